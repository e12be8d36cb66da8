#!/usr/bin/env python3
"""Full-size parity of BASELINE config 2 (N=10000, M=100000, 8 GB) against the reference binary itself.

Runs on the GPU box: the device-generated matrix is downloaded and written as a marker-major .bin, the phenotype is
simulated from it, then BOTH command lines read the same files: vampomi_b200/bin/main_meth (GPU) and
oracle/_ref/main_meth_ref (the patched reference, all host cores). Reports per-iteration relative differences of
x1_hat / r1, CSV values, CG iteration counts and wall times. Usage: python tools/parity_c2.py [--N 10000 --M 100000 --its 3]
"""
import argparse
import json
import math
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vampomi_b200 as vb  # noqa: E402
from helpers import csv_rows, rel_l2  # noqa: E402
from vampomi_b200 import build, sim  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=10000)
ap.add_argument("--M", type=int, default=100000)
ap.add_argument("--its", type=int, default=3)
ap.add_argument("--dir", default=None)
a = ap.parse_args()
N, M = a.N, a.M
work = a.dir or tempfile.mkdtemp(prefix="vampomi_c2_", dir="/tmp")
os.makedirs(work, exist_ok=True)
report = {"config": f"N={N} M={M} ({N * M * 8 / 1e9:.1f} GB), linear, {a.its} iterations", "runs": []}

t0 = time.time()
sh = vb.Shard(N, M)
sh.generate_iid(99)
sh.compute_stats()
rng = np.random.default_rng(99)
CM = max(int(M * 0.01), 1)
beta = np.zeros(M)
beta[rng.choice(M, CM, replace=False)] = rng.normal(0, math.sqrt(0.5 / CM), CM)
y = sh.Ax(beta * math.sqrt(N)) + rng.normal(0, math.sqrt(0.5), N)
with open(f"{work}/c2.bin", "wb") as f:
    step = max(1, (256 << 20) // (N * 8))
    for j0 in range(0, M, step):
        sh.download(j0, min(step, M - j0)).tofile(f)
sim.write_phen(f"{work}/c2.phen", y)
beta.tofile(f"{work}/c2_ts.bin")
sh.close()
report["prepare_s"] = round(time.time() - t0, 1)

ref_bin = os.path.join(ROOT, "oracle", "_ref", "main_meth_ref")
for tag, extra, tol in (("default_gam1", [], 1e-7), ("gam1_1e-2", ["--gam1", "1e-2"], 1e-9)):
    outs = {}
    for who, binary, more, env in (("gpu", build.MAIN_METH, ["--seed", "5"], {}),
                                   ("ref", ref_bin, ["--verbosity", "1"], {"VAMPOMI_SEED": "5", "OMP_NUM_THREADS": str(os.cpu_count())})):
        od = f"{work}/out_{tag}_{who}"
        os.makedirs(od, exist_ok=True)
        cmd = [binary, "--meth-file", f"{work}/c2.bin", "--phen-file", f"{work}/c2.phen", "--N", str(N), "--Mt", str(M), "--out-dir", od,
               "--out-name", "c2", "--iterations", str(a.its), "--true-signal-file", f"{work}/c2_ts.bin", "--stop-criteria-thr", "0"] + extra + more
        t = time.time()
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=dict(os.environ, **env))
        outs[who] = dict(rc=res.returncode, wall_s=round(time.time() - t, 2), log=res.stdout)
        if res.returncode != 0:
            print(res.stdout[-2000:])
    run = {"case": tag, "tolerance": tol, "gpu_wall_s": outs["gpu"]["wall_s"], "ref_wall_s": outs["ref"]["wall_s"]}
    run["gpu_iter_s"] = [float(x) for x in re.findall(r"Total iteration time = ([0-9.eE+-]+)", outs["gpu"]["log"])]
    run["ref_iter_s"] = [float(x) for x in re.findall(r"Total iteration time = ([0-9.eE+-]+)", outs["ref"]["log"])]
    load = re.findall(r"reading methylation data took ([0-9.eE+-]+)", outs["gpu"]["log"])
    run["gpu_load_s"] = float(load[0]) if load else None
    run["gpu_cg"] = re.findall(r"\[CG\] LMMSE solve: (\d+) iterations, onsager solve: (\d+)", outs["gpu"]["log"])
    devs = []
    for k in range(1, a.its + 1):
        gx, rx = (np.fromfile(f"{work}/out_{tag}_{w}/c2_it_{k}.bin") for w in ("gpu", "ref"))
        gr, rr = (np.fromfile(f"{work}/out_{tag}_{w}/c2_r1_it_{k}.bin") for w in ("gpu", "ref"))
        devs.append([rel_l2(gx, rx), rel_l2(gr, rr)])
    run["rel_l2_x1_r1_per_iteration"] = devs
    gp, rp = (csv_rows(open(f"{work}/out_{tag}_{w}/c2_params.csv", "rb").read()) for w in ("gpu", "ref"))
    run["max_rel_params"] = max(abs(x - y_) / abs(y_) for k in rp for x, y_ in zip(gp[k], rp[k]) if y_ != 0)
    sizes = [os.path.getsize(f"{work}/out_{tag}_{w}/c2_{c}.csv") for w in ("gpu", "ref") for c in ("params", "metrics", "prior")]
    run["csv_sizes_equal"] = sizes[:3] == sizes[3:]
    run["pass"] = bool(all(max(d) < tol for d in devs) and run["csv_sizes_equal"])
    report["runs"].append(run)
print(json.dumps(report, indent=1))
if a.dir is None:
    subprocess.run(["rm", "-rf", work])
