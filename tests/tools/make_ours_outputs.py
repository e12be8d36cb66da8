#!/usr/bin/env python3
"""Runs bin/main_meth (GPU) on the inputs of the `linear_small` fixture — inference, `se` association test, out-of-sample test
mode — and keeps ITS output files (CSV bytes as written, the last r1 / p-value vectors) under tests/golden/ours_linear_small/.
They are what tests/test_postprocessing.py feeds to the reference's own post-processing scripts (scripts/p_vals.py,
scripts/metrics.py's readers), which exist only where /root/reference does — i.e. not on the GPU box. Run on a GPU box:
    gpurun -- 'python tests/tools/make_ours_outputs.py gpurun_out/ours_linear_small'   and copy the directory into tests/golden/.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import csv_rows, golden_inputs, load_golden  # noqa: E402
from vampomi_b200 import build, sim  # noqa: E402

dest = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "ours_linear_small")
os.makedirs(dest, exist_ok=True)
g = load_golden("linear_small")
its = int(g["iterations"])
with tempfile.TemporaryDirectory() as d:
    golden_inputs(g, d)
    os.makedirs(f"{d}/out")
    base = ["--meth-file", f"{d}/ex.bin", "--phen-file", f"{d}/ex.phen", "--N", str(g["N"]), "--Mt", str(g["M"]), "--out-dir", f"{d}/out",
            "--out-name", "g"]

    def run(args):
        res = subprocess.run([build.MAIN_METH] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert res.returncode == 0, res.stdout[-2000:]

    run(base + ["--iterations", str(its), "--true-signal-file", f"{d}/ex_ts.bin", "--stop-criteria-thr", "0", "--seed", str(g["probe_seed"])])
    gam1_last = csv_rows(open(f"{d}/out/g_params.csv", "rb").read())[its][1]
    run(base + ["--run-mode", "association_test", "--pval-method", "se", "--r1-file", f"{d}/out/g_r1_it_{its}.bin", "--gam1", repr(gam1_last)])
    Nt = int(g["N_test"])
    sim.write_dataset(d, "tst", Nt, int(g["M"]), float(g["lam"]), float(g["h2"]), int(g["data_seed"]) + 1000)
    run(["--meth-file-test", f"{d}/tst.bin", "--phen-file-test", f"{d}/tst.phen", "--N-test", str(Nt), "--Mt", str(g["M"]), "--out-dir", f"{d}/out",
         "--out-name", "g", "--run-mode", "test", "--estimate-file", f"{d}/out/g_it_1.bin", "--test-iter-range", f"1,{its}"])
    for f in ("g_params.csv", "g_metrics.csv", "g_prior.csv", "g_test.csv", f"g_r1_it_{its}.bin", f"g_it_{its}_pval_se.bin"):
        shutil.copy(f"{d}/out/{f}", os.path.join(dest, f))
print("wrote", sorted(os.listdir(dest)))
