#!/usr/bin/env python3
"""Regenerates tests/golden/*.npz from the reference itself (oracle/_ref/main_meth_ref, built by oracle/build_ref.py
from /root/reference with the four scripted patches). Runs only in the build container (the GPU box has no
/root/reference); the fixtures are committed. Inputs are NOT stored: tests rebuild them from the seeds recorded in
each fixture with vampomi_b200.sim.simulate and verify the recorded SHA-256 first.
"""
import hashlib
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402
from oracle import vamp_oracle as vo  # noqa: E402
from vampomi_b200 import sim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = {
    "linear_small": dict(N=300, M=800, lam=0.05, h2=0.6, data_seed=11, probe_seed=3, iterations=6, model="linear", extra=[]),
    "linear_readme": dict(N=1000, M=2000, lam=0.1, h2=0.8, data_seed=1234, probe_seed=7, iterations=10, model="linear", extra=[]),
    "linear_ragged": dict(N=333, M=517, lam=0.05, h2=0.5, data_seed=5, probe_seed=9, iterations=5, model="linear",
                          extra=["--EM-max-iter", "3", "--learn-prior-delay", "2", "--rho", "0.7"]),
    "linear_wellcond": dict(N=300, M=800, lam=0.05, h2=0.6, data_seed=11, probe_seed=3, iterations=6, model="linear",
                            extra=["--gam1", "1e-2"]),
    "linear_two_comp": dict(N=200, M=300, lam=0.1, h2=0.7, data_seed=31, probe_seed=2, iterations=5, model="linear",
                            extra=["--vars", "0,0.001", "--probs", "0.9,0.1", "--learn-vars", "0", "--gam1", "1e-2"]),
    "linear_alpha_scale": dict(N=250, M=400, lam=0.1, h2=0.7, data_seed=32, probe_seed=4, iterations=4, model="linear",
                               extra=["--alpha-scale", "0.5", "--gam1", "1e-2", "--merge-vars-thr", "0.9"]),
    "linear_stops_early": dict(N=300, M=200, lam=0.1, h2=0.8, data_seed=33, probe_seed=6, iterations=40, model="linear",
                               extra=["--gam1", "1e-2"], stop_thr="0.01"),
    "linear_warm_start": dict(N=200, M=300, lam=0.1, h2=0.7, data_seed=34, probe_seed=8, iterations=3, model="linear",
                              extra=["--gam1", "1e-2"], warm_from=3),
    "probit_ragged": dict(N=333, M=217, lam=0.1, h2=0.5, data_seed=35, probe_seed=10, iterations=5, model="bin_class",
                          extra=["--gam1", "1e-2", "--rho", "0.8"]),
    # the headline's aspect ratio (Mt/N = 42.5) in small: many more markers than samples, sparse effects, CLI-default prior
    "linear_wide": dict(N=100, M=4000, lam=0.01, h2=0.5, data_seed=41, probe_seed=15, iterations=6, model="linear",
                        extra=["--gam1", "1e-2"]),
    # (probit needs a start that keeps tau1 off its 1e-11 clamp: once the z-channel precision collapses, 1 - alpha2 cancels to
    # ~1e-13 and no two builds of the reference agree any more — N=160, M=2400, --gam1 1e-2 is such a case)
    "probit_wide": dict(N=200, M=3000, lam=0.01, h2=0.5, data_seed=44, probe_seed=17, iterations=4, model="bin_class",
                        extra=["--gam1", "1e-1"]),
    # the CG iteration cap ends both solves (done = 3), a tight tolerance makes them long and unequal, and the EM loop
    # runs to its own convergence test
    "linear_cg_cap": dict(N=200, M=500, lam=0.05, h2=0.6, data_seed=51, probe_seed=21, iterations=4, model="linear",
                          extra=["--gam1", "1e-2", "--CG-max-iter", "3"]),
    "linear_tight_cg": dict(N=220, M=450, lam=0.05, h2=0.7, data_seed=52, probe_seed=22, iterations=4, model="linear",
                            extra=["--gam1", "1e-2", "--CG-err-tol", "1e-9", "--rho", "1.0"]),
    "linear_em_conv": dict(N=240, M=480, lam=0.08, h2=0.6, data_seed=53, probe_seed=23, iterations=5, model="linear",
                           extra=["--gam1", "1e-2", "--EM-max-iter", "8", "--EM-err-thr", "0.05", "--learn-prior-delay", "0"]),
    "linear_h2": dict(N=210, M=420, lam=0.1, h2=0.8, data_seed=54, probe_seed=24, iterations=4, model="linear",
                      extra=["--gam1", "1e-2", "--h2", "0.8", "--rho", "0.3"]),
    # the headline's aspect ratio AND the CLI-default --gam1 1e-6 (the headline benchmark's own start): held to the measured
    # distance between the reference's two builds (x1_O2 / r1_O2), see tests/helpers.py
    "linear_wide_default": dict(N=100, M=4000, lam=0.01, h2=0.5, data_seed=41, probe_seed=15, iterations=6, model="linear", extra=[]),
    # gam2 / gamw ~ 4e6: the regime in which products recycled from the solves' residuals would cancel catastrophically
    # (the recycled / onepass schedules recompute them by explicit passes there), with a tight CG tolerance
    "linear_large_gam2": dict(N=220, M=450, lam=0.05, h2=0.7, data_seed=55, probe_seed=27, iterations=5, model="linear",
                              extra=["--gam1", "1e4", "--CG-err-tol", "1e-10"]),
    # covariates (--C / --cov-file; oracle patch P5 makes the reference's main load them): effects fitted in iteration 1
    "linear_cov": dict(N=240, M=400, lam=0.05, h2=0.6, data_seed=61, probe_seed=25, iterations=4, model="linear",
                       extra=["--gam1", "1e-2"], C=3),
    "probit_cov": dict(N=300, M=400, lam=0.05, h2=0.5, data_seed=62, probe_seed=26, iterations=4, model="bin_class",
                       extra=["--gam1", "1e-2"], C=2),
    "probit_small": dict(N=400, M=600, lam=0.05, h2=0.5, data_seed=21, probe_seed=13, iterations=6, model="bin_class",
                         extra=["--gam1", "1e-2"]),
}


def run_ref(args, seed, threads=4, binary=None):
    env = dict(os.environ, VAMPOMI_SEED=str(seed), OMP_NUM_THREADS=str(threads))
    res = subprocess.run([binary or build_ref.OUT_BIN] + args, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(res.stdout[-3000:])
    return res.stdout


def cg_counts(log, tol=1e-5, max_iter=500):
    """Per VAMP iteration (k1, k2) from the --verbosity 1 log (src/vamp.cpp:696-751). Every executed iteration prints one
    '[CG] it = i' line, except the one in which the onsager solve's own test fires: that one breaks before both prints
    (:718-719). So k1 = number of '[CG]' lines, and k2 = that number + 1 unless the solve ended on the residual test (last
    printed residual < tol) or on the iteration cap."""
    out = []
    for block in log.split("iteration = ")[1:]:
        lm, ons = block.split("CG took")[0], block.split("CG took")[1].split("onsager took")[0]
        k1 = len(re.findall(r"\[CG\] it = ", lm))
        n_cg = len(re.findall(r"\[CG\] it = ", ons))
        last_rel = re.findall(r"\|\|r_it\|\| / \|\|RHS\|\| = ([0-9.e+-]+)", ons)
        ended_by_residual = bool(last_rel) and float(last_rel[-1]) < tol
        k2 = n_cg if (ended_by_residual or n_cg >= max_iter) else n_cg + 1
        out.append((k1, k2))
    return np.array(out, dtype=np.int64)


def cg_counts_probit(log, tol=1e-5, max_iter=500):
    """The probit loop (src/vamp_probit.cpp:296-311) prints no timing lines between its two solves; the Onsager solve marks its
    iterations with '[CG onsager] it = i' (src/vamp.cpp:723-724), so the '[CG] it' lines before the first of those belong to
    the LMMSE solve and the ones after it to the Onsager solve. Same rules as cg_counts for k2."""
    out = []
    for block in log.split("iteration = ")[1:]:
        parts = block.split("[CG onsager]", 1)
        lm, ons = parts[0], (parts[1] if len(parts) > 1 else "")
        k1 = len(re.findall(r"\[CG\] it = ", lm))
        n_cg = len(re.findall(r"\[CG\] it = ", ons))
        last_rel = re.findall(r"\|\|r_it\|\| / \|\|RHS\|\| = ([0-9.e+-]+)", ons)
        ended_by_residual = bool(last_rel) and float(last_rel[-1]) < tol
        k2 = n_cg if (ended_by_residual or n_cg >= max_iter) else n_cg + 1
        out.append((k1, k2))
    return np.array(out, dtype=np.int64)


def make_case(name, c):
    with tempfile.TemporaryDirectory() as d:
        X, y, beta = sim.write_dataset(d, "ex", c["N"], c["M"], c["lam"], c["h2"], c["data_seed"], binary=c["model"] == "bin_class")
        os.makedirs(os.path.join(d, "out"))
        cov = None
        if c.get("C"):
            cov, y = sim.simulate_covariates(c["N"], c["C"], c["data_seed"], y=y, binary=c["model"] == "bin_class")
            sim.write_phen(f"{d}/ex.phen", y)
            sim.write_covariates(f"{d}/ex.cov", cov)
        args = ["--meth-file", f"{d}/ex.bin", "--phen-file", f"{d}/ex.phen", "--N", str(c["N"]), "--Mt", str(c["M"]),
                "--out-dir", f"{d}/out", "--out-name", "g", "--iterations", str(c["iterations"]), "--true-signal-file",
                f"{d}/ex_ts.bin", "--model", c["model"], "--stop-criteria-thr", c.get("stop_thr", "0"), "--verbosity", "1"] + c["extra"]
        if c.get("C"):
            args += ["--C", str(c["C"]), "--cov-file", f"{d}/ex.cov"]
        init = None
        if c.get("warm_from"):
            # --estimate-file start (src/main_meth.cpp:75-80, src/vamp.cpp:71-79 with patch P1): first produce an estimate
            os.makedirs(os.path.join(d, "pre"))
            pre = [a_.replace(f"{d}/out", f"{d}/pre") for a_ in args]
            run_ref(pre, c["probe_seed"])
            init = np.fromfile(f"{d}/pre/g_it_{c['warm_from']}.bin")
            args += ["--estimate-file", f"{d}/pre/g_it_{c['warm_from']}.bin"]
        log = run_ref(args, c["probe_seed"])
        ran = len([f for f in os.listdir(f"{d}/out") if f.startswith("g_it_")])
        c = dict(c, iterations=ran)          # early stop: fewer iterations than asked for
        x1 = np.stack([np.fromfile(f"{d}/out/g_it_{k}.bin") for k in range(1, c["iterations"] + 1)])
        r1 = np.stack([np.fromfile(f"{d}/out/g_r1_it_{k}.bin") for k in range(1, c["iterations"] + 1)])
        blob = {k: np.frombuffer(open(f"{d}/out/g_{k}.csv", "rb").read(), dtype=np.uint8) for k in ("params", "metrics", "prior")}
        fix = dict(x1=x1, r1=r1, csv_params=blob["params"], csv_metrics=blob["metrics"], csv_prior=blob["prior"],
                   N=c["N"], M=c["M"], lam=c["lam"], h2=c["h2"], data_seed=c["data_seed"], probe_seed=c["probe_seed"],
                   iterations=c["iterations"], model=c["model"], extra=np.array(c["extra"], dtype="U32"),
                   sha256_A=hashlib.sha256(X.tobytes()).hexdigest(), sha256_phen=hashlib.sha256(open(f"{d}/ex.phen", "rb").read()).hexdigest())
        fix["stop_thr"] = float(c.get("stop_thr", "0"))
        if cov is not None:
            fix["C"] = c["C"]
            fix["cov_eff"] = np.array([float(v) for v in re.findall(r"cov_eff\[\d+\] = ([-+0-9.eE]+|-?nan|-?inf)", log)][:c["C"]])
        if init is not None:
            fix["x1hat_init"] = init
        ex = dict(zip(c["extra"][::2], c["extra"][1::2]))
        if c["model"] == "linear":
            fix["cg_iters"] = cg_counts(log, float(ex.get("--CG-err-tol", 1e-5)), int(ex.get("--CG-max-iter", 500)))
        else:
            fix["cg_iters"] = cg_counts_probit(log, float(ex.get("--CG-err-tol", 1e-5)), int(ex.get("--CG-max-iter", 500)))
        if "--gam1" not in ex and name != "linear_small":
            # CLI-default gam1 = 1e-6: the same run by the IEEE-strict (-O2) build, vectors and params rows (see linear_small below)
            os.makedirs(os.path.join(d, "out2"))
            args2 = [a.replace(f"{d}/out", f"{d}/out2") for a in args]
            run_ref(args2, c["probe_seed"], binary=build_ref.OUT_BIN_STRICT)
            fix["x1_O2"] = np.stack([np.fromfile(f"{d}/out2/g_it_{k}.bin") for k in range(1, c["iterations"] + 1)])
            fix["r1_O2"] = np.stack([np.fromfile(f"{d}/out2/g_r1_it_{k}.bin") for k in range(1, c["iterations"] + 1)])
            fix["csv_params_O2"] = np.frombuffer(open(f"{d}/out2/g_params.csv", "rb").read(), dtype=np.uint8)
        if name == "linear_small":
            # the same run by an IEEE-strict (-O2) build of the same patched sources: how far the reference is from
            # ITSELF when only compiler flags change — the floor any independent implementation can be held to
            os.makedirs(os.path.join(d, "out2"))
            args2 = [a.replace(f"{d}/out", f"{d}/out2") for a in args]
            run_ref(args2, c["probe_seed"], binary=build_ref.OUT_BIN_STRICT)
            fix["x1_O2"] = np.stack([np.fromfile(f"{d}/out2/g_it_{k}.bin") for k in range(1, c["iterations"] + 1)])
            fix["r1_O2"] = np.stack([np.fromfile(f"{d}/out2/g_r1_it_{k}.bin") for k in range(1, c["iterations"] + 1)])
            fix["csv_params_O2"] = np.frombuffer(open(f"{d}/out2/g_params.csv", "rb").read(), dtype=np.uint8)
            # association tests and out-of-sample test mode from this run's saved files (src/main_meth.cpp:112-265)
            params = vo.read_csv_rows(f"{d}/out/g_params.csv")
            last = c["iterations"]
            gam1_last = params[last][1]
            run_ref(["--meth-file", f"{d}/ex.bin", "--phen-file", f"{d}/ex.phen", "--N", str(c["N"]), "--Mt", str(c["M"]),
                     "--out-dir", f"{d}/out", "--out-name", "g", "--run-mode", "association_test", "--pval-method", "se",
                     "--r1-file", f"{d}/out/g_r1_it_{last}.bin", "--gam1", repr(gam1_last)], c["probe_seed"])
            run_ref(["--meth-file", f"{d}/ex.bin", "--phen-file", f"{d}/ex.phen", "--N", str(c["N"]), "--Mt", str(c["M"]),
                     "--out-dir", f"{d}/out", "--out-name", "g", "--run-mode", "association_test", "--pval-method", "loo",
                     "--estimate-file", f"{d}/out/g_it_{last}.bin"], c["probe_seed"])
            fix["se_gam1"] = gam1_last
            fix["pval_se"] = np.fromfile(f"{d}/out/g_it_{last}_pval_se.bin")
            fix["pval_loo"] = np.fromfile(f"{d}/out/g_it_{last}_pval_loo.bin")
            Nt = 200
            sim.write_dataset(d, "tst", Nt, c["M"], c["lam"], c["h2"], c["data_seed"] + 1000)
            run_ref(["--meth-file-test", f"{d}/tst.bin", "--phen-file-test", f"{d}/tst.phen", "--N-test", str(Nt), "--Mt", str(c["M"]),
                     "--out-dir", f"{d}/out", "--out-name", "g", "--run-mode", "test", "--estimate-file", f"{d}/out/g_it_1.bin",
                     "--test-iter-range", f"1,{last}"], c["probe_seed"])
            fix["N_test"] = Nt
            fix["csv_test"] = np.frombuffer(open(f"{d}/out/g_test.csv", "rb").read(), dtype=np.uint8)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **fix)
        print(name, "ok:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in fix.items() if k in ("x1", "cg_iters", "csv_params")})


if __name__ == "__main__":
    if not build_ref.build() or not build_ref.build(strict=True):
        sys.exit("oracle/_ref is not available (no /root/reference here?)")
    os.makedirs(GOLDEN, exist_ok=True)
    for n, c in CASES.items():
        if len(sys.argv) > 1 and n not in sys.argv[1:]:
            continue
        make_case(n, c)
