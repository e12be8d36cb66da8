#!/usr/bin/env python3
"""Measured parity of the CUDA path against every fixture produced by the reference binary (tests/golden/*.npz): the
largest relative L2 deviation of x1_hat / r1 over the iterations, the largest relative deviation of the params / metrics
CSV values, and whether the CG iteration counts are identical. Prints one JSON object (kept under profiles/)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import csv_rows, extra_kwargs, golden_covariates, golden_inputs, load_golden, rel_l2, standardize_phen, tolerances  # noqa: E402
from vampomi_b200 import capi  # noqa: E402

CASES = ["linear_wide_default", "linear_large_gam2", "linear_cov", "probit_cov", "linear_small", "linear_readme", "linear_ragged", "linear_wellcond", "linear_two_comp", "linear_alpha_scale",
         "linear_stops_early", "linear_warm_start", "probit_small", "probit_ragged", "linear_wide", "probit_wide", "linear_cg_cap", "linear_tight_cg",
         "linear_em_conv", "linear_h2"]
SCHED = {"onepass": 3, "recycled": 2, "fused": 1, "plain": 0}
schedule = sys.argv[1] if len(sys.argv) > 1 else "onepass"
out = {"_schedule": schedule}
for name in CASES:
    g = load_golden(name)
    A, y_txt, beta = golden_inputs(g)
    model = g["model"]
    y = standardize_phen(y_txt) if model == "linear" else y_txt
    kw = dict(gamw=2.0, seed=int(g["probe_seed"]), fuse_passes=SCHED[schedule])
    kw.update(extra_kwargs(g))
    if "h2" in kw:
        kw["gamw"] = 1.0 / (1.0 - kw.pop("h2"))              # src/main_meth.cpp:52
    sh = capi.Shard(int(g["N"]), int(g["M"]))
    sh.upload(A)
    sh.compute_stats(kw.pop("alpha_scale", 1.0))
    sol = capi.Solver(sh, y, model=model, true_signal=beta, x1hat_init=g.get("x1hat_init"), **kw)
    if "C" in g:
        sol.set_covariates(golden_covariates(g))
    want_p, want_m = csv_rows(g["csv_params"]), csv_rows(g["csv_metrics"])
    dev_vec, dev_csv, cg_same, dev_o2 = 0.0, 0.0, True, 0.0
    for k in range(1, int(g["iterations"]) + 1):
        r = sol.step()
        dev_vec = max(dev_vec, rel_l2(r["x1"], g["x1"][k - 1]), rel_l2(r["r1"], g["r1"][k - 1]))
        if "x1_O2" in g:                                     # default-gam1 fixtures: distance to the IEEE-strict build of the reference
            dev_o2 = max(dev_o2, rel_l2(r["x1"], g["x1_O2"][k - 1]), rel_l2(r["r1"], g["r1_O2"][k - 1]))
        for got, want in ((r["params"], want_p.get(k, [])), (r["metrics"], want_m.get(k, []))):
            for a, b in zip(got, want):
                if b is not None and np.isfinite(b) and b != 0 and abs(b) > 1e-6:      # the CSV keeps 15 decimals: tiny values carry few digits
                    dev_csv = max(dev_csv, abs(a - b) / abs(b))
        cg_same &= (r["k1"], r["k2"]) == tuple(g["cg_iters"][k - 1])
    tol_vec, tol_csv = tolerances(g)
    out[name] = dict(iterations=int(g["iterations"]), gam1_start=kw.get("gam1", 1e-6), max_rel_l2_x1_r1=dev_vec, max_rel_csv=dev_csv,
                     cg_counts_identical=bool(cg_same), tolerance_vec=tol_vec, tolerance_csv=tol_csv)
    if "x1_O2" in g:
        out[name]["max_rel_l2_vs_reference_O2_build"] = dev_o2
        out[name]["reference_builds_apart"] = max(max(rel_l2(g["x1_O2"][k], g["x1"][k]), rel_l2(g["r1_O2"][k], g["r1"][k])) for k in range(1, int(g["iterations"])))
    sol.close()
    sh.close()
print(json.dumps(out, indent=1))
