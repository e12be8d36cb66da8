#!/usr/bin/env python3
"""Smallest run that touches every kernel once, for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
both storages, every A^T p / A x implementation incl. the bulk-copy pipelines, CG (cold, warm, onsager), denoiser,
EM sums, probit z-channel, dots, se / loo, generator, file ingest."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vampomi_b200 as vb  # noqa: E402
from vampomi_b200 import capi  # noqa: E402

N, M = 333, 517
rng = np.random.default_rng(0)
A = rng.standard_normal((M, N))
y = rng.standard_normal(N)
for storage in ("f64", "f32"):
    sh = vb.Shard(N, M, storage=storage)
    sh.upload(A)
    sh.compute_stats()
    p, x = rng.standard_normal(N), rng.standard_normal(M)
    for knobs in (dict(), dict(atx_impl=0), dict(atx_impl=2, atx_cols=4, atx_unroll=2), dict(ax_impl=1, atx_impl=1), dict(center_split=1, atx_impl=0),
                  dict(ax_rv=4, ax_unroll=2, ax_impl=0)):
        for k, v in knobs.items():
            sh.set_tuning(k, v)
        sh.Ax(x); sh.ATx(p)
    for k, v in dict(ax_impl=0, atx_impl=3, center_split=0, ax_rv=0, ax_unroll=0, atx_cols=0, atx_unroll=0).items():
        sh.set_tuning(k, v)
    sh.set(capi.V_V, x)
    sh.cg_solve(capi.V_V, capi.V_X2, 2.0, 1.5, tol=1e-7)
    sh.cg_solve(capi.V_V, capi.V_X2, 2.0, 1.5, warm_start=True, tol=1e-9)
    sh.cg_solve(capi.V_V, capi.V_QINV_BERN, 2.0, 1.5, onsager_mode=True)
    sol = vb.Solver(sh, y, true_signal=x * 0.01, gam1=1e-2, seed=1)
    for _ in range(3):
        sol.step()
    sol.close()
    yb = (y > 0).astype(float)
    sol = vb.Solver(sh, yb, model="bin_class", true_signal=x * 0.01, gam1=1e-2, seed=1)
    for _ in range(2):
        sol.step()
    sol.close()
    sh.pvals_se(x, 2.0)
    sh.set(capi.V_USER_N1, y)
    sh.loo_sums(capi.V_USER_N1)
    sh.generate_iid(3)
    with tempfile.TemporaryDirectory() as d:
        A.tofile(os.path.join(d, "a.bin"))
        sh.load_file(os.path.join(d, "a.bin"))
    sh.close()
print("sanitize case done")
