#!/usr/bin/env python3
"""Times the matrix kernels over their tuning knobs on one GPU (CUDA events inside the library, vampomi_time_kernel)
and prints GB/s per variant. Usage: python tools/sweep.py [--N 20000 --M 106250 --reps 10] > gpurun_out/sweep.txt"""
import argparse
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import vampomi_b200 as vb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=20000)
ap.add_argument("--M", type=int, default=106250)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--quick", action="store_true")
a = ap.parse_args()

sh = vb.Shard(a.N, a.M)
sh.generate_iid(1)
sh.compute_stats()
rng = np.random.default_rng(0)
sh.Ax(rng.standard_normal(a.M))
sh.ATx(rng.standard_normal(a.N))
gb = a.N * a.M * 8 / 1e9
results = []


def t(which, label, **knobs):
    for k, v in knobs.items():
        sh.set_tuning(k, v)
    sh.time_kernel(which, 2)
    ms = sh.time_kernel(which, a.reps)
    rec = dict(kernel=label, ms=ms, gbs=gb / (ms * 1e-3), **knobs)
    results.append(rec)
    print(json.dumps(rec), flush=True)


ax_variants = [(1, 2), (1, 4), (1, 8), (2, 2), (2, 4), (2, 8), (4, 2), (4, 4)]
atx_variants = [(1, 2), (1, 4), (1, 8), (2, 2), (2, 4), (2, 8), (4, 2), (4, 4)]
occ = [0] if a.quick else [0, 1, 2, 3, 4, 6, 8]
for (rv, u), o in itertools.product(ax_variants, occ):
    t(0, "ax", ax_rv=rv, ax_unroll=u, ax_ctas_per_sm=o)
for (c, u), o in itertools.product(atx_variants, occ):
    t(1, "atx", atx_cols=c, atx_unroll=u, atx_ctas_per_sm=o)
for o in ([0] if a.quick else [0, 1, 2, 3]):
    t(0, "ax", ax_impl=1, ax_ctas_per_sm=o)
    t(1, "atx", atx_impl=1, atx_ctas_per_sm=o)
sh.set_tuning("ax_impl", 0)
sh.set_tuning("atx_impl", 0)
t(2, "stats")
t(3, "loo")
best = {k: max((r for r in results if r["kernel"] == k), key=lambda r: r["gbs"]) for k in ("ax", "atx")}
print("BEST", json.dumps(best))
