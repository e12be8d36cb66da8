#!/usr/bin/env python3
"""How fast is a marker block that FITS the L2 re-read? Times the plain streaming read (vampomi_time_kernel which=4) and the
A x / A^T p kernels on blocks of 106 250 markers x N rows for small N (27-218 MB), 30 back-to-back launches each: the data
behind DESIGN.md's "one HBM read per CG iteration" idea (row blocks that stay in the 126 MB L2 between A_t p and A_t^T q)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vampomi_b200 as vb  # noqa: E402

M = 106250
for N in (32, 64, 96, 112, 128, 144, 160, 192, 256, 512, 2048):
    sh = vb.Shard(N, M)
    sh.generate_iid(1)
    sh.compute_stats()
    ld = (N + 15) // 16 * 16
    mb = M * ld * 8 / 1e6
    rec = dict(N=N, MB=round(mb, 1))
    for which, name in ((4, "read_probe"), (0, "ax"), (1, "atx"), (5, "ax_2vec"), (6, "atx_2vec")):
        sh.time_kernel(which, 3)
        ms = sh.time_kernel(which, 30)
        rec[name + "_gbs"] = round(mb / ms)          # MB / ms = GB/s
    print(json.dumps(rec), flush=True)
    sh.close()
