#!/usr/bin/env python3
"""The fused pass beyond 20 480 samples (16-CTA clusters) against the two multi-vector passes it replaces, and at other sample counts:
python tools/gram_large_n.py  ->  one JSON line per (N, M)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vampomi_b200 as vb  # noqa: E402

for N, M in ((40000, 53000), (32000, 66000), (24000, 88000), (20000, 106250), (12000, 177000), (5000, 425000)):
    sh = vb.Shard(N, M)
    sh.generate_iid(1)
    sh.compute_stats()
    gb = N * M * 8 / 1e9
    out = {"N": N, "M": M, "GB": round(gb, 2)}
    for name, which in (("gram_K2", 9), ("gram_K1", 10), ("ax_multi_K2", 5), ("atx_multi_K2", 6)):
        sh.time_kernel(which, 3)
        ms = sh.time_kernel(which, 20)
        out[name] = {"ms": round(ms, 3), "gbs": round(gb / ms * 1e3)}
    out["fused_over_two_passes"] = round((out["ax_multi_K2"]["ms"] + out["atx_multi_K2"]["ms"]) / out["gram_K2"]["ms"], 2)
    print(json.dumps(out), flush=True)
    sh.close()
