#!/usr/bin/env python3
"""Ingest path: writes a marker-major FP64 file and times vampomi_load_file (pread -> pinned ring ->
HBM) for several reader-thread counts. Usage: python tools/load_bench.py [--N 20000 --M 50000]"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vampomi_b200 as vb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=20000)
ap.add_argument("--M", type=int, default=50000)
a = ap.parse_args()
d = tempfile.mkdtemp(prefix="vampomi_load_", dir="/tmp")
path = os.path.join(d, "m.bin")
sh = vb.Shard(a.N, a.M)
sh.generate_iid(1)
with open(path, "wb") as f:
    step = max(1, (256 << 20) // (a.N * 8))
    for j0 in range(0, a.M, step):
        sh.download(j0, min(step, a.M - j0)).tofile(f)
gb = a.N * a.M * 8 / 1e9
ref = sh.download(0, 64)
res = []
for threads in (1, 2, 4, 8, 16):
    sh.set_tuning("load_threads", threads)
    t = time.time()
    sh.load_file(path)
    dt = time.time() - t
    assert np.array_equal(sh.download(0, 64), ref)
    res.append(dict(threads=threads, s=round(dt, 3), gbs=round(gb / dt, 2)))
print(json.dumps(dict(file_gb=gb, note="file just written: reads come from the page cache", results=res)))
os.unlink(path)
os.rmdir(d)
