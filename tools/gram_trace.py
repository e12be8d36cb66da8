#!/usr/bin/env python3
"""Needs a library built with the experiment shapes: make -C vampomi_b200/csrc clean all EXTRA_NVFLAGS=-DVAMPOMI_GRAM_EXPERIMENTS
Hand-over time stamps of the fused pass (gram_shape 15 = shape 11 with clock64() stamps written by cluster 0's rank 0 instead of the
products): where a step's ~1300 SM cycles go. Prints, over the traced steps, the mean and quartiles of every leg.
slots per step: 0-9 compute warp w finished its dot (arrives on REDBAR); 10-19 warp w saw WREADY of the step; 20 / 21 warp 0 before / after the
ring barrier; 24 communication warp saw REDBAR; 25 sent; 26 saw FULL; 27 set WREADY."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vampomi_b200 as vb  # noqa: E402
from vampomi_b200 import capi  # noqa: E402

N, M = 20000, 106250
sh = vb.Shard(N, M)
sh.generate_iid(1)
sh.compute_stats()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
base, traced = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (11, 15)    # 11 / 15: lane sums through shared memory; 16 / 17: quadruple sums
sh.set_tuning("gram_shape", base)
sh.time_kernel(9 if K == 2 else 10, 200)                         # heat up to the sustained clock
sh.set_tuning("gram_shape", traced)
sh.time_kernel(9 if K == 2 else 10, 1)
t = sh.get(capi.V_TMP_M1).view(np.int64)[: 3000 * 32].reshape(3000, 32)
t = t[200:2900].astype(np.float64)                               # steady state
W = 10
dot_done = t[:, 0:W]
wr_seen = t[:, 10:10 + W]
red_seen, sent, full_seen, wr_set = t[:, 24], t[:, 25], t[:, 26], t[:, 27]


def q(name, x):
    x = np.asarray(x).ravel()
    print(json.dumps({"leg": name, "mean": round(float(x.mean()), 1), "p25": float(np.percentile(x, 25)), "p50": float(np.percentile(x, 50)),
                      "p75": float(np.percentile(x, 75)), "p95": float(np.percentile(x, 95))}))


q("step period (communication warp, REDBAR seen s -> s+1)", np.diff(red_seen))
q("last dot done -> REDBAR seen by the communication warp", red_seen - dot_done.max(1))
q("first dot done -> last dot done (spread over the compute warps)", dot_done.max(1) - dot_done.min(1))
q("REDBAR seen -> sent (partial sums read, added, st.async)", sent - red_seen)
q("sent -> FULL seen (round trip, slowest rank of the cluster)", full_seen - sent)
q("FULL seen -> WREADY set", wr_set - full_seen)
q("WREADY set -> seen by the compute warps", wr_seen - wr_set[:, None])
q("REDBAR seen -> WREADY set (communication chain of a step)", wr_set - red_seen)
q("WREADY set (s) -> REDBAR seen (s+1): the communication warp idle", red_seen[1:] - wr_set[:-1])
q("per warp: dot done (s+1) -> WREADY seen (s): wait before the deferred axpy", wr_seen[:-1] - dot_done[1:])
q("per warp: WREADY seen (s) -> dot done (s+2): axpy + ring wait + dot", dot_done[2:] - wr_seen[:-2])
q("warp 0: ring barrier wait", t[:, 21] - t[:, 20])
q("warp 0: ring barrier passed -> dot done", t[:, 0] - t[:, 21])
for w in range(W):
    q(f"warp {w}: dot done relative to the step's last", dot_done[:, w] - dot_done.max(1))
