#!/usr/bin/env python3
"""Timings (and property checks) of the BASELINE configs that are not the headline bench line, on one GPU:
  C4  probit VAMP (--model bin_class) N=20000, M=400000 synthetic case/control, 5 iterations
  C5  association se / loo and out-of-sample test-mode passes at N=20000, M=850000 from estimates of a short run
Prints one JSON object. Usage: python tools/run_configs.py > gpurun_out/configs.json"""
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vampomi_b200 as vb  # noqa: E402
from vampomi_b200 import capi  # noqa: E402

out = {}


def simulate(sh, N, Mt, seed, binary):
    rng = np.random.default_rng(seed)
    CM = max(int(Mt * 0.01), 1)
    beta = np.zeros(Mt)
    beta[rng.choice(Mt, CM, replace=False)] = rng.normal(0, math.sqrt(0.5 / CM), CM)
    y = sh.Ax(beta * math.sqrt(N)) + rng.normal(0, math.sqrt(0.5), N)
    if binary:
        return (y > 0).astype(np.float64), beta
    return y * math.sqrt((N - 1) / float(((y - y.mean()) ** 2).sum())), beta


# ---- C4 ----
N, Mt = 20000, 400000
sh = vb.Shard(N, Mt)
sh.generate_iid(4)
sh.compute_stats()
y, beta = simulate(sh, N, Mt, 4, True)
sol = vb.Solver(sh, y, model="bin_class", true_signal=beta, gam1=1e-2, seed=3)
its = []
for k in range(5):
    t = time.time()
    r = sol.step()
    its.append(dict(it=r["it"], s=round(time.time() - t, 4), k1=r["k1"], k2=r["k2"], passes=r["matrix_passes"],
                    acc1=r["metrics"][4], acc2=r["metrics"][10], corr_x2=r["metrics"][11]))
out["schedule"] = "onepass (default)"
out["C4_probit_N20000_M400000"] = dict(iterations=its, gbs=[round(i["passes"] * N * Mt * 8 / i["s"] / 1e9) for i in its],
                                       finite=bool(np.all(np.isfinite(r["x1"]))))
sol.close()
sh.close()

# ---- C5 ----
N, Mt = 20000, 850000
sh = vb.Shard(N, Mt)
sh.generate_iid(5)
sh.compute_stats()
y, beta = simulate(sh, N, Mt, 5, False)
sol = vb.Solver(sh, y, model="linear", true_signal=beta, seed=3)
for k in range(3):
    r = sol.step()
x1s, r1s, gam1 = r["x1"], r["r1"], r["params"][1]
sol.close()
t = time.time(); p_se = sh.pvals_se(r1s, gam1); t_se = time.time() - t
# loo: A x1, residual, one streaming pass for the sums (driver.cpp run_association)
t = time.time()
sh.set(capi.V_Y, y); sh.set(capi.V_X1, x1s * math.sqrt(N)); sh.ax_dev(capi.V_X1, capi.V_Z1)
sh.lincomb(capi.V_USER_N1, 1.0, capi.V_Y, -1.0, capi.V_Z1, 1.0)
t_prep = time.time() - t
sums = sh.loo_sums(capi.V_USER_N1)
t_loo = time.time() - t
loo_kernel_ms = sh.time_kernel(3, 5)
t = time.time(); z = sh.Ax(x1s * math.sqrt(N)); t_test = time.time() - t
# test mode as main_meth runs it: four saved estimates per pass over the test matrix (vampomi_ax_multi_dev)
for v_ in (capi.V_X1, capi.V_X2, capi.V_R1, capi.V_R2):
    sh.set(v_, x1s * math.sqrt(N))
sh.ax_multi_dev([capi.V_X1, capi.V_X2, capi.V_R1, capi.V_R2], [capi.V_Z1, capi.V_Z2, capi.V_USER_N0, capi.V_USER_N1])
sh.get(capi.V_Z1)
t = time.time()
sh.ax_multi_dev([capi.V_X1, capi.V_X2, capi.V_R1, capi.V_R2], [capi.V_Z1, capi.V_Z2, capi.V_USER_N0, capi.V_USER_N1])
z4 = sh.get(capi.V_USER_N1)
t_test4 = time.time() - t
out["C5_assoc_test_N20000_M850000"] = dict(se_s=round(t_se, 4), se_frac_below_0_05=float((p_se < 0.05).mean()),
                                           loo_end_to_end_s=round(t_loo, 4), loo_prepare_s=round(t_prep, 4),
                                           loo_sums_kernel_ms=round(loo_kernel_ms, 3),
                                           loo_sums_kernel_gbs=round(N * Mt * 8 / loo_kernel_ms / 1e6),
                                           test_mode_pass_s=round(t_test, 4), test_mode_four_estimates_pass_s=round(t_test4, 4),
                                           four_pass_rel_l2_vs_single=float(np.linalg.norm(z4 - z) / np.linalg.norm(z)), sums_finite=bool(np.all(np.isfinite(sums))),
                                           sum_x_matches_mean=bool(np.allclose(sums[:1000, 0] / N, sh.stats()[0][:1000], rtol=1e-10, atol=1e-12)))
sh.close()
print(json.dumps(out, indent=1))
