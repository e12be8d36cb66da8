"""End-to-end parity of the VAMP loop on the GPU: per-iteration x1_hat / r1, params / metrics rows and CG iteration
counts against fixtures produced by the reference binary (tests/golden) and against the numpy oracle; the main_meth
command line against the same fixtures, byte layout of the CSVs included; plus full-size property checks."""
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import vamp_oracle as vo
from vampomi_b200 import build, capi, sim
from vampomi_b200.capi import V_QINV_BERN, V_USER_M0, V_USER_M1, V_USER_N0, V_USER_N1, V_V, V_X1, V_X2, V_Y
from helpers import (assert_rows_close, csv_rows, extra_kwargs, golden_inputs, load_golden, oracle_run, rel_l2,
                     standardize_phen, tolerances)

pytestmark = pytest.mark.gpu

CASES = ["linear_small", "linear_readme", "linear_ragged", "linear_wellcond", "linear_two_comp", "linear_alpha_scale",
         "linear_stops_early", "linear_warm_start", "probit_small", "probit_ragged", "linear_wide", "probit_wide", "linear_cg_cap", "linear_tight_cg",
         "linear_em_conv", "linear_h2"]


def solver_for(g, A, y_txt, beta, **over):
    model = g["model"]
    y = standardize_phen(y_txt) if model == "linear" else y_txt
    kw = dict(gamw=2.0, seed=int(g["probe_seed"]))
    kw.update(extra_kwargs(g))
    if "h2" in kw:
        kw["gamw"] = 1.0 / (1.0 - kw.pop("h2"))              # src/main_meth.cpp:52
    kw.update(over)
    sh = capi.Shard(int(g["N"]), int(g["M"]))
    sh.upload(A)
    sh.compute_stats(kw.pop("alpha_scale", 1.0))
    return sh, capi.Solver(sh, y, model=model, true_signal=beta, x1hat_init=g.get("x1hat_init"), **kw)


# schedules of the matrix passes: "onepass" (recycled + CG iterations that read the block once: fused A^T q / A A^T q pass), "recycled" (fused + A x2_hat, A Q^-1 u and A^T A of both kept by the solves themselves),
# "fused" (products that are known together share one read of the block),
# "plain" (one product per pass, A^T y and A x2_hat cached), "reference" (plain + the passes the reference repeats)
SCHEDULES = {"onepass": dict(fuse_passes=3, redundant_passes=0), "recycled": dict(fuse_passes=2, redundant_passes=0), "fused": dict(fuse_passes=1, redundant_passes=0), "plain": dict(fuse_passes=0, redundant_passes=0),
             "reference": dict(fuse_passes=0, redundant_passes=1)}


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("schedule", list(SCHEDULES))
def test_solver_matches_reference_fixture(name, schedule):
    g = load_golden(name)
    rel_vec, rel_csv = tolerances(g)
    A, y_txt, beta = golden_inputs(g)
    sh, sol = solver_for(g, A, y_txt, beta, **SCHEDULES[schedule])
    want_params, want_metrics = csv_rows(g["csv_params"]), csv_rows(g["csv_metrics"])
    got_params, got_metrics = {}, {}
    for k in range(1, int(g["iterations"]) + 1):
        r = sol.step()
        assert r["it"] == k
        assert rel_l2(r["x1"], g["x1"][k - 1]) < rel_vec, f"x1_hat it {k}"
        assert rel_l2(r["r1"], g["r1"][k - 1]) < rel_vec, f"r1 it {k}"
        got_params[k], got_metrics[k] = r["params"], r["metrics"]
        if g["model"] == "linear":
            assert (r["k1"], r["k2"]) == tuple(g["cg_iters"][k - 1]), f"CG iteration counts it {k}"
            base = 2 * (r["k1"] + r["k2"])
            if schedule == "reference":   # the reference's own pass count: 6 (it = 1) / 8 (it > 1) + 2(k1+k2), SURVEY.md §3.1
                assert r["matrix_passes"] == base + (6 if k == 1 else 8)
            elif schedule == "plain":     # A^T y cached, A x2_hat computed once
                assert r["matrix_passes"] == base + (5 if k == 1 else 6)
            elif schedule == "fused":     # A^T y once; both solves in lock-step; A [x2, Q^-1 u] and A^T [.., A x2] one pass each
                assert r["matrix_passes"] == 2 * max(r["k1"], r["k2"]) + (3 if k == 1 else 2)
            elif schedule == "onepass":   # A [p0 p1 x1] once, then ONE fused pass per lock-step CG iteration (and A^T y once)
                assert r["matrix_passes"] == max(r["k1"], r["k2"]) + 1 + (1 if k == 1 else 0)
            else:                         # nothing but the lock-step solves (and A^T y once)
                assert r["matrix_passes"] == 2 * max(r["k1"], r["k2"]) + (1 if k == 1 else 0)
        else:
            if schedule == "fused":       # A^T p2, lock-step solves, A [x2, x2/sqrt(N)]
                assert r["matrix_passes"] == 2 * max(r["k1"], r["k2"]) + 2
            elif schedule == "recycled":  # A^T p2, lock-step solves
                assert r["matrix_passes"] == 2 * max(r["k1"], r["k2"]) + 1
            elif schedule == "onepass":   # A^T p2, A [p0 p1 x1/sqrt(N)], one fused pass per lock-step CG iteration
                assert r["matrix_passes"] == max(r["k1"], r["k2"]) + 2
            else:
                assert r["matrix_passes"] == 2 * (r["k1"] + r["k2"]) + 4
    assert_rows_close(got_params, want_params, rel_csv, "params")
    assert_rows_close(got_metrics, want_metrics, rel_csv, "metrics")
    if float(g.get("stop_thr", 0)) > 0:          # the reference stopped here on its own NMSE test (src/vamp.cpp:419-423)
        assert r["nmse"] < float(g["stop_thr"])
    sol.close()
    sh.close()


@pytest.mark.parametrize("name", ["linear_wellcond", "probit_small"])
def test_solver_matches_oracle_tightly(name):
    """GPU vs the numpy oracle on the well-conditioned fixtures: both restate the same arithmetic, so they agree far
    below the 1e-9 contract."""
    g = load_golden(name)
    A, y_txt, beta = golden_inputs(g)
    v = oracle_run(g, A, y_txt, beta)
    sh, sol = solver_for(g, A, y_txt, beta)
    for k in range(1, int(g["iterations"]) + 1):
        r = sol.step()
        assert rel_l2(r["x1"], v.dump[k][0]) < 1e-10 and rel_l2(r["r1"], v.dump[k][1]) < 1e-10
        assert np.allclose(r["probs"], v.history[k - 1]["probs"], rtol=1e-9)
    sol.close()
    sh.close()


def run_cli(args, **kw):
    res = subprocess.run([build.MAIN_METH] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    assert res.returncode == 0, res.stdout[-3000:]
    return res.stdout


@pytest.mark.parametrize("schedule", ["onepass", "recycled", "fused", "plain", "reference"])
def test_main_meth_schedule_flag(schedule, tmp_path):
    """--schedule only changes which products share a read of the block: same files for all three."""
    g = load_golden("linear_wellcond")
    rel_vec, rel_csv = tolerances(g)
    d = str(tmp_path)
    golden_inputs(g, d)
    os.makedirs(tmp_path / "out")
    its = int(g["iterations"])
    out = run_cli(["--meth-file", f"{d}/ex.bin", "--phen-file", f"{d}/ex.phen", "--N", g["N"], "--Mt", g["M"], "--out-dir", f"{d}/out",
                   "--out-name", "g", "--iterations", its, "--true-signal-file", f"{d}/ex_ts.bin", "--stop-criteria-thr", 0,
                   "--seed", g["probe_seed"], "--schedule", schedule] + list(g["extra"]))
    assert f"--schedule {schedule}" in out
    for k in range(1, its + 1):
        assert rel_l2(np.fromfile(f"{d}/out/g_it_{k}.bin"), g["x1"][k - 1]) < rel_vec
        assert rel_l2(np.fromfile(f"{d}/out/g_r1_it_{k}.bin"), g["r1"][k - 1]) < rel_vec
    for kind in ("params", "metrics"):
        assert_rows_close(csv_rows(open(f"{d}/out/g_{kind}.csv", "rb").read()), csv_rows(bytes(g[f"csv_{kind}"])), rel_csv, kind)
    res = subprocess.run([build.MAIN_METH, "--schedule", "sideways"], stdout=subprocess.PIPE, text=True)
    assert res.returncode == 1 and "FATAL" in res.stdout


@pytest.mark.parametrize("name", ["linear_small", "probit_small", "linear_stops_early", "linear_warm_start", "linear_alpha_scale",
                                  "linear_two_comp"])
def test_main_meth_command_line_outputs(name, tmp_path):
    g = load_golden(name)
    rel_vec, rel_csv = tolerances(g)
    d = str(tmp_path)
    golden_inputs(g, d)
    os.makedirs(tmp_path / "out")
    its = int(g["iterations"])
    early = float(g.get("stop_thr", 0)) > 0
    args = ["--meth-file", f"{d}/ex.bin", "--phen-file", f"{d}/ex.phen", "--N", g["N"], "--Mt", g["M"], "--out-dir", f"{d}/out",
            "--out-name", "g", "--iterations", 40 if early else its, "--true-signal-file", f"{d}/ex_ts.bin", "--model", g["model"],
            "--stop-criteria-thr", g.get("stop_thr", 0), "--seed", g["probe_seed"], "--run-mode", "inference"] + list(g["extra"])
    if "x1hat_init" in g:
        g["x1hat_init"].tofile(f"{d}/init_it_3.bin")
        args += ["--estimate-file", f"{d}/init_it_3.bin"]
    out = run_cli(args)
    assert "iteration = 1" in out and "x1_hat NMSE" in out
    if early:
        assert "...stopping criteria fulfilled" in out and not os.path.exists(f"{d}/out/g_it_{its + 1}.bin")
    for k in range(1, its + 1):
        assert rel_l2(np.fromfile(f"{d}/out/g_it_{k}.bin"), g["x1"][k - 1]) < rel_vec
        assert rel_l2(np.fromfile(f"{d}/out/g_r1_it_{k}.bin"), g["r1"][k - 1]) < rel_vec
    for kind in ("params", "metrics", "prior"):
        got = open(f"{d}/out/g_{kind}.csv", "rb").read()
        want = bytes(g[f"csv_{kind}"])
        assert len(got) == len(want), f"{kind}.csv size"
        if kind != "prior" or g["model"] == "linear":
            assert np.array_equal(np.frombuffer(got, np.uint8) == 0, np.frombuffer(want, np.uint8) == 0), f"{kind}.csv NUL layout"
        if kind == "prior" and g["model"] == "linear":
            assert got == want
        if kind != "prior":
            assert_rows_close(csv_rows(got), csv_rows(want), rel_csv, kind)


def test_association_and_test_modes(tmp_path):
    g = load_golden("linear_small")
    d = str(tmp_path)
    golden_inputs(g, d)
    os.makedirs(tmp_path / "out")
    last = int(g["iterations"])
    for k in range(1, last + 1):
        g["x1"][k - 1].tofile(f"{d}/out/g_it_{k}.bin")
        g["r1"][k - 1].tofile(f"{d}/out/g_r1_it_{k}.bin")
    common = ["--meth-file", f"{d}/ex.bin", "--phen-file", f"{d}/ex.phen", "--N", g["N"], "--Mt", g["M"], "--out-dir", f"{d}/out",
              "--out-name", "g"]
    run_cli(common + ["--run-mode", "association_test", "--pval-method", "se", "--r1-file", f"{d}/out/g_r1_it_{last}.bin",
                      "--gam1", repr(float(g["se_gam1"]))])
    assert np.allclose(np.fromfile(f"{d}/out/g_it_{last}_pval_se.bin"), g["pval_se"], rtol=1e-10, atol=1e-300)
    run_cli(common + ["--run-mode", "association_test", "--pval-method", "loo", "--estimate-file", f"{d}/out/g_it_{last}.bin"])
    assert np.allclose(np.fromfile(f"{d}/out/g_it_{last}_pval_loo.bin"), g["pval_loo"], rtol=1e-7, atol=1e-300)
    Nt = int(g["N_test"])
    sim.write_dataset(d, "tst", Nt, int(g["M"]), float(g["lam"]), float(g["h2"]), int(g["data_seed"]) + 1000)
    run_cli(["--meth-file-test", f"{d}/tst.bin", "--phen-file-test", f"{d}/tst.phen", "--N-test", Nt, "--Mt", g["M"], "--out-dir",
             f"{d}/out", "--out-name", "g", "--run-mode", "test", "--estimate-file", f"{d}/out/g_it_1.bin", "--test-iter-range",
             f"1,{last}"])
    got = open(f"{d}/out/g_test.csv", "rb").read()
    want = bytes(g["csv_test"])
    assert len(got) == len(want) and got[:39] == want[:39]
    assert_rows_close(csv_rows(got), csv_rows(want), 1e-8, "test.csv")
    # four saved estimates share one pass over the test matrix (ranges that are not a multiple of four, a missing file:
    # like the reference, an unreadable estimate is a zero vector, src/utilities.cpp:251-267, i.e. R2 = 1 - |y|^2/(sd^2 n))
    os.remove(f"{d}/out/g_test.csv")
    run_cli(["--meth-file-test", f"{d}/tst.bin", "--phen-file-test", f"{d}/tst.phen", "--N-test", Nt, "--Mt", g["M"], "--out-dir",
             f"{d}/out", "--out-name", "g", "--run-mode", "test", "--estimate-file", f"{d}/out/g_it_1.bin", "--test-iter-range",
             f"1,{last + 2}"])
    rows = csv_rows(open(f"{d}/out/g_test.csv", "rb").read())
    assert sorted(rows) == list(range(1, last + 3))
    assert_rows_close({k: rows[k] for k in range(1, last + 1)}, csv_rows(want), 1e-8, "test.csv, longer range")
    assert rows[last + 1][0] == rows[last + 2][0] and rows[last + 1][0] < 0.01
    # README's literal association command (test-file flags) fails like the reference: FATAL, exit 1
    res = subprocess.run([build.MAIN_METH, "--meth-file-test", f"{d}/ex.bin", "--phen-file-test", f"{d}/ex.phen", "--N", "300", "--Mt", "800",
                          "--run-mode", "association_test"], stdout=subprocess.PIPE, text=True)
    assert res.returncode == 1 and "FATAL: could not open phenotype file" in res.stdout


def test_full_size_properties():
    """One 8-GPU shard of the headline configuration (N = 20 000, M = 106 250; 17 GB), device-generated: properties that
    do not need a CPU pass over the matrix — adjointness <A x, p> = <x, A^T p>, linearity, statistics of the synthetic
    block, kernel variants agreeing, and the CG solution satisfying its own residual test through independent calls."""
    N, M = 20000, 106250
    sh = capi.Shard(N, M)
    sh.generate_iid(7)
    sh.compute_stats()
    mave, msig = sh.stats()
    assert abs(mave.mean()) < 1e-3 and abs(msig.mean() - 1) < 1e-2
    rng = np.random.default_rng(0)
    x, x2, p = rng.standard_normal(M), rng.standard_normal(M), rng.standard_normal(N)
    Ax, ATp = sh.Ax(x), sh.ATx(p)
    assert abs(Ax @ p - x @ ATp) <= 1e-11 * math.sqrt((Ax @ Ax) * (p @ p))
    assert rel_l2(sh.Ax(0.5 * x - 2 * x2), 0.5 * Ax - 2 * sh.Ax(x2)) < 1e-12
    for knobs in (dict(ax_rv=1, ax_unroll=8), dict(ax_rv=4, ax_unroll=4), dict(atx_impl=0, atx_cols=1, atx_unroll=8), dict(atx_impl=0, atx_cols=4, atx_unroll=4),
                  dict(atx_impl=2, atx_cols=2, atx_unroll=4), dict(ax_impl=1, atx_impl=1)):
        for k, v in knobs.items():
            sh.set_tuning(k, v)
        assert rel_l2(sh.Ax(x), Ax) < 1e-13 and rel_l2(sh.ATx(p), ATp) < 1e-13
    # CG: ||(tau A^T A + gam2) mu - v|| / ||v|| below the tolerance, evaluated with separate operator calls
    tau, gam2 = 1.7, 0.9
    v = rng.standard_normal(M)
    sh.set(V_V, v)
    it, rel, _ = sh.cg_solve(V_V, V_X2, tau, gam2, tol=1e-8, max_iter=200)
    mu = sh.get(V_X2)
    res = v - (tau * sh.ATx(sh.Ax(mu)) + gam2 * mu)
    assert 1 < it < 200 and np.linalg.norm(res) / np.linalg.norm(v) < 2e-8
    assert abs(np.linalg.norm(res) / np.linalg.norm(v) - rel) < 1e-9
    sh.close()


@pytest.mark.parametrize("name", ["linear_wellcond", "probit_small"])
def test_f32_storage_vamp_matches_oracle_on_rounded_matrix(name, tmp_path):
    """--storage f32 (opt-in, outside the reference's contract): the whole VAMP run equals the oracle's run on the matrix
    rounded to FP32, to the same 1e-9 — the mode changes the data that is held, not the arithmetic."""
    g = load_golden(name)
    A, y_txt, beta = golden_inputs(g)
    A32 = A.astype(np.float32).astype(np.float64)
    v = oracle_run(g, A32, y_txt, beta)
    model = g["model"]
    y = standardize_phen(y_txt) if model == "linear" else y_txt
    sh = capi.Shard(int(g["N"]), int(g["M"]), storage="f32")
    sh.upload(A)
    sh.compute_stats()
    kw = dict(gamw=2.0, seed=int(g["probe_seed"]))
    kw.update(extra_kwargs(g))
    if "h2" in kw:
        kw["gamw"] = 1.0 / (1.0 - kw.pop("h2"))              # src/main_meth.cpp:52
    sol = capi.Solver(sh, y, model=model, true_signal=beta, **kw)
    for k in range(1, int(g["iterations"]) + 1):
        r = sol.step()
        assert rel_l2(r["x1"], v.dump[k][0]) < 1e-9 and rel_l2(r["r1"], v.dump[k][1]) < 1e-9
    sol.close()
    sh.close()
    # and through the command line
    d = str(tmp_path)
    golden_inputs(g, d)
    os.makedirs(tmp_path / "out")
    run_cli(["--meth-file", f"{d}/ex.bin", "--phen-file", f"{d}/ex.phen", "--N", g["N"], "--Mt", g["M"], "--out-dir", f"{d}/out", "--out-name", "g",
             "--iterations", 3, "--true-signal-file", f"{d}/ex_ts.bin", "--model", model, "--stop-criteria-thr", "0", "--seed", g["probe_seed"],
             "--storage", "f32"] + list(g["extra"]))
    for k in range(1, 4):
        assert rel_l2(np.fromfile(f"{d}/out/g_it_{k}.bin"), v.dump[k][0]) < 1e-9
