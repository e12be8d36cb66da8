"""the reference's own post-processing scripts must work unchanged on OUR output files. The files under
tests/golden/ours_linear_small/ were written by bin/main_meth on a B200 (tests/tools/make_ours_outputs.py); the scripts live
in /root/reference/scripts, which exists in the build container only — so this test runs here, on committed outputs.
  scripts/p_vals.py:41,58-62   strips the NUL holes of _params.csv, takes gam1 of the target iteration, re-derives the se p-values
  scripts/metrics.py:40,57,76,91   csv.reader over NUL-stripped rows; the columns it reads pin our column order
  scripts/roc.py:41-58             reads M doubles of p-values and of the true signal, ROC / AUC / FDR / TPR with scikit-learn
metrics.py and roc.py import matplotlib, which this image does not have: they run unchanged against tests/stubs/matplotlib, a
test double that draws nothing and records every plotted series."""
import csv
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import csv_rows, golden_inputs, load_golden

OURS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ours_linear_small")
SCRIPTS = "/root/reference/scripts"
needs_outputs = pytest.mark.skipif(not os.path.isfile(os.path.join(OURS, "g_params.csv")), reason="tests/golden/ours_linear_small not generated")


@needs_outputs
@pytest.mark.skipif(not os.path.isfile(os.path.join(SCRIPTS, "p_vals.py")), reason="the reference's scripts are not on this machine")
def test_reference_p_vals_script_on_our_files(tmp_path):
    g = load_golden("linear_small")
    its, M, N = int(g["iterations"]), int(g["M"]), int(g["N"])
    for f in os.listdir(OURS):
        os.symlink(os.path.join(OURS, f), tmp_path / f)
    res = subprocess.run([sys.executable, os.path.join(SCRIPTS, "p_vals.py"), "--out-name", "script_pvals", "--csv-params", str(tmp_path / "g_params.csv"),
                          "--r1-file", str(tmp_path / f"g_r1_it_{its}.bin"), "--it", str(its), "--M", str(M), "--N", str(N)],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout[-2000:]
    script = np.fromfile(tmp_path / "script_pvals.bin")
    ours = np.fromfile(os.path.join(OURS, f"g_it_{its}_pval_se.bin"))
    assert script.shape == ours.shape == (M,)
    # the script's scipy evaluation of our r1 / our gam1 against our own se run-mode output, and against the reference binary's
    assert np.allclose(script, ours, rtol=1e-9, atol=1e-300)
    assert np.allclose(ours, g["pval_se"], rtol=1e-6, atol=1e-300)       # gam1 / r1 of a default-gam1 run: its parity floor applies


@needs_outputs
def test_metrics_script_readers_parse_our_csvs():
    """The reader logic of scripts/metrics.py (:38-95), restated because the script itself imports matplotlib (absent here):
    csv.reader over NUL-stripped rows, header skipped, the very columns it converts with float()."""
    g = load_golden("linear_small")
    its = int(g["iterations"])

    def rows(name):
        with open(os.path.join(OURS, name), newline="", encoding="utf-8") as f:
            rd = csv.reader((row.replace("\0", "") for row in f), delimiter=",")
            next(rd, None)
            return [r for r in rd]

    test = rows("g_test.csv")
    assert [int(r[0]) for r in test] == list(range(1, its + 1))
    r2_test, corr2_test = [float(r[1]) for r in test], [float(r[2]) for r in test]                   # metrics.py:44-49
    want = csv_rows(g["csv_test"])
    assert np.allclose(r2_test, [want[k][0] for k in range(1, its + 1)], rtol=1e-6) and np.allclose(corr2_test, [want[k][1] for k in range(1, its + 1)], rtol=1e-6, equal_nan=True)
    met = rows("g_metrics.csv")
    r2_den, corr_train, r2_lmmse = [float(r[1]) for r in met], [float(r[2]) for r in met], [float(r[3]) for r in met]   # :61-66
    wm = csv_rows(g["csv_metrics"])
    assert len(met) == its and np.allclose(r2_den, [wm[k][0] for k in range(1, its + 1)], rtol=1e-6)
    assert np.isnan(corr_train[0]) and np.allclose(corr_train[1:], [wm[k][1] for k in range(2, its + 1)], rtol=1e-6)
    assert np.allclose(r2_lmmse, [wm[k][2] for k in range(1, its + 1)], rtol=1e-6)
    par = rows("g_params.csv")
    gam1, gamw = [float(r[2]) for r in par], [float(r[5]) for r in par]                               # :80-87
    wp = csv_rows(g["csv_params"])
    assert np.allclose(gam1, [wp[k][1] for k in range(1, its + 1)], rtol=1e-6) and np.allclose(gamw, [wp[k][4] for k in range(1, its + 1)], rtol=1e-6)
    assert rows("g_prior.csv") == []                 # linear model: header only (src/vamp.cpp:392 is commented out) -> metrics.py's lam stays empty
    # byte-for-byte the reference's own files where nothing numerical is involved
    assert open(os.path.join(OURS, "g_prior.csv"), "rb").read() == bytes(g["csv_prior"])
    for name, key in (("g_params.csv", "csv_params"), ("g_metrics.csv", "csv_metrics"), ("g_test.csv", "csv_test")):
        got, ref = open(os.path.join(OURS, name), "rb").read(), bytes(g[key])
        assert len(got) == len(ref) and np.array_equal(np.frombuffer(got, dtype=np.uint8) == 0, np.frombuffer(ref, dtype=np.uint8) == 0), name


def _run_script(script, args, tmp_path):
    """Runs an unmodified reference script with the matplotlib test double first on the path; returns (stdout, plotted series)."""
    log = tmp_path / "plots.jsonl"
    env = dict(os.environ, MPL_STUB_LOG=str(log), PYTHONPATH=os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs"))
    res = subprocess.run([sys.executable, os.path.join(SCRIPTS, script)] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    assert res.returncode == 0, res.stdout[-2000:]
    return res.stdout, [json.loads(line) for line in open(log)]


@needs_outputs
@pytest.mark.skipif(not os.path.isfile(os.path.join(SCRIPTS, "metrics.py")), reason="the reference's scripts are not on this machine")
def test_reference_metrics_script_on_our_files(tmp_path):
    """scripts/metrics.py, unchanged, on OUR four CSV files: it must parse every row (NUL holes, column order, `-nan` cells) and
    the series it would draw must be the reference binary's own values."""
    g = load_golden("linear_small")
    its = int(g["iterations"])
    for f in os.listdir(OURS):
        os.symlink(os.path.join(OURS, f), tmp_path / f)
    out, plots = _run_script("metrics.py", ["--csv-metrics", str(tmp_path / "g_metrics.csv"), "--csv-test", str(tmp_path / "g_test.csv"),
                                            "--csv-params", str(tmp_path / "g_params.csv"), "--csv-prior", str(tmp_path / "g_prior.csv"),
                                            "--iterations", str(its)], tmp_path)
    series = {p["label"]: p["series"] for p in plots if p["kind"] == "plot"}
    assert set(series) == {"Denoising", "LMMSE", "Test", "gamw", "gam1"}
    wm, wt, wp = csv_rows(g["csv_metrics"]), csv_rows(g["csv_test"]), csv_rows(g["csv_params"])
    want = {"Denoising": [wm[k][0] for k in range(1, its + 1)], "LMMSE": [wm[k][2] for k in range(1, its + 1)], "Test": [wt[k][0] for k in range(1, its + 1)],
            "gamw": [wp[k][4] for k in range(1, its + 1)], "gam1": [wp[k][1] for k in range(1, its + 1)]}
    for label, (x, y) in series.items():
        assert x == list(range(1, its + 1)) and np.allclose(y, want[label], rtol=1e-6), label     # a default-gam1 run: its parity floor applies
    assert any(p["kind"] == "savefig" and p["label"].endswith("g_metrics.png") for p in plots)


@needs_outputs
@pytest.mark.skipif(not os.path.isfile(os.path.join(SCRIPTS, "roc.py")), reason="the reference's scripts are not on this machine")
def test_reference_roc_script_on_our_pvalues(tmp_path):
    """scripts/roc.py, unchanged, on the p-value file our `association_test --pval-method se` run wrote: the table it prints (markers
    under the Bonferroni threshold, AUC, FDR, TPR) must equal the one it prints for the reference binary's p-values."""
    g = load_golden("linear_small")
    its, M = int(g["iterations"]), int(g["M"])
    _, _, beta = golden_inputs(g)                            # the effects the fixture's phenotype was simulated with
    beta.tofile(tmp_path / "beta.bin")
    tables = []
    for name, pv in (("ours", np.fromfile(os.path.join(OURS, f"g_it_{its}_pval_se.bin"))), ("reference", np.asarray(g["pval_se"], dtype=np.float64))):
        pv.tofile(tmp_path / f"{name}_pvals.bin")
        out, _ = _run_script("roc.py", ["--pval", str(tmp_path / f"{name}_pvals.bin"), "--true-signal", str(tmp_path / "beta.bin"),
                                        "--out-name", name, "--it", str(its), "--M", str(M)], tmp_path)
        row = [line for line in out.splitlines() if line.startswith("|") and "AUC" not in line]
        assert len(row) == 1, out[-1000:]
        tables.append([float(c) for c in row[0].strip("| ").split("|")])
    assert tables[0][0] == its and tables[0] == tables[1], tables
