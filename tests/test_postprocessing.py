"""the reference's own post-processing scripts must work unchanged on OUR output files. The files under
tests/golden/ours_linear_small/ were written by bin/main_meth on a B200 (tests/tools/make_ours_outputs.py); the scripts live
in /root/reference/scripts, which exists in the build container only — so this test runs here, on committed outputs.
  scripts/p_vals.py:41,58-62   strips the NUL holes of _params.csv, takes gam1 of the target iteration, re-derives the se p-values
  scripts/metrics.py:40,57,76,91   csv.reader over NUL-stripped rows; the columns it reads pin our column order"""
import csv
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import csv_rows, load_golden

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
