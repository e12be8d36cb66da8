"""SURVEY.md §8 b2 from the reference's side of the seam: oracle/_ref/main_meth_ref_gpudata is the reference's OWN main_meth.cpp,
vamp.cpp (+ vamp_probit.cpp), utilities.cpp and options.cpp (oracle patches P1-P3 only) linked against the `class data` adapter of
INTEGRATION.md §2 (oracle/ref_shims/data_gpu.cpp) and libvampomi_cuda.so instead of src/data.cpp. Its VAMP loop calls data::Ax /
data::ATx / data::pvals_loo (src/data.hpp:47-90; call sites src/vamp.cpp:232,303,508,518,519,653,654) — and gets our CUDA
operators. Its output files must equal the fixtures the unmodified reference produced."""
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import assert_rows_close, csv_rows, golden_inputs, load_golden, rel_l2

pytestmark = pytest.mark.gpu
ADAPTER_BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "main_meth_ref_gpudata")


def run(args, seed):
    res = subprocess.run([ADAPTER_BIN] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                         env=dict(os.environ, VAMPOMI_SEED=str(seed), OMP_NUM_THREADS="4"))
    assert res.returncode == 0, res.stdout[-3000:]
    return res.stdout


@pytest.mark.parametrize("name", ["linear_wellcond", "probit_small", "linear_ragged"])
def test_reference_vamp_loop_over_our_class_data(name, tmp_path):
    if not os.path.isfile(ADAPTER_BIN):
        pytest.skip("oracle/_ref/main_meth_ref_gpudata is not present (built only where /root/reference exists)")
    g = load_golden(name)
    d = str(tmp_path)
    golden_inputs(g, d)
    os.makedirs(tmp_path / "out")
    its = int(g["iterations"])
    log = run(["--meth-file", f"{d}/ex.bin", "--phen-file", f"{d}/ex.phen", "--N", g["N"], "--Mt", g["M"], "--out-dir", f"{d}/out", "--out-name", "g",
               "--iterations", its, "--true-signal-file", f"{d}/ex_ts.bin", "--model", g["model"], "--stop-criteria-thr", "0", "--verbosity", "1"]
              + list(g["extra"]), g["probe_seed"])
    # this IS the reference's loop (compiled -O2): default-gam1 fixtures are compared with the -O2 twin of the fixture
    x1_want = g["x1_O2"] if "x1_O2" in g else g["x1"]
    r1_want = g["r1_O2"] if "r1_O2" in g else g["r1"]
    for k in range(1, its + 1):
        assert rel_l2(np.fromfile(f"{d}/out/g_it_{k}.bin"), x1_want[k - 1]) < 1e-9, f"x1_hat it {k}"
        assert rel_l2(np.fromfile(f"{d}/out/g_r1_it_{k}.bin"), r1_want[k - 1]) < 1e-9, f"r1 it {k}"
    want = csv_rows(g["csv_params_O2"] if "csv_params_O2" in g else g["csv_params"])
    assert_rows_close(csv_rows(open(f"{d}/out/g_params.csv", "rb").read()), want, 1e-8, "params")
    # CG iteration counts from the reference's own log lines
    counts = []
    for block in log.split("iteration = ")[1:]:
        lm, _, ons = block.partition("[CG onsager]")
        k1, n = len(re.findall(r"\[CG\] it = ", lm)), len(re.findall(r"\[CG\] it = ", ons))
        last = re.findall(r"\|\|r_it\|\| / \|\|RHS\|\| = ([0-9.e+-]+)", ons)
        counts.append((k1, n if (last and float(last[-1]) < 1e-5) else n + 1))
    assert counts == [tuple(c) for c in g["cg_iters"]]


def test_reference_association_loo_over_our_class_data(tmp_path):
    """association_test / loo through the reference's own main (src/main_meth.cpp:245-264): data::Ax and data::pvals_loo are ours."""
    if not os.path.isfile(ADAPTER_BIN):
        pytest.skip("oracle/_ref/main_meth_ref_gpudata is not present")
    g = load_golden("linear_small")
    d = str(tmp_path)
    golden_inputs(g, d)
    os.makedirs(tmp_path / "out")
    last = int(g["iterations"])
    g["x1"][last - 1].tofile(f"{d}/out/g_it_{last}.bin")
    run(["--meth-file", f"{d}/ex.bin", "--phen-file", f"{d}/ex.phen", "--N", g["N"], "--Mt", g["M"], "--out-dir", f"{d}/out", "--out-name", "g",
         "--run-mode", "association_test", "--pval-method", "loo", "--estimate-file", f"{d}/out/g_it_{last}.bin"], g["probe_seed"])
    assert np.allclose(np.fromfile(f"{d}/out/g_it_{last}_pval_loo.bin"), g["pval_loo"], rtol=1e-7, atol=1e-300)
