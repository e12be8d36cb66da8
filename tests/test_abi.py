"""CPU-side checks of the drop-in boundary: the shared library loads, exports exactly the symbols include/*.h
declare, the host-only helpers agree with the oracle, and compute entry points fail loudly without a GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import vamp_oracle as vo
from vampomi_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("vampomi.h", "vampomi_host.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(vampomi_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_exactly_the_declared_abi(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", build.LIB_PATH], stdout=subprocess.PIPE, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line and line.split()[-1].startswith("vampomi_")}
    declared = declared_symbols()
    assert declared == exported, f"header-only: {declared - exported}; undeclared exports: {exported - declared}"
    assert declared == set(capi.exported_symbols()), "ctypes table out of sync with include/*.h"
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.vampomi_abi_version() == 1


@pytest.mark.parametrize("Mt,nranks", [(2000, 1), (2000, 3), (850000, 8), (17, 5), (8, 8), (100001, 7)])
def test_divide_work_matches_reference_rule(lib, Mt, nranks):
    total = 0
    for r in range(nranks):
        M, S = capi.divide_work(Mt, nranks, r)
        assert (M, S) == vo.divide_work(Mt, nranks, r)
        assert S == total
        total += M
    assert total == Mt


def test_no_silent_cpu_fallback(lib):
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.VampomiError, match="no CUDA device"):
        capi.Shard(64, 32)
    # the command line fails the same way: FATAL line, exit status 1 — never a CPU run
    res = subprocess.run([build.MAIN_METH, "--meth-file", "/nonexistent.bin", "--phen-file", __file__, "--N", "2", "--Mt", "2"],
                         stdout=subprocess.PIPE, text=True)
    assert res.returncode == 1 and "FATAL" in res.stdout


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "main_meth_ref")


@pytest.mark.parametrize("argv", [["--bogus", "1"], ["--meth-file"], ["--iterations", "0", "--meth-file", "x"],
                                  ["--N", "-3", "--meth-file", "x"], ["--learn-vars", "-1"], ["--out-dir", "x"]])
def test_cli_option_errors_match_reference(lib, argv):
    got = subprocess.run([build.MAIN_METH] + argv, stdout=subprocess.PIPE, text=True)
    assert got.returncode == 1
    assert "FATAL" in got.stdout
    if os.path.isfile(REF_BIN):       # present in the build container (and shipped to the GPU box as a binary)
        want = subprocess.run([REF_BIN] + argv, stdout=subprocess.PIPE, text=True)
        assert want.returncode == got.returncode
        fatal = [l for l in want.stdout.splitlines() if "FATAL" in l]
        assert fatal and fatal[0] in got.stdout


def test_cli_accepts_every_reference_flag(lib):
    flags = ["--meth-file", "--meth-file-test", "--phen-file", "--phen-file-test", "--cov-file", "--cov-file-test",
             "--estimate-file", "--r1-file", "--cov-estimate-file", "--true-signal-file", "--run-mode", "--model",
             "--pval-method", "--N", "--N-test", "--Mt", "--Mt-test", "--C", "--out-dir", "--out-name", "--iterations",
             "--test-iter-range", "--rho", "--gam1", "--h2", "--alpha-scale", "--probit-var", "--num-mix-comp", "--probs",
             "--vars", "--learn-vars", "--learn-prior-delay", "--CG-max-iter", "--CG-err-tol", "--EM-max-iter",
             "--EM-err-thr", "--stop-criteria-thr", "--merge-vars-thr", "--verbosity"]
    argv = []
    for f in flags:
        argv += [f, {"--run-mode": "association_test", "--pval-method": "none", "--test-iter-range": "1,2", "--probs": "0.5,0.5",
                     "--vars": "0,1", "--phen-file": __file__}.get(f, "1")]
    res = subprocess.run([build.MAIN_METH] + argv, stdout=subprocess.PIPE, text=True)
    assert "unknown" not in res.stdout and "missing argument" not in res.stdout
    assert "--probit-var1" in res.stdout        # the reference echoes this flag without a space (src/options.cpp:199)


def test_cli_additions_are_validated_before_any_gpu_work(lib):
    """The flags the reference does not have (--gpus, --storage, --schedule, --seed) are parsed and echoed like its own;
    bad values end with the reference's FATAL wording and exit status 1 without touching a GPU."""
    for argv in (["--schedule", "sideways"], ["--storage", "f16"], ["--gpus", "0"], ["--schedule"]):
        res = subprocess.run([build.MAIN_METH] + argv, stdout=subprocess.PIPE, text=True)
        assert res.returncode == 1 and "FATAL" in res.stdout, argv
    res = subprocess.run([build.MAIN_METH, "--meth-file", "x", "--phen-file", "y", "--out-dir", "/tmp", "--out-name", "t", "--schedule",
                          "fused", "--storage", "f32", "--seed", "7", "--gpus", "2"], stdout=subprocess.PIPE, text=True)
    assert res.returncode == 1            # --N / --Mt missing: stops after the echo, before any GPU work
    for echo in ("--schedule fused", "--storage f32", "--seed 7", "--gpus 2"):
        assert echo in res.stdout


def test_argument_validation_needs_no_gpu(lib):
    """Bad arguments are rejected with VAMPOMI_ERR_ARG (1) and a message before any CUDA call."""
    import ctypes as C
    h = C.c_void_p()
    for args in ((0, 1, 10, 1, 0, 0), (0, 100, 0, 1, 0, 0), (0, 100, 10, 0, 0, 0), (0, 100, 10, 2, 2, 0), (0, 100, 3, 4, 0, 0),
                 (0, 100, 10, 1, 0, 7)):
        assert lib.vampomi_create_ex(*args, C.byref(h)) == 1
        assert lib.vampomi_last_error()
    assert lib.vampomi_create_ex(0, 100, 10, 1, 0, 0, None) == 1
    M, S = C.c_longlong(), C.c_longlong()
    assert lib.vampomi_divide_work(10, 0, 0, C.byref(M), C.byref(S)) == 1
    assert lib.vampomi_divide_work(10, 2, 2, C.byref(M), C.byref(S)) == 1
    assert lib.vampomi_shard(None, C.byref(M), C.byref(S)) == 1
    assert lib.vampomi_set_tuning(None, b"ax_rv", 1) == 1
    assert lib.vampomi_comm_get_unique_id(None) == 1
    assert lib.vampomi_solver_create(None, None, None, None, None, None) == 1
    buf = C.create_string_buffer(8)
    assert lib.vampomi_host_csv_row(1, None, 0, buf, 8) == 6 and buf.value == b"    1\n"      # empty row: "%5d" + newline
    assert lib.vampomi_host_csv_row(1, None, 0, buf, 3) == -1                               # does not fit
