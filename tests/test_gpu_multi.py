"""Marker sharding over several GPUs of one node (NCCL all-reduce of the N-length A x partial sums and of the packed
scalar sums): results must not depend on the number of shards. Skipped on single-GPU boxes."""
import os
import subprocess

import numpy as np
import pytest

from vampomi_b200 import build, capi
from helpers import assert_rows_close, csv_rows, golden_inputs, load_golden, rel_l2, tolerances

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,gpus", [("linear_wellcond", 2), ("probit_small", 2), ("linear_ragged", 3), ("linear_wellcond", 4)])
def test_main_meth_is_shard_invariant(name, gpus, tmp_path):
    if capi.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    g = load_golden(name)
    rel_vec, rel_csv = tolerances(g)
    d = str(tmp_path)
    golden_inputs(g, d)
    os.makedirs(tmp_path / "out")
    its = int(g["iterations"])
    args = ["--meth-file", f"{d}/ex.bin", "--phen-file", f"{d}/ex.phen", "--N", g["N"], "--Mt", g["M"], "--out-dir", f"{d}/out",
            "--out-name", "g", "--iterations", its, "--true-signal-file", f"{d}/ex_ts.bin", "--model", g["model"],
            "--stop-criteria-thr", "0", "--seed", g["probe_seed"], "--gpus", gpus] + list(g["extra"])
    res = subprocess.run([build.MAIN_METH] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:]
    for k in range(1, its + 1):
        x1 = np.fromfile(f"{d}/out/g_it_{k}.bin")
        assert x1.size == int(g["M"])
        assert rel_l2(x1, g["x1"][k - 1]) < rel_vec
        assert rel_l2(np.fromfile(f"{d}/out/g_r1_it_{k}.bin"), g["r1"][k - 1]) < rel_vec
    for kind in ("params", "metrics"):
        assert_rows_close(csv_rows(open(f"{d}/out/g_{kind}.csv", "rb").read()), csv_rows(g[f"csv_{kind}"]), rel_csv, kind)


@pytest.mark.parametrize("gpus", [2, 4, 8])
def test_peer_exchange_equals_nccl(gpus, tmp_path):
    """The fused peer-memory all-reduce (mode 2) and the NCCL path (VAMPOMI_XCHG=0) must give the same iterates: both add
    the same shard contributions, only the order of the G additions may differ."""
    if capi.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    g = load_golden("linear_wellcond")
    d = str(tmp_path)
    golden_inputs(g, d)
    its = 4
    outs = {}
    for mode in ("1", "0"):
        os.makedirs(tmp_path / f"out{mode}")
        args = ["--meth-file", f"{d}/ex.bin", "--phen-file", f"{d}/ex.phen", "--N", g["N"], "--Mt", g["M"], "--out-dir", f"{d}/out{mode}",
                "--out-name", "g", "--iterations", its, "--true-signal-file", f"{d}/ex_ts.bin", "--stop-criteria-thr", "0", "--seed",
                g["probe_seed"], "--gpus", gpus] + list(g["extra"])
        res = subprocess.run([build.MAIN_METH] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                             timeout=600, env=dict(os.environ, VAMPOMI_XCHG=mode))
        assert res.returncode == 0, res.stdout[-3000:]
        assert ("NVLink peer memory" in res.stdout) == (mode == "1"), res.stdout[:2000]
        outs[mode] = res.stdout
    for k in range(1, its + 1):
        a, b = np.fromfile(f"{d}/out1/g_it_{k}.bin"), np.fromfile(f"{d}/out0/g_it_{k}.bin")
        assert rel_l2(a, b) < 1e-12
        assert rel_l2(a, g["x1"][k - 1]) < 1e-9
    cg = lambda s: [l for l in s.splitlines() if l.startswith("[CG] LMMSE solve")]
    assert cg(outs["1"]) == cg(outs["0"])


def test_association_and_test_modes_on_two_gpus(tmp_path):
    """se / loo p-values and the out-of-sample test mode with the markers split over 2 GPUs: every shard writes its S*8
    slice of the same output file; results equal the single-rank reference fixtures."""
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from vampomi_b200 import sim
    g = load_golden("linear_small")
    d = str(tmp_path)
    golden_inputs(g, d)
    os.makedirs(tmp_path / "out")
    last = int(g["iterations"])
    for k in range(1, last + 1):
        g["x1"][k - 1].tofile(f"{d}/out/g_it_{k}.bin")
        g["r1"][k - 1].tofile(f"{d}/out/g_r1_it_{k}.bin")
    common = ["--meth-file", f"{d}/ex.bin", "--phen-file", f"{d}/ex.phen", "--N", g["N"], "--Mt", g["M"], "--out-dir", f"{d}/out", "--out-name", "g",
              "--gpus", 2]

    def run(args):
        res = subprocess.run([build.MAIN_METH] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert res.returncode == 0, res.stdout[-3000:]

    run(common + ["--run-mode", "association_test", "--pval-method", "se", "--r1-file", f"{d}/out/g_r1_it_{last}.bin", "--gam1", repr(float(g["se_gam1"]))])
    assert np.allclose(np.fromfile(f"{d}/out/g_it_{last}_pval_se.bin"), g["pval_se"], rtol=1e-10, atol=1e-300)
    run(common + ["--run-mode", "association_test", "--pval-method", "loo", "--estimate-file", f"{d}/out/g_it_{last}.bin"])
    assert np.allclose(np.fromfile(f"{d}/out/g_it_{last}_pval_loo.bin"), g["pval_loo"], rtol=1e-7, atol=1e-300)
    Nt = int(g["N_test"])
    sim.write_dataset(d, "tst", Nt, int(g["M"]), float(g["lam"]), float(g["h2"]), int(g["data_seed"]) + 1000)
    run(["--meth-file-test", f"{d}/tst.bin", "--phen-file-test", f"{d}/tst.phen", "--N-test", Nt, "--Mt", g["M"], "--out-dir", f"{d}/out", "--out-name", "g",
         "--run-mode", "test", "--estimate-file", f"{d}/out/g_it_1.bin", "--test-iter-range", f"1,{last}", "--gpus", 2])
    assert_rows_close(csv_rows(open(f"{d}/out/g_test.csv", "rb").read()), csv_rows(g["csv_test"]), 1e-8, "test.csv")


@pytest.mark.parametrize("write_gpus,resume_gpus", [(2, 1), (1, 2)])
def test_checkpoint_written_on_one_sharding_resumes_on_another(write_gpus, resume_gpus, tmp_path):
    """--checkpoint-every / --resume-from across GPU counts: the checkpoint holds the marker vectors at their GLOBAL offsets and one
    completion record per writing rank, so a run checkpointed on 2 GPUs continues on 1 (and the reverse) with the reference's iterates."""
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    g = load_golden("linear_wellcond")
    d = str(tmp_path)
    golden_inputs(g, d)
    os.makedirs(tmp_path / "out")
    its = int(g["iterations"])
    base = ["--meth-file", f"{d}/ex.bin", "--phen-file", f"{d}/ex.phen", "--N", g["N"], "--Mt", g["M"], "--out-dir", f"{d}/out", "--out-name", "g",
            "--true-signal-file", f"{d}/ex_ts.bin", "--stop-criteria-thr", 0, "--seed", g["probe_seed"]] + list(g["extra"])

    def run(args):
        res = subprocess.run([build.MAIN_METH] + [str(a) for a in base + args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert res.returncode == 0, res.stdout[-3000:]
        return res.stdout

    run(["--iterations", 3, "--checkpoint-every", 3, "--gpus", write_gpus])
    ck = f"{d}/out/g_checkpoint_it_3.bin"
    recs = np.frombuffer(open(ck, "rb").read()[2048:2048 + 24 * write_gpus], dtype=np.int64).reshape(write_gpus, 3)
    assert np.all(recs[:, 0] == 3) and recs[:, 2].sum() == int(g["M"]) and recs[0, 1] == 0        # (iteration, first marker, markers) per writer
    out = run(["--iterations", its, "--resume-from", ck, "--gpus", resume_gpus])
    assert "resuming after iteration 3" in out
    for k in range(1, its + 1):
        assert rel_l2(np.fromfile(f"{d}/out/g_it_{k}.bin"), g["x1"][k - 1]) < 1e-9
        assert rel_l2(np.fromfile(f"{d}/out/g_r1_it_{k}.bin"), g["r1"][k - 1]) < 1e-9
    for kind in ("params", "metrics"):
        assert_rows_close(csv_rows(open(f"{d}/out/g_{kind}.csv", "rb").read()), csv_rows(g[f"csv_{kind}"]), 1e-8, kind)
