"""World-size-2 run of the marker-sharded VAMP loop on CPU (gloo): each rank owns the column block divide_work gives
it, N-vectors are replicated, A x and the scalar sums go through an all-reduce — the communication pattern the GPU
build runs over NCCL. Shard results must agree with the single-shard run and land at byte offset S*8 of shared files."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir, name):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import golden_inputs, load_golden, oracle_run
    from oracle import vamp_oracle as vo
    from vampomi_b200 import capi
    g = load_golden(name)
    g["iterations"] = 4
    A, y_txt, beta = golden_inputs(g)
    M, S = capi.divide_work(int(g["M"]), world, rank)            # the product's own work split (host-side C ABI)
    oracle_run(g, A[S:S + M], y_txt, beta[S:S + M], out_dir=out_dir, comm=vo.TorchComm(), S=S, Mt=int(g["M"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("name", ["linear_wellcond", "probit_small"])
def test_two_shards_match_one_shard(name, tmp_path, lib):
    from helpers import REL_CSV, REL_VEC, assert_rows_close, csv_rows, load_golden, rel_l2
    port = 29500 + (2 * os.getpid() + (name != "linear_wellcond")) % 4000      # a port of its own per case: no wait for the last one to free
    mp.spawn(_worker, args=(2, port, str(tmp_path), name), nprocs=2, join=True)
    g = load_golden(name)
    for k in range(1, 5):
        x1 = np.fromfile(tmp_path / f"o_it_{k}.bin")
        r1 = np.fromfile(tmp_path / f"o_r1_it_{k}.bin")
        assert x1.size == int(g["M"])                              # both shards wrote their S*8 slice of ONE file
        assert rel_l2(x1, g["x1"][k - 1]) < REL_VEC and rel_l2(r1, g["r1"][k - 1]) < REL_VEC
    got = csv_rows(open(tmp_path / "o_params.csv", "rb").read())
    want = {k: v for k, v in csv_rows(g["csv_params"]).items() if k <= 4}
    assert_rows_close(got, want, REL_CSV, "params")
