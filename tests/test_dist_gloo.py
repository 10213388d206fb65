"""world_size-2 gloo test of the multi-GPU host logic (sharding + summary gather) on the CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fers_core_b200 import dist as fdist


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 64, 65, 8192):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = fdist.shard_range(n, world, r)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [fdist.shard_range(n, world, r)[1] - fdist.shard_range(n, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    factors = np.arange(11 * 3, dtype=np.float64).reshape(11, 3)
    mine = fdist.shard_factors(factors, world, rank)
    counts = [fdist.shard_range(11, world, r)[1] - fdist.shard_range(11, world, r)[0] for r in range(world)]
    # stand-in for a per-combination result summary computed on this rank
    local = torch.from_numpy(mine.sum(axis=1, keepdims=True).copy())
    allv = fdist.gather_summary(local, counts)
    t = fdist.max_over_ranks(float(rank + 1), torch.device("cpu"))
    # result tensors [rows, 12, columns]: every rank contributes its own columns (with padding columns behind them)
    full = torch.arange(5 * 12 * 11, dtype=torch.float64).reshape(5, 12, 11)
    lo, hi = fdist.shard_range(11, world, rank)
    local = torch.full((5, 12, 8), -1.0, dtype=torch.float64)
    local[..., : hi - lo] = full[..., lo:hi]
    got_cols = fdist.gather_columns(local, counts)

    class FakeEngine:                                  # run_linear of a stand-in engine: U[:, c] = factors[c].sum()
        def run_linear(self, fac, detail=False):
            u = torch.zeros((4, 8), dtype=torch.float64)
            u[:, : fac.shape[0]] = torch.from_numpy(fac.sum(axis=1))[None, :]
            return {"U": u}
    sh = fdist.run_linear_sharded(FakeEngine(), factors, gather=("U", "R"))
    if rank == 0:
        q.put((allv.numpy().reshape(-1).tolist(), t, bool(torch.equal(got_cols, full)), sh["U"][0].tolist(), sh["columns"]))
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, tmax, cols_ok, sharded_u, my_cols = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(11 * 3, dtype=np.float64).reshape(11, 3).sum(axis=1)
    assert np.allclose(got, want)
    assert tmax == 2.0
    assert cols_ok and np.allclose(sharded_u, want) and my_cols == (0, 6)
