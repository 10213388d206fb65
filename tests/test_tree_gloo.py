"""The tree-parallel solve (fers_core_b200.chol: fers_b200_chol_plan_partition + tree_schedule, executed on the GPU by
TreeParallelSolver) on the CPU: every rank of a gloo group runs the SAME schedule on its own share of the fronts with the
numpy emulator of the numeric phase (tests/chol_emulator.py) and real send / recv / all_gather between the processes.  Proves
ownership, the order of levels and messages, and what each message must carry -- without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fers_core_b200 import _lib, chol, synth
from oracle import fers_oracle
from tests import chol_emulator


def _problem():
    model = synth.grid_frame(9, 8, 6, n_combos=4)
    asm = fers_oracle.Assembly(model)
    K = asm.K_elastic().tobsr(blocksize=(6, 6))
    K.sort_indices()
    rowptr, colidx = chol.host_pattern(model.n_nodes, model.m_n1, model.m_n2)
    assert np.array_equal(rowptr, K.indptr) and np.array_equal(colidx, K.indices)
    fixed = (model.bc_code.reshape(-1) == 1).astype(np.uint8)
    rng = np.random.default_rng(3)
    b = rng.normal(size=(model.n_dof, 3))
    return model, rowptr, colidx, np.ascontiguousarray(K.data), fixed, b


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model, rowptr, colidx, vals, fixed, b = _problem()
        plan = chol.CholPlan(rowptr, colidx, model.xyz)
        n_out = 4 + 4 * world
        info = (C.c_longlong * n_out)()
        _lib.check(plan.lib.fers_b200_chol_plan_partition(plan.handle, rank, world, info, n_out), "partition")
        k, ls = int(info[0]), int(info[1])
        n_top = (2 << k) - 1
        ids = [int(info[4 + t]) for t in range(n_top)]
        top = [[ids[(1 << d) - 1 + j] for j in range(1 << d)] for d in range(k + 1)]
        ex = plan.export()                                   # this rank's level lists and K blocks
        em = chol_emulator.Emulator(ex, rowptr, colidx, vals.reshape(len(vals), 36), fixed)
        lp = ex["level_ptr"]

        def mine(lo, hi):                                    # the fronts this rank keeps on levels [lo, hi)
            return [int(f) for f in ex["level_fronts"][lp[lo]:lp[min(hi, len(lp) - 1)]]]
        kept = mine(0, ls + 1)
        sched = chol.tree_schedule(k, ls, top)
        nd = lambda fid: em.dofs[fid]                        # DOFs of a factorised front: own nodes, then boundary
        fr = ex["fronts"]

        def front_dofs(fid):                                 # (own, boundary) DOFs of ANY front, from the plan
            f = fr[fid]
            nodes = ex["nodes"][f["nodes_off"]:f["nodes_off"] + f["n1"] + f["n2"]]
            d = (6 * nodes[:, None] + np.arange(6)).reshape(-1)
            return d[:6 * f["n1"]], d[6 * f["n1"]:]
        for step in sched["factor"]:
            if step[0] == "levels":
                em.factor(mine(step[1], step[2]))
            else:
                _, fid, _depth, src, dst = step
                m2 = 6 * int(fr[fid]["n2"])
                if rank == src:
                    dist.send(torch.from_numpy(np.ascontiguousarray(em.schur[fid])), dst)
                elif rank == dst:
                    t = torch.empty((m2, m2), dtype=torch.float64)
                    dist.recv(t, src)
                    em.schur[fid] = t.numpy()
        x0 = em.start(b)
        x = x0.copy()
        for step in sched["solve"]:
            kind = step[0]
            if kind == "fwd":
                em.forward(x, mine(step[1], step[2]))
            elif kind == "bwd":
                em.backward(x, mine(step[1], step[2]))
            elif kind == "w":                                # what the subtree of `fid` took off its boundary rows
                _, fid, _depth, src, dst = step
                bnd = front_dofs(fid)[1]
                if rank == src:
                    dist.send(torch.from_numpy(np.ascontiguousarray(x[bnd] - x0[bnd])), dst)
                elif rank == dst:
                    t = torch.empty((len(bnd), b.shape[1]), dtype=torch.float64)
                    dist.recv(t, src)
                    x[bnd] += t.numpy()
            elif kind == "anc":                              # solution rows of an upper front and its ancestors
                _, chain, src, dst = step
                own = np.concatenate([front_dofs(f)[0] for f in chain])
                if rank == src:
                    dist.send(torch.from_numpy(np.ascontiguousarray(x[own])), dst)
                elif rank == dst:
                    t = torch.empty((len(own), b.shape[1]), dtype=torch.float64)
                    dist.recv(t, src)
                    x[own] = t.numpy()
            elif kind == "gather":
                own = np.concatenate([front_dofs(f)[0] for f in kept]) if kept else np.zeros(0, np.int64)
                sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
                dist.all_gather(sizes, torch.tensor([len(own)]))
                width = int(max(s.item() for s in sizes))
                pad_i = torch.full((width,), -1, dtype=torch.int64); pad_i[:len(own)] = torch.from_numpy(own)
                pad_v = torch.zeros((width, b.shape[1]), dtype=torch.float64); pad_v[:len(own)] = torch.from_numpy(x[own])
                all_i = [torch.empty_like(pad_i) for _ in range(world)]
                all_v = [torch.empty_like(pad_v) for _ in range(world)]
                dist.all_gather(all_i, pad_i); dist.all_gather(all_v, pad_v)
                for r in range(world):
                    if r != rank:
                        sel = all_i[r] >= 0
                        x[all_i[r][sel].numpy()] = all_v[r][sel].numpy()
        out[rank] = (x, sorted(kept))
        plan.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_tree_schedule_solves_the_system_over_gloo(world):
    model, rowptr, colidx, vals, fixed, b = _problem()
    a = chol_emulator.expand_masked(rowptr, colidx, vals.reshape(len(vals), 36), fixed)
    want = np.linalg.solve(a, np.where(fixed.astype(bool)[:, None], 0.0, b))
    whole = chol.CholPlan(rowptr, colidx, model.xyz)
    n_fronts = whole.info.n_fronts
    whole.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29570 + world, out), nprocs=world, join=True)
    assert set(out.keys()) == set(range(world))
    seen = []
    for r in range(world):
        x, kept = out[r]
        assert np.abs(x - want).max() <= 1e-9 * np.abs(want).max(), r          # every rank ends with the whole solution
        seen += kept
    assert sorted(seen) == list(range(n_fronts))                                # each front factorised exactly once
