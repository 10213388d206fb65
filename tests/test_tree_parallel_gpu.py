"""Tree-parallel sparse direct solve (chol.TreeParallelSolver, Engine(tree_parallel=...)) on two GPUs against the
single-GPU solve of the same model: same displacements to round-off, same residual class.  Needs two GPUs (skipped on the
one-GPU test box; `gpurun --gpus 2 -- python -m pytest tests/test_tree_parallel_gpu.py -m gpu`)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from fers_core_b200 import engine, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        model = synth.grid_frame(14, 12, 8, n_combos=8)
        fac = model.combo_factors[:8]
        tp = engine.Engine(model, dev, solver="cholesky", tree_parallel=(rank, world))
        u_tp = tp.run_linear(fac)["U"].clone()
        u_tp2 = tp.run_linear(fac)["U"]                       # a second pass re-uses the plan and the buffers (the load
        assert float((u_tp - u_tp2).abs().max() / u_tp.abs().max()) < 1e-12      # vectors are summed with atomics: round-off)
        one = engine.Engine(model, dev, solver="cholesky")
        u_1 = one.run_linear(fac)["U"]
        out[rank] = (float((u_tp - u_1).abs().max() / u_1.abs().max()), float(tp.stats.relres), float(one.stats.relres),
                     tp._chol.fronts_kept, tp._chol.link_bytes["factor"])
    finally:
        dist.destroy_process_group()


def test_two_gpus_give_the_single_gpu_solution():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29541, out), nprocs=2, join=True)
    assert set(out.keys()) == {0, 1}
    for rank in (0, 1):
        diff, res_tp, res_1, kept, link = out[rank]
        assert diff < 1e-11 and res_tp < 1e-11 and res_1 < 1e-11, out[rank]
        assert kept > 0
    assert out[0][4] == 0 and out[1][4] > 0                    # the Schur complement travels from rank 1 to rank 0
