"""Tree-parallel sparse solve on two GPUs against the single-GPU solve of the same model (chol.TreeParallelSolver):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/tree_probe.py [nx nz ny] [steps]
Prints the largest relative difference of the displacements, the residual of both solves and the time per pass."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fers_core_b200 import engine, synth  # noqa: E402

nx, nz, ny = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (24, 24, 12)
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
assert world >= 2 and (world & (world - 1)) == 0
model = synth.grid_frame(nx, nz, ny, n_combos=64)
factors = model.combo_factors[:64]


def timed(eng, n):
    for _ in range(2):
        rec = eng.run_linear(factors)
    torch.cuda.synchronize(dev); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        rec = eng.run_linear(factors)
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return rec, float(t.item())


tp = engine.Engine(model, dev, solver="cholesky", tree_parallel=(rank, world))
rec_tp, ms_tp = timed(tp, steps)
u_tp = rec_tp["U"].clone()
res_tp = tp.stats.relres
sol = tp._chol
kept = (sol.fronts_kept, sol.k_blocks_kept)
link = dict(sol.link_bytes)
del tp, rec_tp
torch.cuda.empty_cache()
one = engine.Engine(model, dev, solver="cholesky")              # every rank solves the whole model on its own GPU
rec_1, ms_1 = timed(one, steps)
diff = float((u_tp - rec_1["U"]).abs().max() / rec_1["U"].abs().max())
print(f"rank {rank}: {model.n_dof} DOF x 64; fronts kept {kept[0]}, K blocks kept {kept[1]}; tree-parallel {ms_tp:.1f} ms per pass "
      f"(residual {res_tp:.2e}), one GPU {ms_1:.1f} ms (residual {one.stats.relres:.2e}); speed-up {ms_1 / ms_tp:.2f}; "
      f"max rel diff of U {diff:.2e}; link bytes {link}", flush=True)
dist.barrier()
dist.destroy_process_group()
