"""Where chol_plan_ms goes (pattern D2H, plan, solver buffers + upload)."""
import sys, time
import torch
sys.path.insert(0, ".")
import bench  # noqa: E402
from fers_core_b200 import chol, engine, synth  # noqa: E402

model = bench.pin_model(synth.grid_frame(60, 60, 30, n_combos=64))
dev = torch.device("cuda", 0)
for rep in range(3):
    eng = engine.Engine(model, dev)
    eng.assemble(eng.d_xyz0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rowptr = eng.d_rowptr.cpu().numpy(); colidx = eng.d_colidx.cpu().numpy()
    t1 = time.perf_counter()
    plan = chol.CholPlan(rowptr, colidx, model.xyz, node_restrained=eng.h_node_restrained)
    t2 = time.perf_counter()
    sol = chol.CholSolver(plan, dev, max_nrhs=64)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"pattern D2H {1e3*(t1-t0):.2f} ms, CholPlan {1e3*(t2-t1):.2f} ms, CholSolver {1e3*(t3-t2):.2f} ms, index MB {plan.info.index_bytes/1e6:.1f}", flush=True)
    del sol, plan, eng
