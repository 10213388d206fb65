"""One sparse-Cholesky factorisation of the headline frame (for an ncu launch list):
python tools/factor_once.py [nx nz ny].  Prints the per-level plan table (fronts, own nodes, front nodes, Mflop)."""
import sys

import torch

sys.path.insert(0, ".")
from fers_core_b200 import chol, engine, synth  # noqa: E402

nx, nz, ny = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (60, 60, 30)
model = synth.grid_frame(nx, nz, ny, n_combos=64)
eng = engine.Engine(model, "cuda:0", solver="cholesky")
eng.assemble(eng.d_xyz0)
plan = chol.CholPlan(eng.d_rowptr.cpu().numpy(), eng.d_colidx.cpu().numpy(), model.xyz)
print("levels [fronts, max own nodes, max front nodes, Mflop]:")
for l, row in enumerate(plan.levels()):
    print(l, row.tolist())
sol = chol.CholSolver(plan, eng.dev, max_nrhs=64)
info = sol.factorize(eng.d_vals, eng.d_diag_slot, eng.d_fixed, eng.d_ks)
torch.cuda.synchronize()
print("info", info)
f_case, peq_case = eng.load_cases(eng.d_xyz0)
d_f = eng.combine(f_case, model.combo_factors[:64], 64)
u = torch.empty((model.n_dof, 64), dtype=torch.float64, device=eng.dev)
sol.solve(d_f, u, eng.d_fixed)
torch.cuda.synchronize()
print("solved", float(u.abs().max()))
