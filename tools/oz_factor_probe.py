"""Factorisation time and solve residual with the int8 tensor-core Schur updates on / off:
python tools/oz_factor_probe.py nx nz ny [settings "S:min_k,..." default "0:1024,8:1024,7:1024,6:1024"]"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from fers_core_b200 import _lib, chol, engine, synth, wire  # noqa: E402

nx, nz, ny = (int(a) for a in sys.argv[1:4])
settings = sys.argv[4] if len(sys.argv) > 4 else "0:1024,8:1024,7:1024,6:1024"
nrhs = 64
model = synth.grid_frame(nx, nz, ny, n_combos=nrhs)
eng = engine.Engine(model, "cuda:0", solver="cholesky")
lib = eng.lib
f_case, _ = eng.load_cases(eng.d_xyz0)
eng.assemble(eng.d_xyz0)
d_f = eng.combine(f_case, model.combo_factors[:nrhs], nrhs)
restr = (model.bc_code.reshape(-1, 6) == wire.BC_FIXED).all(axis=1)
plan = chol.CholPlan(eng.d_rowptr.cpu().numpy(), eng.d_colidx.cpu().numpy(), model.xyz, node_restrained=restr)
sol = chol.CholSolver(plan, eng.dev, max_nrhs=nrhs)
u = torch.empty((model.n_dof, nrhs), dtype=torch.float64, device=eng.dev)
ku = torch.empty_like(u); r = torch.empty_like(u); norms = torch.empty((2, nrhs), dtype=torch.float64, device=eng.dev)
u_ref = None


def relres():
    _lib.check(lib.fers_b200_spmm(model.n_nodes, nrhs, eng.d_rowptr.data_ptr(), eng.d_colidx.data_ptr(), eng.d_vals.data_ptr(),
                                  eng.d_fixed.data_ptr(), None, u.data_ptr(), ku.data_ptr(), eng.stream), "spmm")
    _lib.check(lib.fers_b200_residual(model.n_dof, nrhs, d_f.data_ptr(), ku.data_ptr(), eng.d_fixed.data_ptr(), r.data_ptr(),
                                      norms.data_ptr(), eng.stream), "residual")
    nh = norms.cpu().numpy()
    return float(np.sqrt(nh[0] / nh[1]).max())


for spec in settings.split(","):
    S, min_k = (int(v) for v in spec.split(":"))
    _lib.check(lib.fers_b200_chol_set_ozaki(S, min_k), "set_ozaki")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for rep in range(3):
        e0.record()
        info = sol.factorize(eng.d_vals, eng.d_diag_slot, eng.d_fixed, eng.d_ks, check=False)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    lib.fers_b200_chol_set_profile(1)
    sol.factorize(eng.d_vals, eng.d_diag_slot, eng.d_fixed, eng.d_ks)
    prof = sol.profile()
    lib.fers_b200_chol_set_profile(0)
    sol.solve(d_f, u, eng.d_fixed)
    rr = [relres()]
    for _ in range(2):
        sol.solve(r, u, eng.d_fixed, accumulate=True)
        rr.append(relres())
    if u_ref is None:
        u_ref = u.clone()
    du = float(((u - u_ref).abs().amax(dim=0) / u_ref.abs().amax(dim=0)).max())
    print(json.dumps({"ozaki_slices": S, "min_k": min_k, "factor_ms": min(ts), "factor_tflops": plan.info.flops / min(ts) / 1e9,
                      "syrk_ms_profiled": prof["syrk_ms"], "relres_after_0_1_2_refinements": rr,
                      "max_rel_diff_of_u_vs_first_setting_after_refinement": du}), flush=True)
