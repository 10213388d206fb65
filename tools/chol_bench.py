"""Sparse-Cholesky timing on a synthetic grid frame: python tools/chol_bench.py nx nz ny [nrhs] [leaf]."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from fers_core_b200 import _lib, chol, engine, synth  # noqa: E402

nx, nz, ny = (int(a) for a in sys.argv[1:4])
nrhs = int(sys.argv[4]) if len(sys.argv) > 4 else 64
leaf = int(sys.argv[5]) if len(sys.argv) > 5 else chol.DEFAULT_LEAF_NODES
model = synth.grid_frame(nx, nz, ny, n_combos=nrhs)
eng = engine.Engine(model, "cuda:0", solver="cholesky")
lib = eng.lib
f_case, peq_case = eng.load_cases(eng.d_xyz0)
eng.assemble(eng.d_xyz0)
d_f = eng.combine(f_case, model.combo_factors[:nrhs], nrhs)
t0 = time.perf_counter()
plan = chol.CholPlan(eng.d_rowptr.cpu().numpy(), eng.d_colidx.cpu().numpy(), model.xyz, leaf_nodes=leaf)
t_plan = time.perf_counter() - t0
print(f"dof {model.n_dof} fronts {plan.info.n_fronts} levels {plan.info.n_levels} Tflop {plan.info.flops/1e12:.3f} "
      f"factor GB {plan.info.factor_doubles*8/1e9:.2f} plan {t_plan*1e3:.0f} ms", flush=True)
sol = chol.CholSolver(plan, eng.dev, max_nrhs=nrhs)
lib.fers_b200_chol_set_profile(1)
import os
variants = [int(v) for v in os.environ.get("CHOL_VARIANTS", "0").split(",")]
for rep in variants:
    lib.fers_b200_chol_set_variant(rep)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    info = sol.factorize(eng.d_vals, eng.d_diag_slot, eng.d_fixed, eng.d_ks)
    torch.cuda.synchronize()
    t_f = time.perf_counter() - t0
    u = torch.empty((model.n_dof, nrhs), dtype=torch.float64, device=eng.dev)
    t0 = time.perf_counter()
    sol.solve(d_f, u, eng.d_fixed)
    torch.cuda.synchronize()
    t_s = time.perf_counter() - t0
    p = sol.profile()
    print(f"variant {rep}: info {info} factor {t_f*1e3:.1f} ms ({plan.info.flops/t_f/1e12:.2f} TF/s) solve {t_s*1e3:.1f} ms", flush=True)
    print("   ", {k: round(v, 2) for k, v in p.items()}, flush=True)
    gemm_ms = p["update_ms"] + p["syrk_ms"]
    print(f"    gemm share {gemm_ms/p['factor_ms']:.2f}  gemm TF/s (all flops / gemm time) {plan.info.flops/gemm_ms/1e9:.2f}", flush=True)
# residual
ku = torch.empty_like(u); r = torch.empty_like(u); norms = torch.empty((2, nrhs), dtype=torch.float64, device=eng.dev)
for it in range(3):
    _lib.check(lib.fers_b200_spmm(model.n_nodes, nrhs, eng.d_rowptr.data_ptr(), eng.d_colidx.data_ptr(), eng.d_vals.data_ptr(),
                                  eng.d_fixed.data_ptr(), None, u.data_ptr(), ku.data_ptr(), eng.stream), "spmm")
    _lib.check(lib.fers_b200_residual(model.n_dof, nrhs, d_f.data_ptr(), ku.data_ptr(), eng.d_fixed.data_ptr(), r.data_ptr(),
                                      norms.data_ptr(), eng.stream), "residual")
    nh = norms.cpu().numpy()
    print(f"refine {it}: max relres {np.sqrt(nh[0]/nh[1]).max():.3e}  |u|max {float(u.abs().max()):.6e}", flush=True)
    sol.solve(r, u, eng.d_fixed, accumulate=True)
lib.fers_b200_chol_set_profile(0)
for spec in os.environ.get("CHOL_OUTERS", "4x1").split(","):
    outer, group = (int(v) for v in spec.split("x"))
    lib.fers_b200_chol_set_outer(outer, group)
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sol.factorize(eng.d_vals, eng.d_diag_slot, eng.d_fixed, eng.d_ks, check=False)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        sol.solve(d_f, u, eng.d_fixed)
        torch.cuda.synchronize()
        print(f"outer {outer} group {group}: unprofiled factor {(t1-t0)*1e3:.1f} ms ({plan.info.flops/(t1-t0)/1e12:.2f} TF/s) solve {(time.perf_counter()-t1)*1e3:.1f} ms", flush=True)
lib.fers_b200_chol_set_outer(4, 1)
