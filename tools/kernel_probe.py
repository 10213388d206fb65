"""HBM-bound kernels of the path on the configs[3] frame, a few launches each (timing with CUDA events;
the same command under ncu captures one launch per kernel):  python tools/kernel_probe.py [nx nz ny]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from fers_core_b200 import _lib, engine, synth  # noqa: E402

nx, nz, ny = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (60, 60, 30)
model = synth.grid_frame(nx, nz, ny, n_combos=64)
dev = torch.device("cuda", 0)
lib = _lib.load()
peak = 6556.5


def timed(fn, reps=5, pre=None):
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = []
    for _ in range(reps):
        if pre is not None:
            pre()
        a0.record()
        fn()
        a1.record()
        torch.cuda.synchronize(dev)
        out.append(a0.elapsed_time(a1))
    return float(np.median(out)), float(min(out))


def report(name, nbytes, t):
    med, best = t
    print(f"{name:34s} {med*1e3:8.1f} us (best {best*1e3:7.1f})  {nbytes/med/1e6:7.0f} GB/s  {nbytes/med/1e6/peak:5.3f} of HBM", flush=True)


res = {}
for variant in (engine.ASSEMBLY_GATHER, 0):
    eng = engine.Engine(model, dev, assembly_variant=variant)
    t = timed(lambda: eng.assemble(eng.d_xyz0), reps=3)
    print(f"engine.assemble variant {variant} (zero + kernels + BC): {t[0]*1e3:.1f} us", flush=True)
    res[variant] = eng.d_vals.clone()
print("gather vs scatter max rel diff", float((res[0] - res[4]).abs().max() / res[0].abs().max()))
nnzb, nn, nm, nd = eng.nnzb, model.n_nodes, model.n_members, model.n_dof
eng = engine.Engine(model, dev)
eng.assemble(eng.d_xyz0)
P = _lib.ptr
report("k_assemble_gather", nm * 64 + nnzb * 288, timed(lambda: _lib.check(lib.fers_b200_assemble_beams_gather(
    nn, nm, P(eng.d_xyz0), P(eng.d_n1), P(eng.d_n2), P(eng.d_m_prop), P(eng.d_type), P(eng.d_roll), None, None, P(eng.d_prop),
    eng.n_prop, None, 0, None, 0, 0, P(eng.d_m_slot), P(eng.d_hm_ptr), P(eng.d_hm_list), eng.n_dup_half_members,
    P(eng.d_diag_slot), P(eng.d_member_cache), P(eng.d_vals), eng.stream), "gather")))
for nr in (64, 1):
    x = torch.randn((nd, nr), dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    b = nnzb * 292 + (nn + 1) * 4 + 2 * nd * nr * 8 + nd
    for v in ((0, 1, 2) if nr == 64 else (0,)):
        lib.fers_b200_set_spmm_variant(v)
        report(f"spmm nrhs={nr} variant {v}", b, timed(lambda: _lib.check(lib.fers_b200_spmm(
            nn, nr, P(eng.d_rowptr), P(eng.d_colidx), P(eng.d_vals), P(eng.d_fixed), None, P(x), P(y), eng.stream), "spmm")))
lib.fers_b200_set_spmm_variant(1)
