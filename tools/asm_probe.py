"""Gather assembly on the configs[3] frame: the element-lane kernel (kind 0) against the per-half-member kernel (kind 1),
CUDA events around single launches.  The same command under ncu captures both:  python tools/asm_probe.py [nx nz ny]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from fers_core_b200 import _lib, engine, synth  # noqa: E402

nx, nz, ny = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (60, 60, 30)
model = synth.grid_frame(nx, nz, ny, n_combos=8)
dev = torch.device("cuda", 0)
lib = _lib.load()
eng = engine.Engine(model, dev)
P = _lib.ptr
nbytes = model.n_members * 64 + eng.nnzb * 288


def launch():
    _lib.check(lib.fers_b200_assemble_beams_gather(
        model.n_nodes, model.n_members, P(eng.d_xyz0), P(eng.d_n1), P(eng.d_n2), P(eng.d_m_prop), P(eng.d_type), P(eng.d_roll), None, None,
        P(eng.d_prop), eng.n_prop, None, 0, None, 0, 0, P(eng.d_m_slot), P(eng.d_hm_ptr), P(eng.d_hm_list), eng.n_dup_half_members,
        P(eng.d_diag_slot), None, P(eng.d_vals), eng.stream), "gather")


out = {}
for kind in (1, 0):
    _lib.check(lib.fers_b200_assemble_set_gather_kernel(kind), "kind")
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(7):              # 10 launches per event pair: one launch alone carries ~10 us of host launch gap
        a0.record()
        for _ in range(10):
            launch()
        a1.record()
        torch.cuda.synchronize(dev)
        ts.append(a0.elapsed_time(a1) * 1e2)
    out[kind] = eng.d_vals.clone()
    med = float(np.median(ts))
    print(f"kind {kind}: {med:7.1f} us (best {min(ts):7.1f})  {nbytes / med / 1e3:6.0f} GB/s", flush=True)
print("max rel diff", float((out[0] - out[1]).abs().max() / out[1].abs().max()))
lib.fers_b200_assemble_set_gather_kernel(0)
