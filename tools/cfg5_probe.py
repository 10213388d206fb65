"""Phase timing of the configs[4] batch path (1024 frames x 270 DOF, P-Delta): python tools/cfg5_probe.py [frames]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from fers_core_b200 import _lib, engine, synth  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
big, _ = synth.portal_batch(frames, seed=synth.BATCH_SEED)
dev = torch.device("cuda", 0)
lib = _lib.load()
acc = {}


def timed(obj, name, label=None):
    fn = getattr(obj, name)
    label = label or name

    def wrap(*a, **k):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize(dev)
        acc.setdefault(label, [0, 0.0])
        acc[label][0] += 1
        acc[label][1] += (time.perf_counter() - t0) * 1e3
        return out
    setattr(obj, name, wrap)


for name in ("_upload_static", "_upload_load_cases", "_build_pattern", "load_cases", "assemble", "combine", "recover"):
    timed(engine.Engine, name)
for name in ("fers_b200_dense_expand_batched", "fers_b200_dense_cholesky", "fers_b200_dense_solve", "fers_b200_member_axial"):
    timed(lib, name)
for rep in range(3):
    acc.clear()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    stats = {}
    engine.analyze_frame_batch(big, synth.PORTAL_NODES, dev, stats_out=stats)
    torch.cuda.synchronize(dev)
    tot = (time.perf_counter() - t0) * 1e3
    print(f"pass {rep}: {tot:.1f} ms, iterations max {int(stats['frame_iterations'].max())}", flush=True)
    for k, (n, ms) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print(f"    {k:34s} {n:3d} x  {ms / n:7.2f} ms = {ms:7.2f} ms", flush=True)
    print(f"    {'(unaccounted host / torch ops)':34s}          {tot - sum(v[1] for v in acc.values()):7.2f} ms", flush=True)
