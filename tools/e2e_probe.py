"""Where the end-to-end time of Engine(model).run_linear(factors) goes (synchronising after every stage)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from fers_core_b200 import engine, synth  # noqa: E402

model = bench.pin_model(synth.grid_frame(60, 60, 30, n_combos=64))
dev = torch.device("cuda", 0)


def timed(obj, name):
    fn = getattr(obj, name)

    def wrap(*a, **k):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize(dev)
        print(f"    {name:24s} {(time.perf_counter() - t0) * 1e3:8.2f} ms", flush=True)
        return out
    setattr(obj, name, wrap)


for name in ("_upload_static", "_upload_load_cases", "_build_pattern", "load_cases", "assemble", "combine", "solve", "recover"):
    timed(engine.Engine, name)
for rep in range(3):
    print(f"pass {rep}", flush=True)
    t0 = time.perf_counter()
    eng = engine.Engine(model, dev)
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    out = eng.run_linear(model.combo_factors)
    torch.cuda.synchronize(dev)
    print(f"  engine {1e3 * (t1 - t0):.1f} ms, run_linear {1e3 * (time.perf_counter() - t1):.1f} ms, plan waited {eng.timings.get('chol_plan_ms'):.1f} ms, plan thread {eng.timings.get('chol_plan_thread_ms', -1):.1f} ms", flush=True)
    del eng, out

# where the constructor's time outside the three stages goes
import cProfile, pstats
for rep in range(2):
    pr = cProfile.Profile()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    pr.enable()
    eng = engine.Engine(model, dev)
    torch.cuda.synchronize(dev)
    pr.disable()
    print(f"constructor under cProfile: {(time.perf_counter() - t0) * 1e3:.1f} ms")
    pstats.Stats(pr).sort_stats("tottime").print_stats(8)
    out = eng.run_linear(model.combo_factors)
    torch.cuda.synchronize(dev)
    del eng, out
