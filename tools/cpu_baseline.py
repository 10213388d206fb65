#!/usr/bin/env python
"""CPU baseline of one pass of the hot path on a synthetic frame (oracle/cpu_frame: C++ / OpenMP / OpenBLAS multifrontal).
    python tools/cpu_baseline.py [--nx 60 --nz 60 --ny 30] [--threads N] [--passes 1]
Prints one JSON line with the phase times and DOF*combinations/s.  Test / bench infrastructure only."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

p = argparse.ArgumentParser()
p.add_argument("--nx", type=int, default=60); p.add_argument("--nz", type=int, default=60); p.add_argument("--ny", type=int, default=30)
p.add_argument("--combos", type=int, default=64); p.add_argument("--threads", type=int, default=0); p.add_argument("--passes", type=int, default=1)
a = p.parse_args()

from fers_core_b200 import synth          # noqa: E402
from oracle import cpu_frame              # noqa: E402

if a.threads:
    cpu_frame.set_threads(a.threads)
model = synth.grid_frame(a.nx, a.nz, a.ny, n_combos=a.combos)
cf = cpu_frame.CpuFrame(model)
out = []
for i in range(a.passes):
    t0 = time.perf_counter()
    cf.run_linear(model.combo_factors)
    dt = time.perf_counter() - t0
    out.append({"pass_s": dt, "dof_combos_per_s": model.n_dof * a.combos / dt, "rel_residual": cf.relres,
                **{k: round(v, 4) for k, v in cf.timings.items()}})
print(json.dumps({"frame": [a.nx, a.nz, a.ny], "dof": model.n_dof, "combinations": a.combos, "threads": cpu_frame.threads(),
                  "host_cores": os.cpu_count(), "factor_tflop": cf.flops / 1e12,
                  "factor_gflops": cf.flops / 1e9 / out[-1]["factor_s"], "passes": out}))
