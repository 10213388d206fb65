#!/bin/bash
O=gpurun_out; T=${1:-r2v}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_pdelta_closed_forms.py tests/test_ozaki_gpu.py -q -x -k "frame_batch or pdelta or ozaki or planes or int8 or dense" 2>&1 | tail -4
python bench.py --workload cfg5 --steps 5 --warmup 3 > $O/${T}_bench_cfg5.json 2> $O/${T}_bench_cfg5.err; echo "cfg5 rc=$?"; tail -3 $O/${T}_bench_cfg5.err; python - <<'P'
import json
d=json.loads(open("gpurun_out/r2v_bench_cfg5.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["config"].get("pdelta_iterations_mean"), d["config"].get("pdelta_iterations_enqueued"), d["e2e"]["ms_per_step"], d["gpu_launches"])
P
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $O/${T}_bench.json 2>$O/${T}_bench.err; python tools/bench_summary.py < $O/${T}_bench.json 2>&1 | grep -E "value=|factor_ms|class_ms|k_oz" | cut -c1-400
