#!/bin/bash
# probe round: TS-mode layout + issue rates + bulk-copy ceiling, new assembly kernel, int8 kernel with A through tensor memory
O=gpurun_out; T=${1:-r2n}
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/ts_probe tools/micro/ts_probe.cu 2>/dev/null && timeout 120 /tmp/ts_probe > $O/${T}_ts_probe.txt 2>&1; echo "ts_probe rc=$?"; cat $O/${T}_ts_probe.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "plain_gather or assembled_values or gather_assembly" 2>&1 | tail -3
FERS_OZ_TS=1 timeout 600 python -m pytest tests/test_ozaki_gpu.py -q -x 2>&1 | tail -3
for ts in 0 1; do echo "== FERS_OZ_TS=$ts"; FERS_OZ_TS=$ts timeout 300 python tools/oz_probe.py --shapes 4860x720,13680x2880,11160x5580 --slices 7,6 2>&1 | tail -8; done
for ts in 0 1; do FERS_OZ_TS=$ts timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $O/${T}_bench_ts$ts.json 2>$O/${T}_bench_ts$ts.err; echo "bench ts=$ts rc=$?"; python tools/bench_summary.py < $O/${T}_bench_ts$ts.json 2>&1 | head -30; done
