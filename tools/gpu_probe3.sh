#!/bin/bash
O=gpurun_out; T=${1:-r2q}
FERS_OZ_WIDE=1 timeout 600 python -m pytest tests/test_ozaki_gpu.py -q -x 2>&1 | tail -5
for w in 0 1; do echo "== FERS_OZ_WIDE=$w"; FERS_OZ_WIDE=$w timeout 300 python tools/oz_probe.py --shapes 4860x720,9000x1536,13680x2880,11160x5580 --slices 7,6 2>&1 | tail -8; done
FERS_OZ_WIDE=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $O/${T}_bench_wide.json 2>$O/${T}_bench_wide.err; echo "bench wide rc=$?"; python tools/bench_summary.py < $O/${T}_bench_wide.json 2>&1 | head -12
ncu --set full --clock-control none --import-source on -k regex:k_assemble_gather_plain -c 1 -o $O/${T}_asm python tools/asm_probe.py > $O/${T}_ncu_asm.log 2>&1
ncu -i $O/${T}_asm.ncu-rep --page raw --csv > $O/${T}_asm_raw.csv
ncu -i $O/${T}_asm.ncu-rep --page source --csv > $O/${T}_asm_source.csv 2>/dev/null
ls -la $O | tail -6
