#!/bin/bash
O=gpurun_out; T=${1:-r2y}
python bench.py --steps 3 --warmup 3 --detail --no-cpu-baseline --no-e2e > $O/${T}_bench_detail.json 2>$O/${T}_bench_detail.err; echo "detail rc=$?"; python tools/bench_summary.py < $O/${T}_bench_detail.json 2>&1 | grep -E "value=|k_assemble" | cut -c1-300
python bench.py --nx 76 --nz 76 --ny 30 --steps 3 --warmup 3 --no-cpu-baseline > $O/${T}_bench_1p1M_dof.json 2>$O/${T}_bench_1p1M.err; echo "1.1M rc=$?"; tail -2 $O/${T}_bench_1p1M.err; python tools/bench_summary.py < $O/${T}_bench_1p1M_dof.json 2>&1 | grep -E "value=|e2e|k_assemble|k_spm" | cut -c1-300
