#!/bin/bash
# tests + bench lines + ncu launch lists (whole bench pass; one factorisation by tree level).  Usage: tools/gpu_round2.sh TAG
T=${1:-r2u}; O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/${T}_gputest.log 2>&1; echo "gputest rc=$?"; tail -3 $O/${T}_gputest.log
python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; python tools/bench_summary.py < $O/${T}_bench.json | cut -c1-300
python bench.py --workload cfg5 --steps 5 --warmup 3 > $O/${T}_bench_cfg5.json 2> $O/${T}_bench_cfg5.err; echo "cfg5 rc=$?"; cut -c1-200 $O/${T}_bench_cfg5.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_factor_launches.csv python tools/factor_once.py 60 60 30 > $O/${T}_ncu_factor.log 2>&1; echo "ncu factor rc=$?"
gzip -f $O/${T}_factor_launches.csv
