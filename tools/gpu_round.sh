#!/bin/bash
# One GPU round: tests, bench lines (both workloads + the reference arm), ncu launch list.  Usage: tools/gpu_round.sh TAG [full]
# "full" adds the ncu --set full capture of every int8 launch of one factorisation (10+ minutes: ncu replays each launch 39 times)
T=${1:-r2m}
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/${T}_gputest.log 2>&1; echo "gputest rc=$?"; tail -2 $O/${T}_gputest.log
python bench.py --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; python tools/bench_summary.py < $O/${T}_bench.json | cut -c1-260
python bench.py --workload cfg5 --steps 20 --warmup 5 > $O/${T}_bench_cfg5.json 2> $O/${T}_bench_cfg5.err; echo "cfg5 rc=$?"; cut -c1-260 $O/${T}_bench_cfg5.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/${T}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
gzip -f $O/${T}_launches.csv
if [ "$2" = "full" ]; then
  python bench.py --impl reference --steps 20 --warmup 5 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "reference rc=$?"; cut -c1-260 $O/${T}_bench_reference.json
  ncu --set full --clock-control none --import-source on -k regex:k_oz_syrk -c 70 -o $O/${T}_oz_syrk python tools/factor_once.py 60 60 30 > $O/${T}_ncu_oz.log 2>&1; echo "ncu oz rc=$?"
  ncu -i $O/${T}_oz_syrk.ncu-rep --page raw --csv > $O/${T}_oz_syrk_raw.csv 2>/dev/null
  rm -f $O/${T}_oz_syrk.ncu-rep
fi
ls -la $O | tail -6
