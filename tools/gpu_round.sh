python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r1s_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log | cut -c1-200
