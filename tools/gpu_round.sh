set -x
python tools/factor_once.py > gpurun_out/factor_once.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1q_factor_launches.csv python tools/factor_once.py > gpurun_out/ncu_factor.log 2>&1
tail -3 gpurun_out/ncu_factor.log
