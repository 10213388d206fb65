set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/r1f_bench.json 2> gpurun_out/r1f_bench.err; echo "bench rc $?"
tail -c 3000 gpurun_out/r1f_bench.err
cat gpurun_out/r1f_bench.json | head -c 6000
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r1f_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r1f_ncu_l.log 2>&1; echo "ncu list rc $?"
ncu --set full --clock-control none --import-source on -k regex:k_syrk --launch-skip 9 --launch-count 1 -o gpurun_out/r1f_syrk -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r1f_ncu_syrk.log 2>&1; echo "ncu syrk rc $?"
