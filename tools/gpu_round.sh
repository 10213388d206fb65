python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r1p_bench_cfg5.json 2> gpurun_out/r1p_cfg5.err
cut -c1-900 gpurun_out/r1p_bench_cfg5.json
python bench.py --workload cfg5 --steps 5 --warmup 3 2>/dev/null | cut -c1-200
