python -m pytest tests -q -x -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py 2>gpurun_out/r1t_bench.err | tail -1 > gpurun_out/r1t_bench.json; cut -c1-300 gpurun_out/r1t_bench.json
python bench.py --workload cfg5 2>/dev/null | tail -1 > gpurun_out/r1t_bench_cfg5.json; cut -c1-200 gpurun_out/r1t_bench_cfg5.json
