python -m pytest tests -q -x -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --workload cfg5 2>/dev/null | tail -1 > gpurun_out/r1q_bench_cfg5.json; cut -c1-200 gpurun_out/r1q_bench_cfg5.json
python bench.py 2>gpurun_out/r1q_bench.err | tail -1 > gpurun_out/r1q_bench.json; cut -c1-400 gpurun_out/r1q_bench.json
