python -m pytest tests -q -x -m gpu 2>&1 | tail -3
python bench.py 2>gpurun_out/r1s_bench.err | tail -1 > gpurun_out/r1s_bench.json; cut -c1-330 gpurun_out/r1s_bench.json
