set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/r1n_bench.json 2> gpurun_out/r1n_bench.err; echo "bench rc $?"
tail -c 1000 gpurun_out/r1n_bench.err
python tools/bench_summary.py r1n < gpurun_out/r1n_bench.json
