set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu-baseline > gpurun_out/r1m_bench.json 2> gpurun_out/r1m_bench.err; echo "bench rc $?"
tail -c 1000 gpurun_out/r1m_bench.err
python tools/bench_summary.py r1m < gpurun_out/r1m_bench.json
python bench.py --nx 76 --nz 76 --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/r1m_bench_1p1M.json 2>> gpurun_out/r1m_bench.err; echo "bench rc $?"
python tools/bench_summary.py r1m_1.1M < gpurun_out/r1m_bench_1p1M.json
