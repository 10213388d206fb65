set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/r1o_bench.json 2> gpurun_out/r1o_bench.err; echo "bench rc $?"
tail -c 800 gpurun_out/r1o_bench.err
python tools/bench_summary.py r1o < gpurun_out/r1o_bench.json
python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | cut -c1-700
