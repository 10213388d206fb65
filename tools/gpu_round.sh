set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/kernel_probe.py 2>&1 | tail -12
CHOL_VARIANTS=4,0,4,0 python tools/chol_bench.py 60 60 30 2>&1 | tail -16
python bench.py --no-cpu-baseline > gpurun_out/r1h_bench.json 2> gpurun_out/r1h_bench.err; echo "bench rc $?"
tail -c 2000 gpurun_out/r1h_bench.err
python tools/bench_summary.py r1h < gpurun_out/r1h_bench.json
ncu --set full --clock-control none --import-source on -k regex:'k_assemble_gather|k_spmv1' --launch-skip 2 --launch-count 2 -o gpurun_out/r1h_hbm -f python tools/kernel_probe.py > gpurun_out/r1h_ncu.log 2>&1; echo "ncu rc $?"
