set -x
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r1n_bench_n$N.json 2> gpurun_out/r1n_bench_n$N.err; echo "bench rc $?"
tail -c 600 gpurun_out/r1n_bench_n$N.err
python tools/bench_summary.py r1n_n$N < gpurun_out/r1n_bench_n$N.json | head -3
