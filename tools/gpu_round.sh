set -x
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r1p_bench_cfg5_n$N.json 2> gpurun_out/r1p_cfg5_n$N.err; echo "rc $?"
cut -c1-700 gpurun_out/r1p_bench_cfg5_n$N.json; tail -c 300 gpurun_out/r1p_cfg5_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r1p_bench_n$N.json 2> gpurun_out/r1p_n$N.err; echo "rc $?"
cut -c1-300 gpurun_out/r1p_bench_n$N.json
