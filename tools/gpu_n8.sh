#!/bin/bash
# multi-GPU bench lines (weak scaling, the driver's launch form).  Usage: tools/gpu_n8.sh TAG N
T=${1:-r2w}; N=${2:-8}; O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $O/${T}_bench_n$N.json 2> $O/${T}_bench_n$N.err; echo "cfg4 n$N rc=$?"; python tools/bench_summary.py < $O/${T}_bench_n$N.json 2>&1 | grep -E "value=|e2e" | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload cfg5 --steps 5 --warmup 3 > $O/${T}_bench_cfg5_n$N.json 2> $O/${T}_bench_cfg5_n$N.err; echo "cfg5 n$N rc=$?"; python tools/bench_summary.py < $O/${T}_bench_cfg5_n$N.json 2>&1 | grep -E "value=|e2e" | cut -c1-300
