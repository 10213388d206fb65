#!/bin/bash
O=gpurun_out; T=${1:-r2r}
timeout 600 python -m pytest tests/test_ozaki_gpu.py tests/test_gpu_parity.py -q -x -k "ozaki or planes or plane or rows or int8 or plain_gather or assembled_values or gather_assembly" 2>&1 | tail -5
python tools/asm_probe.py 2>&1 | tail -3
for d in 4 3; do echo "== FERS_OZ_DSPLIT=$d"; FERS_OZ_DSPLIT=$d timeout 300 python tools/oz_probe.py --shapes 9000x1536,13680x2880,11160x5580 --slices 7 --no-check 2>&1 | tail -3; done
for cfg in "4 1024" "3 1024" "4 768" "4 512"; do set -- $cfg; echo "== DSPLIT=$1 MIN_K=$2"; FERS_OZ_DSPLIT=$1 FERS_CHOL_OZAKI_MIN_K=$2 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $O/${T}_bench_d$1_k$2.json 2>$O/${T}_bench.err; python tools/bench_summary.py < $O/${T}_bench_d$1_k$2.json 2>&1 | grep -E "value=|k_oz|k_syrk|factor_ms|k_assemble"; done
