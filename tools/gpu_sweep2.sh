#!/bin/bash
# A/B of the extend-add kernel variants (columns per thread / resident CTAs) on configs[3]
TAG=${1:-sweep2}
OUT=gpurun_out/${TAG}_init_cols.txt
: > $OUT
for c in 24 18 12 6; do
  FERS_INIT_COLS=$c python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
b=json.loads(sys.stdin.readlines()[-1]); r=b['roofline']
print('init_cols=$c', 'ms_per_step', round(b['ms_per_step'],2), 'init_ms', round(r.get('class_ms_profiled_step',{}).get('init_ms',-1),2), 'resid', b['config'].get('max_rel_residual'))" | tee -a $OUT
done
