#!/bin/bash
# env-variable sweep of the factorisation knobs on configs[3]: one short bench per setting, ms per pass printed
TAG=${1:-sweep}
OUT=gpurun_out/${TAG}_sweep.txt
: > $OUT
run() {
  label="$1"; shift
  ms=$(env "$@" python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; print(json.loads(sys.stdin.readlines()[-1])['ms_per_step'])")
  echo "$label $ms" | tee -a $OUT
}
run default FERS_NOOP=1
run min_k=512 FERS_CHOL_OZAKI_MIN_K=512
run min_k=640 FERS_CHOL_OZAKI_MIN_K=640
run min_k=1024 FERS_CHOL_OZAKI_MIN_K=1024
run dsplit=2 FERS_OZ_DSPLIT=2
run dsplit=4 FERS_OZ_DSPLIT=4
run group=4 FERS_CHOL_GROUP=4
run group=8 FERS_CHOL_GROUP=8
run outer=3 FERS_CHOL_OUTER=3
run outer=6 FERS_CHOL_OUTER=6
run panel_nf=16 FERS_CHOL_OZAKI_PANEL_NF=16
run panel_kw=1024 FERS_CHOL_OZAKI_PANEL_KW=1024
run default_again FERS_NOOP=1
