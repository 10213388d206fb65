#!/bin/bash
O=gpurun_out; T=${1:-r2s}
run() { echo "== $*"; env "$@" timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $O/${T}_b.json 2>$O/${T}_bench.err; python tools/bench_summary.py < $O/${T}_b.json 2>&1 | grep -E "value=|factor_ms" | cut -c1-200; }
run FERS_CHOL_OZAKI_PANEL_KW=1024 FERS_CHOL_GROUP=4
run FERS_CHOL_OZAKI_PANEL_KW=1280 FERS_CHOL_GROUP=5
run FERS_CHOL_OZAKI_PANEL_KW=1536 FERS_CHOL_GROUP=6 FERS_CHOL_OZAKI_PANEL_K=1536 FERS_CHOL_OZAKI_PANEL_NF=16
run FERS_CHOL_OZAKI_MIN_K=512
run FERS_CHOL_OZAKI_PANEL_KW=2048 FERS_CHOL_GROUP=8
