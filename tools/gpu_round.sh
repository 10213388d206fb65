CHOL_VARIANTS=0 CHOL_OUTERS=4x1 timeout 600 python tools/chol_bench.py 60 60 30 2>&1 | grep -E "refine 0|unprofiled|update_ms" | cut -c1-170
