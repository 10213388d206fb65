set -x
for leaf in 24 48 96 160 256; do
CHOL_VARIANTS=0 CHOL_OUTERS=4x1 python tools/chol_bench.py 60 60 30 64 $leaf 2>&1 | grep "outer\|dof "
done
