python -m pytest tests/test_gpu_parity.py tests/test_unilateral_supports.py tests/test_edge_cases.py -q -x -m gpu 2>&1 | tail -3
python bench.py --workload cfg5 2>/dev/null | tail -1 | cut -c1-200
python tools/cfg5_cprofile.py 2>&1 | grep -v Warning | head -18
