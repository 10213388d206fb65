python -m pytest tests/test_chol_gpu.py -q -x -k host_pattern 2>&1 | tail -4
