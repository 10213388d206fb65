python -m pytest tests/test_unity_checks.py -m gpu -x -q 2>&1 | tail -8
