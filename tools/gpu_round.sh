set -x
python -m pytest tests/test_unity_checks.py tests/test_abi_symbols.py -m gpu -x -q 2>&1 | tail -12
