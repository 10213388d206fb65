set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size or gather" 2>&1 | tail -5
