set -x
python tools/kernel_probe.py 2>&1 | tail -9
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
