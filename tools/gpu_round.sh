python -m pytest tests/test_results_json.py tests/test_lazy_bundle.py tests/test_edge_cases.py tests/test_gpu_parity.py -q -x -m gpu 2>&1 | tail -3
