python -m pytest tests/test_edge_cases.py tests/test_wire_host.py -q -x 2>&1 | tail -15
