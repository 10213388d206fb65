python -m pytest tests -q -x -m gpu 2>&1 | tail -5
python bench.py --workload cfg5 2>/dev/null | tail -1 | cut -c1-200
