"""CPU oracle (test infrastructure only -- see fers_oracle.py header)."""
