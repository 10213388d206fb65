"""``fers_calculations`` stand-in used ONLY to run the reference's own test-suite in this
container.  FERS_SHIM_BACKEND=oracle (default) routes to the CPU oracle, =product to the CUDA
engine (needs a GPU)."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.append(_root)

if os.environ.get("FERS_SHIM_BACKEND", "oracle") == "product":
    from fers_core_b200.engine import calculate_from_json, calculate_from_file  # noqa: F401
else:
    from oracle.oracle_engine import calculate_from_json, calculate_from_file  # noqa: F401
