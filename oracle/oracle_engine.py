"""``calculate_from_json`` twin that runs the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Mirrors the reference boundary (fers_core/fers/fers.py:300) so the reference's own tests can be
pointed at the oracle to pin it (tests/test_oracle_reference_suite.py).
"""
import json

from fers_core_b200 import wire
from fers_core_b200.results import results_to_wire
from . import fers_oracle


def solve_document(doc: dict) -> dict:
    model = wire.parse_document(doc)
    res = fers_oracle.analyze(model)
    res.method = "Linear" if model.options.order == "LINEAR" else "P_DELTA"
    out = dict(doc)
    out["results"] = results_to_wire(model, res)
    return out


def calculate_from_json(input_json: str) -> str:
    return json.dumps(solve_document(json.loads(input_json)))


def calculate_from_file(path: str) -> str:
    with open(path, "r", encoding="utf-8") as fh:
        return calculate_from_json(fh.read())
