"""``calculate_from_json`` twin that runs the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Mirrors the reference boundary (fers_core/fers/fers.py:300) so the reference's own tests can be
pointed at the oracle to pin it (tests/test_oracle_reference_suite.py).
"""
import json

from fers_core_b200 import wire
from fers_core_b200.results import results_to_wire
from . import fers_oracle, unity_oracle


def solve_document(doc: dict) -> dict:
    model = wire.parse_document(doc)
    res = fers_oracle.analyze(model)
    res.method = "Linear" if model.options.order == "LINEAR" else "P_DELTA"
    if model.unity_checks:
        res.unity = _OracleUnity(model, res)
    out = dict(doc)
    out["results"] = results_to_wire(model, res)
    return out


_EC3_LABELS = {"u_n": "Axial (6.2.3 / 6.2.4)", "u_my": "Bending y (6.2.5)", "u_mz": "Bending z (6.2.5)", "u_vy": "Shear y (6.2.6)",
               "u_vz": "Shear z (6.2.6)", "u_nm": "N + M (6.2.1(7))", "u_by": "Flexural buckling y (6.3.1)",
               "u_bz": "Flexural buckling z (6.3.1)", "u_ltb": "LTB (6.3.2)", "u_int": "N + M buckling (6.3.3, k = 1)",
               "C1": "C1 (6.3.2)", "Mcr": "M_cr", "chilt": "chi_LT", "chiy": "chi_y", "chiz": "chi_z"}


class _OracleUnity:
    """Unity-check results of the oracle in the wire layout (UnityCheckResult, pydantic_models.py:2005-2012)."""

    def __init__(self, model, res):
        self.model, self.res = model, res

    @staticmethod
    def _status(u, thresholds):
        t = list(thresholds or [0.8, 0.95, 1.0])
        for name, lim in zip(("Green", "Yellow", "Orange"), t):
            if u == u and u <= lim:
                return name
        return "Red"

    def results(self) -> list:
        out = []
        for d, w in zip(self.model.unity_checks, unity_oracle.evaluate(self.model, self.res)):
            per = []
            for p in w["per_entity"]:
                trace = [{"label": _EC3_LABELS[k], "formula": "", "substituted": "", "unit": "", "value": float(v)}
                         for k, v in (p.get("rules") or {}).items() if k in _EC3_LABELS]
                if p.get("rules"):
                    trace.append({"label": "Governing", "formula": "max of the rules above", "substituted": "", "unit": "",
                                  "value": float(p["utilization"])})
                per.append({"entity": p["entity"], "demand": float(p["demand"]), "capacity": p["capacity"],
                            "utilization": float(p["utilization"]), "status": self._status(p["utilization"], d.get("thresholds")),
                            "governing_combination_id": p["governing_combination_id"], "location_x_frac": None, "message": None,
                            "trace": trace})
            gov = max(per, key=lambda r: r["utilization"]) if per else None
            out.append({"check_id": str(d.get("id")), "name": str(d.get("name", d.get("id"))), "limit_state": d.get("limit_state"),
                        "max_utilization": gov["utilization"] if gov else 0.0, "status": gov["status"] if gov else "Green",
                        "governing": gov, "per_entity": per})
        return out


def calculate_from_json(input_json: str) -> str:
    return json.dumps(solve_document(json.loads(input_json)))


def calculate_from_file(path: str) -> str:
    with open(path, "r", encoding="utf-8") as fh:
        return calculate_from_json(fh.read())
