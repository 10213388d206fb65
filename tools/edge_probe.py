"""Behaviour of the document boundary on degenerate inputs (probe; tests/test_edge_cases.py codifies it)."""
import copy, json, sys, traceback
sys.path.insert(0, ".")
import numpy as np
from tests import helpers
import fers_calculations as fc

def base():
    return copy.deepcopy(helpers.golden("cantilever_end_load")["doc"])

def run(label, doc):
    try:
        out = json.loads(fc.calculate_from_json(json.dumps(doc)))
        r = out.get("results") or out
        keys = list(r.keys())[:8] if isinstance(r, dict) else type(r)
        print(f"[{label}] ok keys={keys}")
        return out
    except Exception as e:
        print(f"[{label}] {type(e).__name__}: {str(e)[:200]}")

d = base(); d["model"]["nodes"][0]["nodal_support"] = None
run("no supports (mechanism)", d)
d = base(); d["analysis"]["load_cases"][0]["nodal_loads"] = []
o = run("load case without loads", d)
d = base(); d["analysis"]["load_cases"] = []; d["analysis"]["load_combinations"] = []
run("no load cases at all", d)
d = base(); d["model"]["nodes"].append({"id": 77, "X": 5, "Y": 5, "Z": 5, "nodal_support": None})
run("floating unconnected node", d)
d = base(); d["model"]["nodes"][1].update(X=0.0, Y=0.0, Z=0.0)
run("zero-length member", d)
d = base(); d["model"]["members"] = []
run("no members", d)
d = base(); d["model"]["nodes"][0]["id"] = 1000000; d["model"]["members"][0]["start_node_id"] = 1000000
run("large node ids", d)
d = base(); d["analysis"]["options"]["solve_loadcases"] = False; d["analysis"]["load_combinations"] = []
run("nothing to solve", d)
d = base(); d["model"]["materials"][0]["e_mod"] = -1.0
run("negative E", d)
d = base(); d["analysis"]["load_cases"][0]["nodal_loads"][0]["magnitude"] = float("nan")
run("NaN load", d)
run("not json", "{")
