"""The one document the reference ships WITH results of its own engine (fers_core/examples/json_input_solver/
805_JSON_Import_Export_with_results.json: fixed-fixed IPE 200, 6 m, 20 kN at mid-span, order Nonlinear, an older engine):
the input without the results and the handful of numbers the tests compare.  Writes tests/golden/engine_result_805.json."""
import json
import os

REFERENCE = os.environ.get("FERS_REFERENCE", "/root/reference")
with open(os.path.join(REFERENCE, "fers_core", "examples", "json_input_solver", "805_JSON_Import_Export_with_results.json"),
          encoding="utf-8") as fh:
    full = json.load(fh)
case = full["results"]["loadcases"]["Center load"]
doc = {k: v for k, v in full.items() if k != "results"}
doc["model"]["shape_paths"] = []
for s in doc["model"]["sections"]:
    s["shape_path"] = None
golden = {"loadcase": "Center load",
          "displacement_nodes": {n: {c: v[c] for c in ("dx", "dy", "dz", "rx", "ry", "rz")} for n, v in case["displacement_nodes"].items()},
          "reaction_nodes": {n: v["nodal_forces"] for n, v in case["reaction_nodes"].items()}}
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "engine_result_805.json")
with open(out, "w", encoding="utf-8") as fh:
    json.dump({"doc": doc, "golden": golden}, fh)
print("->", out)
