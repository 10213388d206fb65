"""The two building-sized models among the reference's examples (fers_core/examples/201_Simple_Shed.py, 202_Solar_Rail.py) as
wire documents, built by the reference's own front end in the build container:

    PYTHONPATH=tests/refshim:/root/reference:. FERS_SHIM_BACKEND=oracle python tests/golden/make_examples_golden.py

Each example script is executed up to the point where it saves its model outside the repository (the part before
``_cloud_public``); ``model.to_dict()`` is what ``FERS.run_analysis`` sends to the engine.  Results and section outlines
(``shape_paths``, viewer data) are dropped.  Writes tests/golden/example_models.json.gz."""
import gzip
import io
import json
import os
from contextlib import redirect_stdout

REFERENCE = os.environ.get("FERS_REFERENCE", "/root/reference")
out = {}
for name in ("201_Simple_Shed", "202_Solar_Rail"):
    path = os.path.join(REFERENCE, "fers_core", "examples", name + ".py")
    src = open(path, encoding="utf-8").read().split("_cloud_public")[0]
    scope = {"__file__": path, "__name__": "example"}
    with redirect_stdout(io.StringIO()):
        exec(compile(src, name, "exec"), scope)
    doc = scope["model"].to_dict()
    doc.pop("results", None)
    doc["model"]["shape_paths"] = []
    for s in doc["model"]["sections"]:
        s["shape_path"] = None
    out[name] = doc
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "example_models.json.gz")
with gzip.open(path, "wt", encoding="utf-8") as fh:
    json.dump(out, fh)
print({k: (len(v["model"]["nodes"]), len(v["model"]["members"])) for k, v in out.items()}, "->", path)
