"""The two building-sized models of the reference's examples (201_Simple_Shed: portal frames, purlins and bracing; 202_Solar_Rail),
as the reference's own front end serialises them (tests/golden/example_models.json.gz, made by
tests/golden/make_examples_golden.py).  The examples print results without asserting them, so the pins here are the ones that
need no engine: global equilibrium of every result column against the loads written in the document, and symmetry of the shed;
then the GPU path against the oracle on every field, through the arrays and through the JSON boundary."""
import gzip
import json
import os

import numpy as np
import pytest

from fers_core_b200 import wire
from oracle import fers_oracle, oracle_engine
from tests import helpers

with gzip.open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "example_models.json.gz"), "rt", encoding="utf-8") as fh:
    DOCS = json.load(fh)
NAMES = sorted(DOCS)


def _applied(doc, factors):
    """Total applied force of a result column from the DOCUMENT (not from the parsed arrays): nodal loads plus distributed
    loads integrated over the member lengths."""
    xyz = {n["id"]: np.array([n["X"], n["Y"], n["Z"]], float) for n in doc["model"]["nodes"]}
    mem = {m["id"]: m for m in doc["model"]["members"]}
    total = np.zeros(3)
    for lc in doc["analysis"]["load_cases"]:
        f = factors.get(lc["id"], 0.0)
        if f == 0.0:
            continue
        for nl in lc.get("nodal_loads") or []:
            total += f * nl["magnitude"] * np.array(nl["direction"], float)
        for dl in lc.get("distributed_loads") or []:
            m = mem[dl["member"]]
            L = np.linalg.norm(xyz[m["end_node_id"]] - xyz[m["start_node_id"]])
            q2 = dl["magnitude"] if dl.get("end_magnitude") is None else dl["end_magnitude"]
            a, b = dl.get("start_frac") or 0.0, 1.0 if dl.get("end_frac") is None else dl["end_frac"]
            total += f * 0.5 * (dl["magnitude"] + q2) * (b - a) * L * np.array(dl["direction"], float)
    return total


def _column_factors(doc):
    cols = []
    if doc["analysis"]["options"].get("solve_loadcases", True):
        cols += [{lc["id"]: 1.0} for lc in doc["analysis"]["load_cases"]]
    for c in doc["analysis"]["load_combinations"]:
        fac = c["load_cases_factors"] if "load_cases_factors" in c else c["load_case_factors"]
        if isinstance(fac, dict):
            cols.append({int(k): float(v) for k, v in fac.items()})
        else:
            cols.append({int(e["load_case_id"]): float(e["factor"]) for e in fac})
    return cols


@pytest.mark.parametrize("name", NAMES)
def test_oracle_results_are_in_equilibrium_with_the_document_loads(name):
    doc = DOCS[name]
    m = wire.parse_document(doc)
    res = fers_oracle.analyze(m)
    cols = _column_factors(doc)
    assert res.U.shape[0] == len(cols) and np.isfinite(res.U).all()
    for c, fac in enumerate(cols):
        applied = _applied(doc, fac)
        scale = max(np.abs(applied).max(), 1.0)
        assert np.abs(res.R[c, :, :3].sum(0) + applied).max() < 1e-8 * scale, (name, c)
        # every member is in equilibrium with its own loads: end forces sum to minus the member load
        assert np.isfinite(res.f_local[c]).all()


def test_shed_is_symmetric_under_symmetric_loads():
    doc = DOCS["201_Simple_Shed"]
    m = wire.parse_document(doc)
    res = fers_oracle.analyze(m)
    snow = next(i for i, lc in enumerate(doc["analysis"]["load_cases"]) if lc["name"].startswith("S"))
    sup = np.nonzero(m.node_support_id >= 0)[0]
    ry = res.R[snow, sup, 1]
    # four column bases; the frame, the snow load and the bracing pattern are mirror images about the ridge plane
    x = m.xyz[sup, 0]
    left, right = sup[x < x.mean()], sup[x > x.mean()]
    assert len(left) == len(right) == 2
    assert res.R[snow, left, 1].sum() == pytest.approx(res.R[snow, right, 1].sum(), rel=1e-6)
    assert ry.sum() > 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_matches_oracle(name):
    from fers_core_b200 import engine
    m = wire.parse_document(DOCS[name])
    helpers.assert_results_close(engine.analyze(m), fers_oracle.analyze(m), rtol=1e-9, label=name)


@pytest.mark.gpu
def test_json_boundary_on_the_shed():
    from fers_core_b200 import engine
    doc = DOCS["201_Simple_Shed"]
    got = json.loads(engine.calculate_from_json(json.dumps(doc)))["results"]
    want = oracle_engine.solve_document(doc)["results"]
    assert set(got["loadcombinations"]) == set(want["loadcombinations"])
    for key, w in want["loadcombinations"].items():
        g = got["loadcombinations"][key]
        for nid, rn in w["reaction_nodes"].items():
            for comp, val in rn["nodal_forces"].items():
                assert g["reaction_nodes"][nid]["nodal_forces"][comp] == pytest.approx(val, rel=1e-8, abs=1e-6)
