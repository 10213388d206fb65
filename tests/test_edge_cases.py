"""Degenerate documents at the reference boundary (fers_calculations.calculate_from_json): what fails loudly, what
solves to zeros, what is tolerated.  The oracle and the GPU path must agree on every case that solves."""
import copy
import json

import numpy as np
import pytest

from fers_core_b200 import wire
from oracle import fers_oracle
from tests import helpers


def _base():
    return copy.deepcopy(helpers.golden("cantilever_end_load")["doc"])


def _unloaded():
    d = _base()
    d["analysis"]["load_cases"][0]["nodal_loads"] = []
    return d


def _floating_node():
    d = _base()
    d["model"]["nodes"].append({"id": 77, "X": 5.0, "Y": 5.0, "Z": 5.0, "nodal_support": None})
    return d


def _large_ids():
    d = _base()
    d["model"]["nodes"][0]["id"] = 2_000_000_000
    d["model"]["members"][0]["start_node_id"] = 2_000_000_000
    d["model"]["members"][0]["id"] = 1_999_999_999
    for ms in d["model"]["member_sets"]:
        ms["member_ids"] = [1_999_999_999]
    return d


def _no_members():
    d = _base()
    d["model"]["members"] = []
    d["model"]["member_sets"] = []
    return d


def _nothing_to_solve():
    d = _base()
    d["analysis"]["options"]["solve_loadcases"] = False
    d["analysis"]["load_combinations"] = []
    return d


def _no_load_cases():
    d = _base()
    d["analysis"]["load_cases"] = []
    d["analysis"]["load_combinations"] = []
    return d


def _mechanism():
    d = _base()
    d["model"]["nodes"][0]["nodal_support"] = None
    return d


SOLVABLE = {"unloaded": _unloaded, "floating_node": _floating_node, "large_ids": _large_ids, "no_members": _no_members}
EMPTY = {"nothing_to_solve": _nothing_to_solve, "no_load_cases": _no_load_cases}


@pytest.mark.parametrize("name", sorted(SOLVABLE))
def test_oracle_on_degenerate_documents(name):
    m = wire.parse_document(SOLVABLE[name]())
    res = fers_oracle.analyze(m)
    assert np.isfinite(res.U).all() and np.isfinite(res.R).all()
    if name in ("unloaded", "no_members"):
        assert not res.U.any() and not res.R.any()
    if name == "floating_node":
        assert not res.U[:, 2].any()                       # zero-stiffness DOFs are held, not solved for
    if name == "large_ids":
        ref = fers_oracle.analyze(wire.parse_document(_base()))
        np.testing.assert_array_equal(res.U, ref.U)


def test_oracle_rejects_a_mechanism():
    with pytest.raises(Exception, match="(?i)singular|mechanism|positive definite"):
        fers_oracle.analyze(wire.parse_document(_mechanism()))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SOLVABLE))
def test_gpu_matches_oracle_on_degenerate_documents(name):
    from fers_core_b200 import engine
    doc = SOLVABLE[name]()
    want = fers_oracle.analyze(wire.parse_document(doc))
    got = engine.analyze(wire.parse_document(doc))
    helpers.assert_results_close(got, want, 1e-9, label=name)
    if name == "floating_node":
        assert not got.U[:, 2].any()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(EMPTY))
def test_gpu_boundary_with_nothing_to_solve(name):
    import fers_calculations
    out = json.loads(fers_calculations.calculate_from_json(json.dumps(EMPTY[name]())))
    res = out.get("results") or out
    assert res["loadcases"] == {} and res["loadcombinations"] == {}


@pytest.mark.gpu
def test_gpu_boundary_errors_are_loud():
    import fers_calculations
    from fers_core_b200 import engine
    with pytest.raises(engine.EngineError, match="mechanism"):
        fers_calculations.calculate_from_json(json.dumps(_mechanism()))
    with pytest.raises(ValueError):
        fers_calculations.calculate_from_json("{ not json")
    with pytest.raises(wire.WireError):
        fers_calculations.calculate_from_json(json.dumps({"model": {}}))
    with pytest.raises(wire.WireError, match="zero length"):
        d = _base()
        d["model"]["nodes"][1].update(X=0.0, Y=0.0, Z=0.0)
        fers_calculations.calculate_from_json(json.dumps(d))
