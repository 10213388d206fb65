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


# ---- PlateElement.local_x_direction (pydantic_models.py:1703): the axes the plate resultants are reported in ----------
def _plate_doc(direction):
    import copy
    doc = copy.deepcopy(helpers.golden("plate_strip_surface_load")["doc"])
    for p in doc["model"]["plate_elements"]:
        p["local_x_direction"] = direction
    return doc


def _rotate_resultants(res, theta):
    """nx ny nxy | mx my mxy | qx qy of axes turned by theta about the normal."""
    c, s = np.cos(theta), np.sin(theta)
    out = np.empty_like(res)
    for o in (0, 3):
        xx, yy, xy = res[..., o], res[..., o + 1], res[..., o + 2]
        out[..., o] = c * c * xx + s * s * yy + 2 * s * c * xy
        out[..., o + 1] = s * s * xx + c * c * yy - 2 * s * c * xy
        out[..., o + 2] = -s * c * xx + s * c * yy + (c * c - s * s) * xy
    out[..., 6] = c * res[..., 6] + s * res[..., 7]
    out[..., 7] = -s * res[..., 6] + c * res[..., 7]
    return out


def test_plate_local_x_direction_turns_the_reported_axes_only():
    from oracle import fers_oracle
    base_doc = _plate_doc(None)
    base = fers_oracle.analyze(wire.parse_document(base_doc))
    same = wire.parse_document(_plate_doc([1, 0, 0]))
    assert not same.p_data[:, wire.PD_THETA].any() and not [w for w in same.warnings if "LOCAL_X" in w[0]]
    model = wire.parse_document(_plate_doc([1.0, 1.0, 5.0]))           # out-of-plane part is projected away
    theta = model.p_data[:, wire.PD_THETA]
    np.testing.assert_allclose(np.abs(theta), np.pi / 4, rtol=1e-14)
    res = fers_oracle.analyze(model)
    np.testing.assert_allclose(res.U, base.U, rtol=1e-9, atol=1e-14)             # isotropic: the answer does not move
    scale = np.abs(base.p_res).max(axis=(0, 1))
    for e in range(model.n_plates):
        want = _rotate_resultants(base.p_res[:, e], theta[e])
        np.testing.assert_allclose(res.p_res[:, e], want, rtol=1e-8, atol=1e-9 * scale.max())
    np.testing.assert_allclose(res.p_cdg, base.p_cdg, rtol=1e-9, atol=1e-14)
    # the strip lies in a global plane: a direction along its normal cannot define local x
    n = np.cross(model.xyz[model.p_nodes[0, 1]] - model.xyz[model.p_nodes[0, 0]], model.xyz[model.p_nodes[0, 2]] - model.xyz[model.p_nodes[0, 0]])
    normal = wire.parse_document(_plate_doc([float(v) for v in n]))
    assert "W_PLATE_LOCAL_X_NORMAL" in [w[0] for w in normal.warnings] and not normal.p_data[:, wire.PD_THETA].any()


@pytest.mark.gpu
def test_gpu_plate_local_x_direction_matches_oracle():
    from fers_core_b200 import engine
    from oracle import fers_oracle
    model = wire.parse_document(_plate_doc([1.0, 1.0, 5.0]))
    want = fers_oracle.analyze(model)
    got = engine.analyze(model)
    helpers.assert_results_close(got, want, 1e-9, label="plate local x")
