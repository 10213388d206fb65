"""Unilateral supports (SupportConditionType PositiveOnly / NegativeOnly, fers_core/supports/supportcondition.py:25-32):
active-set iteration in the oracle (CPU) and on the GPU path, against beam theory."""
import copy

import numpy as np
import pytest

from fers_core_b200 import hostlogic, wire
from oracle import fers_oracle
from tests import helpers

E, IZ = 210e9, 1.321e-5
P, L, A_SUP = 1000.0, 8.0, 4.0


def _doc(kind="PositiveOnly", stiffness=None):
    """Clamped at x = 0, unilateral vertical support at x = 4, tip load at x = 8: pushing down closes the support
    (propped cantilever), pulling up opens it (plain cantilever)."""
    doc = copy.deepcopy(helpers.golden("cantilever_end_load")["doc"])
    free = {"condition_type": "Free", "stiffness": None}
    m = doc["model"]
    m["nodes"] = [{"id": 1, "X": 0, "Y": 0, "Z": 0, "nodal_support": 1}, {"id": 2, "X": A_SUP, "Y": 0, "Z": 0, "nodal_support": 2},
                  {"id": 3, "X": L, "Y": 0, "Z": 0, "nodal_support": None}]
    m["nodal_supports"].append(dict(id=2, classification=None, warping_condition=None, X=free, Z=free, RX=free, RY=free, RZ=free,
                                    Y={"condition_type": kind, "stiffness": stiffness}))
    base = m["members"][0]
    m["members"] = [dict(base, id=1, start_node_id=1, end_node_id=2), dict(base, id=2, start_node_id=2, end_node_id=3)]
    m["member_sets"] = [{"id": 1, "classification": None, "member_ids": [1, 2], "buckling_restraints": []}]
    a = doc["analysis"]
    a["options"]["order"] = "LINEAR"
    lc = a["load_cases"][0]
    a["load_cases"] = [dict(lc, id=1, name="Down", nodal_loads=[{"id": 1, "node": 3, "magnitude": P, "direction": [0, -1, 0], "load_type": "force"}]),
                       dict(lc, id=2, name="Up", nodal_loads=[{"id": 2, "node": 3, "magnitude": P, "direction": [0, 1, 0], "load_type": "force"}])]
    return doc


def _expected(kind):
    ei = E * IZ
    r_closed = P * (3 * L - A_SUP) / (2 * A_SUP)                              # reaction of the closed support
    tip_closed = P * L ** 3 / (3 * ei) - r_closed * A_SUP ** 2 * (3 * L - A_SUP) / (6 * ei)
    tip_open = P * L ** 3 / (3 * ei)
    if kind == "PositiveOnly":      # pushes up only: closed under the downward load
        return {"Down": (-tip_closed, r_closed), "Up": (tip_open, 0.0)}
    return {"Down": (-tip_open, 0.0), "Up": (tip_closed, -r_closed)}


def test_active_set_rules():
    m = wire.parse_document(_doc("NegativeOnly", stiffness=5e6))
    idx, sign, k = hostlogic.unilateral_dofs(m)
    assert idx.tolist() == [7] and sign.tolist() == [-1.0] and k.tolist() == [5e6]
    code, kk = hostlogic.effective_bc(m, np.array([True]))
    assert code[1, 1] == wire.BC_SPRING and kk[1, 1] == 5e6
    code, kk = hostlogic.effective_bc(m, np.array([False]))
    assert code[1, 1] == wire.BC_FREE and kk[1, 1] == 0.0
    ideal = wire.parse_document(_doc("PositiveOnly"))
    assert hostlogic.effective_bc(ideal, np.array([True]))[0][1, 1] == wire.BC_FIXED
    u, r = np.zeros(18), np.zeros(18)
    r[7] = -5.0                                        # the closed support would have to pull: it opens
    assert hostlogic.unilateral_update(ideal, np.array([True]), u, r).tolist() == [False]
    r[7] = 5.0
    assert hostlogic.unilateral_update(ideal, np.array([True]), u, r).tolist() == [True]
    u[7] = -1e-3                                       # the open support is penetrated: it closes
    assert hostlogic.unilateral_update(ideal, np.array([False]), u, r).tolist() == [True]
    u[7] = 1e-3
    assert hostlogic.unilateral_update(ideal, np.array([False]), u, r).tolist() == [False]


@pytest.mark.parametrize("kind", ["PositiveOnly", "NegativeOnly"])
def test_oracle_matches_beam_theory(kind):
    m = wire.parse_document(_doc(kind))
    res = fers_oracle.analyze(m)
    for col, name in enumerate(res.plan.names):
        tip, reac = _expected(kind)[name]
        assert res.U[col, 2, 1] == pytest.approx(tip, rel=1e-9)
        assert res.R[col, 1, 1] == pytest.approx(reac, rel=1e-9, abs=1e-6)


def test_oracle_unilateral_spring():
    k = 2.0e6
    m = wire.parse_document(_doc("PositiveOnly", stiffness=k))
    res = fers_oracle.analyze(m)
    down, up = res.plan.names.index("Down"), res.plan.names.index("Up")
    assert res.U[down, 1, 1] < 0 and res.R[down, 1, 1] == pytest.approx(-k * res.U[down, 1, 1], rel=1e-9)
    assert res.U[up, 1, 1] > 0 and abs(res.R[up, 1, 1]) < 1e-6
    assert res.U[up, 2, 1] == pytest.approx(P * L ** 3 / (3 * E * IZ), rel=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,stiffness", [("PositiveOnly", None), ("NegativeOnly", None), ("PositiveOnly", 2.0e6)])
def test_gpu_matches_oracle(kind, stiffness):
    from fers_core_b200 import engine
    doc = _doc(kind, stiffness)
    want = fers_oracle.analyze(wire.parse_document(doc))
    got = engine.analyze(wire.parse_document(doc))
    helpers.assert_results_close(got, want, 1e-9, label=f"{kind} {stiffness}")
    if stiffness is None:
        for col, name in enumerate(got.plan.names):
            tip, reac = _expected(kind)[name]
            assert got.U[col, 2, 1] == pytest.approx(tip, rel=1e-9)
            assert got.R[col, 1, 1] == pytest.approx(reac, rel=1e-9, abs=1e-6)
