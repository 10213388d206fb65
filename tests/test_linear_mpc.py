"""RigidStrategy.LinearMpc (pydantic_models.py:1047-1049).  The product solves rigid members as stiff members under either
strategy and says so (W_LINEAR_MPC_AS_RIGID_MEMBER).  This file measures what that choice costs: the oracle also carries an
exact-constraint solve (master-slave elimination u = T u_master on the structure without its links, link forces from nodal
equilibrium; fers_oracle.analyze(exact_mpc=True)), pinned on the closed forms of the reference's own rigid-link cantilever
(tests/functionality/test_functionality.py:24-68), and the stiff-member results -- oracle and GPU -- are held against it.
(One augmented-Lagrangian step on the stiff-member solve was tried as a cheap way to exact kinematics on the GPU: it stalls at
5e-7, the rounding noise of a factorisation whose condition number carries the 1e6 penalty.)"""
import copy

import numpy as np
import pytest

from fers_core_b200 import wire
from oracle import fers_oracle
from tests import helpers

P, A = 1000.0, 5.0


def _doc(strategy="LinearMpc", load_node=3, order="LINEAR"):
    """node 1 (fixed) --elastic 5 m-- node 2 --rigid 5 m-- node 3 --rigid 2 m up-- node 4; load at `load_node`."""
    doc = copy.deepcopy(helpers.golden("rigid_link")["doc"])
    m, a = doc["model"], doc["analysis"]
    assert [n["id"] for n in m["nodes"]] == [1, 2, 3]
    m["nodes"].append(dict(m["nodes"][2], id=4, X=m["nodes"][2]["X"], Y=2.0, nodal_support=None))
    rigid = [mm for mm in m["members"] if str(mm.get("member_type", "")).lower() == "rigid"][0]
    m["members"].append(dict(rigid, id=max(mm["id"] for mm in m["members"]) + 1, start_node_id=3, end_node_id=4))
    for ms in m["member_sets"]:
        ms["member_ids"] = [mm["id"] for mm in m["members"]]
    a["options"].update(rigid_strategy=strategy, order=order, solve_loadcases=True)
    lc = a["load_cases"][0]
    lc["nodal_loads"] = [{"id": 1, "node": load_node, "magnitude": P, "direction": [0, -1, 0], "load_type": "force"},
                         {"id": 2, "node": 4, "magnitude": 300.0, "direction": [0, 0, 1], "load_type": "force"}]
    return doc


def test_exact_constraint_solve_closed_forms():
    model = wire.parse_document(_doc())
    assert "W_LINEAR_MPC_AS_RIGID_MEMBER" in [w[0] for w in model.warnings]      # what the product will do, said up front
    res = fers_oracle.analyze(model, exact_mpc=True)
    u = res.U[0]
    L = 5.0
    # node 3 and node 4 ride on node 2: same rotations, translations by theta x r
    for slave, r in ((2, np.array([A, 0.0, 0.0])), (3, np.array([A, 2.0, 0.0]))):
        assert np.abs(u[slave, 3:] - u[1, 3:]).max() <= 1e-13 * np.abs(u[1, 3:]).max()
        want = u[1, :3] + np.cross(u[1, 3:], r)
        assert np.abs(u[slave, :3] - want).max() <= 1e-12 * np.abs(u[:, :3]).max()
    # elastic span: cantilever with the tip load P and the tip moment -P a (load at node 3) in the XY plane
    EI = float(model.prop[0, wire.P_EIZ])
    assert u[1, 1] == pytest.approx(-(P * L ** 3 / (3 * EI) + P * A * L ** 2 / (2 * EI)), rel=1e-9)
    assert u[1, 5] == pytest.approx(-(P * L ** 2 / (2 * EI) + P * A * L / EI), rel=1e-9)
    # reaction at the fixed node: P up, moment P (L + a)
    assert res.R[0, 0, 1] == pytest.approx(P, rel=1e-9) and res.R[0, 0, 5] == pytest.approx(P * (L + A), rel=1e-9)
    # the link 2 -> 3 in the stiffness sign: node 2 holds it up (+P), node 3 loads it (-P); moment P a at the near end
    rigid_rows = np.nonzero(model.m_type == wire.MT_RIGID)[0]
    link = [m for m in rigid_rows if model.m_n1[m] == 1 and model.m_n2[m] == 2][0]
    fg = res.f_global[0, link]
    assert fg[1] == pytest.approx(P, rel=1e-9) and fg[7] == pytest.approx(-P, rel=1e-9)
    assert abs(fg[5]) == pytest.approx(P * A, rel=1e-9) and abs(fg[11]) <= 1e-6


@pytest.mark.parametrize("load_node", [2, 3, 4])
def test_stiff_member_error_against_the_exact_constraint(load_node):
    model = wire.parse_document(_doc(load_node=load_node))
    exact = fers_oracle.analyze(model, exact_mpc=True)
    stiff = fers_oracle.analyze(model)                                            # what the product mirrors
    for name in ("U", "R", "f_local", "f_global", "sec"):
        x, y = getattr(stiff, name), getattr(exact, name)
        err = np.abs(x - y).max() / np.abs(y).max()
        assert err <= 1e-5, (name, err)                                           # the stiff member is good to ~2e-6 here ...
    rot = np.abs(stiff.U[0, 2, 3:] - stiff.U[0, 1, 3:]).max() / np.abs(stiff.U[0, 1, 3:]).max()
    assert 1e-12 < rot < 1e-5                                                     # ... and is not exactly rigid


@pytest.mark.gpu
def test_gpu_stiff_member_error_against_the_exact_constraint():
    from fers_core_b200 import engine
    model = wire.parse_document(_doc())
    exact = fers_oracle.analyze(model, exact_mpc=True)
    got = engine.analyze(model)
    for name in ("U", "R", "f_local", "sec"):
        x, y = getattr(got, name), getattr(exact, name)
        assert np.abs(x - y).max() <= 1e-5 * np.abs(y).max(), name
    assert "W_LINEAR_MPC_AS_RIGID_MEMBER" in [w[0] for w in list(model.warnings) + list(got.warnings)]
