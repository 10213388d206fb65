"""RigidStrategy.LinearMpc (pydantic_models.py:1047-1049): rigid members as exact kinematic constraints in first-order
analyses.  Two formulations that must meet: the oracle eliminates the slave nodes with a transformation matrix (u = T u_master
on the structure without its links) and takes the link forces from nodal equilibrium; the product re-attaches the elastic
members of a slave to the master through the rigid arms of the end-offset records (hostlogic.mpc_transform), solves that model
on the GPU and recovers slave displacements and link forces on the device (engine._MpcDevice).  Closed forms: the reference's
own rigid-link cantilever (tests/functionality/test_functionality.py:24-68).  Where the transformation does not apply (second
order, supports or plates on slaves, ...) both fall back to stiff members and say so; the stiff member is measured against
the exact constraint here (2e-6 on this model).  (One augmented-Lagrangian step on the stiff-member solve was tried first as a
cheap way to exact kinematics: it stalls at 5e-7, the rounding noise of a factorisation whose condition number carries the
1e6 penalty.)"""
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
    res = fers_oracle.analyze(model)
    assert not [w for w in list(model.warnings) + list(res.warnings) if "MPC" in w[0]]
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
    exact = fers_oracle.analyze(model)
    stiff = fers_oracle.analyze(model, exact_mpc=False)                           # RigidMember strategy
    for name in ("U", "R", "f_local", "f_global", "sec"):
        x, y = getattr(stiff, name), getattr(exact, name)
        err = np.abs(x - y).max() / np.abs(y).max()
        assert err <= 1e-5, (name, err)                                           # the stiff member is good to ~2e-6 here ...
    rot = np.abs(stiff.U[0, 2, 3:] - stiff.U[0, 1, 3:]).max() / np.abs(stiff.U[0, 1, 3:]).max()
    assert 1e-12 < rot < 1e-5                                                     # ... and is not exactly rigid


def test_conditions_agree_between_product_and_oracle_and_fallback_warns():
    from fers_core_b200 import hostlogic
    for kw, applies in ((dict(), True), (dict(order="NONLINEAR"), False), (dict(strategy="RigidMember"), False)):
        model = wire.parse_document(_doc(**kw))
        mt, info = hostlogic.mpc_transform(model)
        wants = model.options.rigid_strategy == "LinearMpc"
        assert (mt is not None) == applies
        assert (fers_oracle.mpc_applies(model) and wants) == applies
        res = fers_oracle.analyze(model)
        assert ("W_LINEAR_MPC_AS_RIGID_MEMBER" in [w[0] for w in res.warnings]) == (wants and not applies)
    # a support on a slave node: both decline
    doc = _doc()
    doc["model"]["nodes"][2]["nodal_support"] = 1       # two supported nodes in one rigid cluster: one of them would be a slave
    doc["model"]["nodes"][3]["nodal_support"] = 1
    model = wire.parse_document(doc)
    assert hostlogic.mpc_transform(model)[0] is None and not fers_oracle.mpc_applies(model)


def test_transformed_model_is_the_eliminated_system():
    """The product's transformed model (members re-attached with arms, loads moved) assembled by the oracle's assembly gives
    the oracle's own T^T K0 T and T^T f on the master DOFs."""
    from fers_core_b200 import hostlogic
    from tests.test_random_models import random_model
    model = _rigid_frame()
    mt, info = hostlogic.mpc_transform(model)
    assert mt is not None and len(info.att_member) > 0
    want = fers_oracle.analyze(model)
    got = fers_oracle.analyze(mt, exact_mpc=False)            # the transformed model has no rigid members left: plain solve
    keep = np.ones(model.n_nodes, bool); keep[info.slaves] = False
    assert np.abs(got.U[:, keep] - want.U[:, keep]).max() <= 1e-9 * np.abs(want.U).max()
    assert np.abs(got.R - want.R).max() <= 1e-9 * np.abs(want.R).max()
    # members that did not sit on a slave report the same end forces; those that did report them at the master (the engine
    # carries them back to the slave node: engine._MpcDevice)
    untouched = model.m_type != wire.MT_RIGID
    untouched[info.att_member] = False
    assert np.abs(got.f_local[:, untouched] - want.f_local[:, untouched]).max() <= 1e-8 * np.abs(want.f_local).max()
    assert np.abs(got.sec[:, model.m_type != wire.MT_RIGID] - want.sec[:, model.m_type != wire.MT_RIGID]).max() <= 1e-8 * np.abs(want.sec).max()


def _rigid_frame():
    from tests.test_random_models import random_model
    model = random_model(4)                      # linear, no plates
    rng = np.random.default_rng(11)
    free_both = (model.node_support_id[model.m_n1] < 0) & (model.node_support_id[model.m_n2] < 0) & (model.m_rel < 0)
    cand = np.nonzero(free_both & (model.m_type == wire.MT_NORMAL))[0]
    used, rigid = set(), []
    for m in rng.permutation(cand)[:6]:           # disjoint node pairs: independent two-node clusters
        a, b = int(model.m_n1[m]), int(model.m_n2[m])
        if a in used or b in used:
            continue
        used.update((a, b)); rigid.append(int(m))
    model.m_type = model.m_type.copy(); model.m_type[rigid] = wire.MT_RIGID
    for lc in model.load_cases:                   # rigid members carry no member loads
        keep = ~np.isin(lc.dl_member, rigid)
        lc.dl_member, lc.dl_data = lc.dl_member[keep], lc.dl_data[keep]
    model.options.rigid_strategy = "LinearMpc"
    return model


@pytest.mark.gpu
@pytest.mark.parametrize("load_node", [3, 4, 2])
def test_gpu_matches_oracle(load_node):
    import json
    import fers_calculations
    from fers_core_b200 import engine
    doc = _doc(load_node=load_node)
    model = wire.parse_document(doc)
    want, got = fers_oracle.analyze(model), engine.analyze(model)
    assert not [w for w in got.warnings if "MPC" in w[0]]
    for name in ("U", "R", "f_local", "f_global", "sec", "lmax", "lmin"):
        x, y = getattr(got, name), getattr(want, name)
        assert np.abs(x - y).max() <= 1e-9 * np.abs(y).max(), (name, np.abs(x - y).max() / np.abs(y).max())
    u = got.U[0]
    assert np.abs(u[2, 3:] - u[1, 3:]).max() == 0.0                                       # exact rigid kinematics
    out = json.loads(fers_calculations.calculate_from_json(json.dumps(doc)))["results"]
    assert "End Load" in out["loadcases"]


@pytest.mark.gpu
def test_gpu_random_frame_with_rigid_beams():
    from fers_core_b200 import engine
    model = _rigid_frame()
    want, got = fers_oracle.analyze(model), engine.analyze(model)
    assert not [w for w in got.warnings if "MPC" in w[0]]
    for name in ("U", "R", "f_local", "f_global", "sec"):
        x, y = getattr(got, name), getattr(want, name)
        assert np.abs(x - y).max() <= 1e-8 * np.abs(y).max(), (name, np.abs(x - y).max() / np.abs(y).max())


@pytest.mark.gpu
def test_gpu_second_order_falls_back_to_stiff_members():
    from fers_core_b200 import engine
    model = wire.parse_document(_doc(order="NONLINEAR"))
    want, got = fers_oracle.analyze(model), engine.analyze(model)
    assert "W_LINEAR_MPC_AS_RIGID_MEMBER" in [w[0] for w in got.warnings]
    # two factorisations of a matrix whose condition number carries the 1e6 penalty agree to ~1e-5, not to 1e-9: the price of
    # the stiff member, and the reason the first-order path uses the exact constraint
    assert np.abs(got.U - want.U).max() <= 1e-4 * np.abs(want.U).max()
