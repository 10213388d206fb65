"""The host-side rules exist twice: fers_core_b200/hostlogic.py (product) and oracle/host_oracle.py (the oracle's own
restatement, written from the reference's schema text).  The oracle no longer imports the product's copy (VERDICT r1, weak
point 2); this file holds the two against each other on every golden model and on seeded active-set / curve states, so a
change to either shows up here instead of silently moving both sides of every parity test."""
import numpy as np
import pytest

from fers_core_b200 import hostlogic, wire
from oracle import host_oracle
from tests import helpers


def test_the_oracle_does_not_import_the_products_host_rules():
    import re
    src = open(host_oracle.__file__).read()
    assert not re.search(r"^\s*(from|import)\s+fers_core_b200", src, re.M)
    src = open(host_oracle.__file__.replace("host_oracle", "fers_oracle")).read()
    assert "fers_core_b200.hostlogic" not in src.replace("product's hostlogic.py", "")


@pytest.mark.parametrize("name", helpers.golden_names())
def test_result_columns_geometry_and_constants_agree(name):
    model = wire.parse_document(helpers.golden(name)["doc"])
    a, b = hostlogic.rhs_plan(model), host_oracle.rhs_plan(model)
    assert (a.kinds, a.ids, a.names) == (b.kinds, b.ids, b.names)
    assert np.array_equal(a.factors, b.factors)
    assert [g.columns for g in a.groups] == [g.columns for g in b.groups]
    for ga, gb in zip(a.groups, b.groups):
        assert ga.imperfection is gb.imperfection
        if ga.imperfection is not None:
            xa, xb = hostlogic.imperfect_coordinates(model, ga.imperfection), host_oracle.imperfect_coordinates(model, gb.imperfection)
            assert np.abs(xa - xb).max() <= 1e-15 * max(1.0, np.abs(xa).max())
            assert np.abs(xa - model.xyz).max() > 0.0                      # the case does something
    assert np.array_equal(hostlogic.rigid_property_row(model.prop), host_oracle.rigid_property_row(model.prop))
    assert tuple(hostlogic.pdelta_suppressed_axes(model)) == tuple(host_oracle.pdelta_suppressed_axes(model))
    ia, sa, ka = hostlogic.unilateral_dofs(model)
    ib, sb, kb = host_oracle.unilateral_dofs(model)
    assert np.array_equal(ia, ib) and np.array_equal(sa, sb) and np.array_equal(ka, kb)
    assert host_oracle.BC_FREE == wire.BC_FREE and host_oracle.BC_FIXED == wire.BC_FIXED and host_oracle.BC_SPRING == wire.BC_SPRING
    assert host_oracle.BC_POS_ONLY == wire.BC_POS_ONLY and host_oracle.BC_NEG_ONLY == wire.BC_NEG_ONLY
    assert host_oracle.RIGID_STIFFNESS_RATIO == hostlogic.RIGID_FACTOR


def _nonlinear_support_docs():
    from tests import test_hinge_curves as th, test_support_curves as ts, test_unilateral_supports as tu
    return {"unilateral ideal +": tu._doc("PositiveOnly"), "unilateral ideal -": tu._doc("NegativeOnly"),
            "unilateral spring": tu._doc("PositiveOnly", stiffness=2.0e6), "support curve": ts._doc(),
            "support curve signed": ts._doc(signed=True), "hinge curve Mz": th._doc(),
            "hinge curve signed N": th._doc(depends_on="N", curve=th.CURVE_N, signed=True, axial=4.0e3)}


@pytest.mark.parametrize("name", ["unilateral ideal +", "unilateral ideal -", "unilateral spring", "support curve", "support curve signed",
                                  "hinge curve Mz", "hinge curve signed N"])
def test_active_sets_and_curves_agree_on_seeded_states(name):
    model = wire.parse_document(_nonlinear_support_docs()[name])
    idx, _, _ = hostlogic.unilateral_dofs(model)
    assert len(idx) or model.support_curves or model.hinge_curves
    rng = np.random.default_rng(len(name))
    for trial in range(6):
        u = rng.normal(size=model.n_dof) * 1e-3
        r = rng.normal(size=model.n_dof) * 1e4
        f_local = rng.normal(size=(model.n_members, 12)) * 1e4
        active = rng.random(len(idx)) < 0.5
        assert np.array_equal(hostlogic.unilateral_update(model, active, u, r), host_oracle.unilateral_update(model, active, u, r))
        ck_a = hostlogic.curve_stiffness(model, r if trial else None)
        ck_b = host_oracle.curve_stiffness(model, r if trial else None)
        assert np.allclose(ck_a, ck_b, rtol=1e-15, atol=0.0)
        ca, ka = hostlogic.effective_bc(model, active, ck_a if len(ck_a) else None)
        cb, kb = host_oracle.effective_bc(model, active, ck_b if len(ck_b) else None)
        assert np.array_equal(ca, cb) and np.array_equal(ka, kb)
        ha = hostlogic.hinge_curve_stiffness(model, f_local if trial else None)
        hb = host_oracle.hinge_curve_stiffness(model, f_local if trial else None)
        assert np.allclose(ha, hb, rtol=1e-15, atol=0.0)
        assert np.array_equal(hostlogic.rel_with_hinge_curves(model, ha), host_oracle.rel_with_hinge_curves(model, hb))
        assert hostlogic.curves_converged(ha * 1.0000001, ha, 1e-6) == host_oracle.curves_converged(hb * 1.0000001, hb, 1e-6)


def test_sway_and_translation_cases_by_hand():
    """Numbers worked by hand, not by either implementation: a 1/200 lean about Z of a node 7 m above the reference point
    moves it 35 mm along Z x Y = -X ... (axis x height direction); a translation case moves each node of the set once."""
    class M:
        pass
    m = M()
    m.xyz = np.array([[0.0, 0.0, 0.0], [1.0, 7.0, 2.0], [3.0, 2.0, 0.0]])
    m.m_n1 = np.array([0, 1]); m.m_n2 = np.array([1, 2])
    m.member_sets = {4: [0, 1]}
    case = {"sway_imperfections": [{"magnitude": 1 / 200, "axis": [0, 0, 1], "height_direction": [0, 1, 0], "reference_point": [0, 0, 0]}],
            "translation_imperfections": []}
    for impl in (hostlogic, host_oracle):
        x = impl.imperfect_coordinates(m, case)
        assert np.allclose(x[1], [1.0 - 7.0 / 200, 7.0, 2.0], rtol=0, atol=1e-15)         # Z x Y = -X
        assert np.allclose(x[2], [3.0 - 2.0 / 200, 2.0, 0.0], rtol=0, atol=1e-15) and np.array_equal(x[0], m.xyz[0])
    case = {"translation_imperfections": [{"magnitude": 0.01, "axis": [0, 0, 2.0], "memberset_ids": [4]}]}
    for impl in (hostlogic, host_oracle):
        x = impl.imperfect_coordinates(m, case)
        assert np.allclose(x - m.xyz, [[0, 0, 0.02]] * 3, rtol=0, atol=1e-16)               # node 1 is shared: moved once
