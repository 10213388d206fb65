"""Load-dependent hinge springs (MemberHinge.stiffness_curve_*, pydantic_models.py:2062-2087; MemberStiffnessCurve
1638-1652, drivers N / Vy / Vz / Mx / My / Mz 583-589): the member-side twin of the support curves.  Known answer: a
cantilever of two members with a rotational spring in the joint between them; the joint moment M = P b is statically
determinate, so the first-order tip deflection is  delta = P L^3 / 3EI + P b^2 / k(driver)  for any curve."""
import copy

import numpy as np
import pytest

from fers_core_b200 import hostlogic, wire
from oracle import fers_oracle as fo
from tests import helpers

A_LEN, B_LEN = 2.5, 2.5
E_MOD, I_Z = 210e9, 13.21e-6
LOADS = [250.0, 1000.0, 1600.0, 3000.0]                      # tip load P (N), one load case each
CURVE_M = [[0.0, 25_000.0], [2_000.0, 25_000.0], [5_000.0, 70_000.0]]             # k(|Mz|): softens nothing, stiffens above 2 kNm
CURVE_N = [[-20_000.0, 10_000.0], [0.0, 30_000.0], [20_000.0, 90_000.0]]          # k(N), signed: tension stiffens


@pytest.fixture(scope="module")
def engine_mod():
    from fers_core_b200 import engine
    return engine


def _k(curve, f):
    pts = np.array(curve)
    return float(np.interp(f, pts[:, 0], pts[:, 1]))


def _doc(where="end_of_a", depends_on="Mz", curve=CURVE_M, signed=None, axial=0.0, order="LINEAR", constant=None):
    doc = copy.deepcopy(helpers.golden("hinge_spring")["doc"])
    m = doc["model"]
    hinge = m["member_hinges"][0]
    hinge["rotational_release_mz"] = constant
    hinge["stiffness_curve_mz"] = None if constant is not None else {"depends_on": depends_on, "points": curve, "signed": signed}
    a, b = m["members"]
    a.update(member_type="Normal", section=b["section"], start_hinge=None, end_hinge=1 if where == "end_of_a" else None)
    b.update(start_hinge=1 if where == "start_of_b" else None, end_hinge=None)
    an = doc["analysis"]
    an["options"]["order"] = order
    if order != "LINEAR":
        an["options"].update(nonlinear_method="P_DELTA", pdelta_formulation="CONSISTENT", tolerance=1e-10, max_iterations=50)
    lc0 = an["load_cases"][0]
    cases = []
    for i, p in enumerate(LOADS):
        loads = [{"id": 2 * i + 1, "node": 3, "magnitude": p, "direction": [0, -1, 0], "load_type": "force"}]
        if axial:
            loads.append({"id": 2 * i + 2, "node": 3, "magnitude": axial, "direction": [1, 0, 0], "load_type": "force"})
        cases.append(dict(lc0, id=i + 1, name=f"P={p:g}", nodal_loads=loads))
    an["load_cases"] = cases
    an["load_combinations"] = []
    return doc


def _closed_form(p, k):
    L = A_LEN + B_LEN
    return -(p * L ** 3 / (3 * E_MOD * I_Z) + p * B_LEN ** 2 / k)


def test_parser_builds_the_curve_records():
    m = wire.parse_document(_doc())
    assert len(m.hinge_curves) == 1
    row, rel_row, slot, comp, pts, signed = m.hinge_curves[0]
    assert (row, slot, comp, signed) == (0, 11, 5, False) and pts.shape == (3, 2)
    assert m.rel[rel_row, slot] == 25_000.0                              # starts at the zero-force stiffness
    m = wire.parse_document(_doc(where="start_of_b", depends_on="N", curve=CURVE_N, signed=True))
    row, rel_row, slot, comp, pts, signed = m.hinge_curves[0]
    assert (row, slot, comp, signed) == (1, 5, 0, True) and m.rel[rel_row, slot] == 30_000.0
    f = np.zeros((2, 12)); f[1, 0] = -5_000.0; f[1, 6] = 5_000.0         # 5 kN of tension in member b
    assert hostlogic.hinge_curve_stiffness(m, f)[0] == pytest.approx(_k(CURVE_N, 5_000.0))
    assert hostlogic.hinge_curve_stiffness(m, -f)[0] == pytest.approx(_k(CURVE_N, -5_000.0))
    assert hostlogic.rel_with_hinge_curves(m, np.array([123.0]))[rel_row, slot] == 123.0 and m.rel[rel_row, slot] == 30_000.0
    for bad in ({"depends_on": "Fz"}, {"points": [[0, 1.0]]}, {"points": [[1, 1.0], [0, 1.0]]}, {"points": [[0, 1.0], [1, -1.0]]}):
        doc = _doc()
        doc["model"]["member_hinges"][0]["stiffness_curve_mz"].update(bad)
        with pytest.raises(wire.WireError):
            wire.parse_document(doc)


@pytest.mark.parametrize("where", ["end_of_a", "start_of_b"])
def test_oracle_moment_driven_spring_closed_form(where):
    res = fo.analyze(wire.parse_document(_doc(where=where)))
    for col, p in enumerate(LOADS):
        assert res.U[col, 2, 1] == pytest.approx(_closed_form(p, _k(CURVE_M, p * B_LEN)), rel=1e-9), p
    # below 2 kNm the curve is flat: same answer as the constant spring
    flat = fo.analyze(wire.parse_document(_doc(where=where, constant=25_000.0)))
    assert res.U[0, 2, 1] == pytest.approx(flat.U[0, 2, 1], rel=1e-12)
    assert abs(res.U[3, 2, 1]) < abs(flat.U[3, 2, 1])


@pytest.mark.parametrize("axial", [8_000.0, -8_000.0])
def test_oracle_signed_axial_driver_tension_positive(axial):
    """First order: the axial load changes nothing but the spring; +X tip load = tension = the stiff side of the curve."""
    for where in ("end_of_a", "start_of_b"):
        res = fo.analyze(wire.parse_document(_doc(where=where, depends_on="N", curve=CURVE_N, signed=True, axial=axial)))
        for col, p in enumerate(LOADS):
            assert res.U[col, 2, 1] == pytest.approx(_closed_form(p, _k(CURVE_N, axial)), rel=1e-9)
    unsigned = fo.analyze(wire.parse_document(_doc(depends_on="N", curve=CURVE_N, signed=False, axial=axial)))
    assert unsigned.U[1, 2, 1] == pytest.approx(_closed_form(LOADS[1], _k(CURVE_N, abs(axial))), rel=1e-9)


def test_oracle_signed_moment_is_continuous_across_the_joint():
    """The same signed curve on the end of member a or on the start of member b sees the same internal moment."""
    curve = [[-10_000.0, 20_000.0], [10_000.0, 80_000.0]]
    a = fo.analyze(wire.parse_document(_doc(where="end_of_a", curve=curve, signed=True)))
    b = fo.analyze(wire.parse_document(_doc(where="start_of_b", curve=curve, signed=True)))
    # (the joint node's own rotation follows member b in one model and member a in the other)
    np.testing.assert_allclose(a.U[:, :, :3], b.U[:, :, :3], rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(a.U[:, 2], b.U[:, 2], rtol=1e-9, atol=1e-15)
    k = [(-LOADS[c] * B_LEN ** 2 / (a.U[c, 2, 1] + LOADS[c] * (A_LEN + B_LEN) ** 3 / (3 * E_MOD * I_Z))) for c in range(len(LOADS))]
    m = LOADS[1] * B_LEN
    assert k[1] == pytest.approx(_k(curve, m), rel=1e-6) or k[1] == pytest.approx(_k(curve, -m), rel=1e-6)


def test_oracle_second_order_converges_and_amplifies():
    lin = fo.analyze(wire.parse_document(_doc(axial=-2_000.0)))
    sec = fo.analyze(wire.parse_document(_doc(axial=-2_000.0, order="NONLINEAR")))
    assert sec.converged.all()
    assert (np.abs(sec.U[:, 2, 1]) > np.abs(lin.U[:, 2, 1])).all()           # compression amplifies, and raises the joint moment


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["moment", "axial_signed", "second_order"])
def test_gpu_matches_oracle(engine_mod, case):
    doc = {"moment": _doc(where="start_of_b"),
           "axial_signed": _doc(depends_on="N", curve=CURVE_N, signed=True, axial=8_000.0),
           "second_order": _doc(axial=-2_000.0, order="NONLINEAR")}[case]
    want = fo.analyze(wire.parse_document(doc))
    got = engine_mod.analyze(wire.parse_document(doc))
    helpers.assert_results_close(got, want, 1e-9 if case != "second_order" else 1e-8, label="hinge curve " + case)
    if case == "moment":
        for col, p in enumerate(LOADS):
            assert got.U[col, 2, 1] == pytest.approx(_closed_form(p, _k(CURVE_M, p * B_LEN)), rel=1e-9)


def test_hinge_capacities_are_checked_after_the_analysis():
    """MemberHinge.max_moment_* / max_tension_* (pydantic_models.py:2069-2074): no behaviour is stated for them, so the
    analysis does not enforce them -- it reports where the end forces exceed them.  The joint moment is M = P b
    (statically determinate): with a capacity of 3 kNm the load cases P = 1600 N (4.0 kNm) and 3000 N (7.5 kNm) exceed it,
    the governing one is named, and a capacity above the largest moment stays silent."""
    for cap, expect in ((3.0e3, True), (8.0e3, False)):
        doc = _doc(constant=40_000.0)
        doc["model"]["member_hinges"][0]["max_moment_mz"] = cap
        doc["model"]["member_hinges"][0]["max_tension_vx"] = 1.0          # the cantilever carries no axial force: never exceeded
        model = wire.parse_document(doc)
        assert [(c[1], c[2], c[3]) for c in model.hinge_caps] == [(1, 0, 1.0), (1, 5, cap)]
        res = fo.analyze(model)
        notes = [msg for code, msg in res.warnings if code == "W_HINGE_CAPACITY_EXCEEDED"]
        assert bool(notes) == expect
        if expect:
            assert len(notes) == 1 and "Mz = 7500" in notes[0] and f"utilisation {7500.0 / cap:.3f}" in notes[0]
        # the product states the rule a second time (hostlogic); both read the same end forces
        assert hostlogic.hinge_capacity_notes(model, res.f_local, fo.rhs_plan(model).names) == \
            [(c, msg) for c, msg in res.warnings if c == "W_HINGE_CAPACITY_EXCEEDED"]
        assert any(c == "W_HINGE_CAPACITY_IGNORED" for c, _ in model.warnings)


@pytest.mark.gpu
def test_gpu_reports_the_same_capacity_notes(engine_mod):
    doc = _doc(constant=40_000.0)
    doc["model"]["member_hinges"][0]["max_moment_mz"] = 3.0e3
    model = wire.parse_document(doc)
    got = [w for w in engine_mod.analyze(model).warnings if w[0] == "W_HINGE_CAPACITY_EXCEEDED"]
    want = [w for w in fo.analyze(model).warnings if w[0] == "W_HINGE_CAPACITY_EXCEEDED"]
    assert len(got) == 1 and got[0][1].split("=")[0] == want[0][1].split("=")[0] and "utilisation 2.500" in got[0][1]
