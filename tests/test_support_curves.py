"""Load-dependent support springs (SupportStiffnessCurve, pydantic_models.py:1249-1265; SupportCondition.spring_curve,
fers_core/supports/supportcondition.py:19-23).  Known answer: the reference's example 131
(fers_core/examples/131_StiffnessCurve_Cantilever.py): 5 m column, rotational base spring whose stiffness follows the
axial reaction Fz, 1 kN lateral load, first-order closed form  delta = F h^2 / k(N) + F h^3 / (3 E I)  (:141-146)."""
import copy

import numpy as np
import pytest

from fers_core_b200 import hostlogic, wire
from oracle import fers_oracle as fo
from tests import helpers

H, F_H = 5.0, 1000.0
AXIAL_KN = [0, 10, 20, 30, 40, 50, 60, 70, 80]
CURVE = [[-80_000, 6_250_000], [-60_000, 5_000_000], [-40_000, 3_750_000], [-20_000, 2_500_000], [0, 2_500_000],
         [20_000, 2_500_000], [40_000, 3_750_000], [60_000, 5_000_000], [80_000, 6_250_000]]     # example 131:52-66
E_MOD, I_BEND = 210e9, 5.696e-5


@pytest.fixture(scope="module")
def engine_mod():
    from fers_core_b200 import engine
    return engine


def _interpolate_k(n_kn):
    """example 131:126-139"""
    fz = n_kn * 1000.0
    if fz <= CURVE[0][0]:
        return CURVE[0][1]
    if fz >= CURVE[-1][0]:
        return CURVE[-1][1]
    for (f0, k0), (f1, k1) in zip(CURVE[:-1], CURVE[1:]):
        if f0 <= fz <= f1:
            return k0 + (fz - f0) / (f1 - f0) * (k1 - k0)
    return CURVE[-1][1]


def _doc(order="LINEAR", signed=None, curve=True):
    doc = copy.deepcopy(helpers.golden("cantilever_end_load")["doc"])
    m = doc["model"]
    m["nodes"] = [{"id": 1, "X": 0, "Y": 0, "Z": 0, "nodal_support": 1}, {"id": 2, "X": 0, "Y": 0, "Z": H, "nodal_support": None}]
    sec = m["sections"][0]
    sec.update(i_y=I_BEND, i_z=I_BEND, j=5.92e-7, area=7.81e-3)     # HEB 200; i_y = i_z: neutral on SURVEY A.7
    fixed = {"condition_type": "Fixed", "stiffness": None}
    ry = {"condition_type": "Spring", "stiffness": None,
          "stiffness_curve": {"depends_on": "Fz", "points": CURVE, "signed": signed}} if curve else \
        {"condition_type": "Spring", "stiffness": 2_500_000.0}
    m["nodal_supports"] = [dict(id=1, classification=None, warping_condition=None, X=fixed, Y=fixed, Z=fixed, RX=fixed, RZ=fixed, RY=ry)]
    a = doc["analysis"]
    a["options"]["order"] = order
    if order != "LINEAR":
        a["options"].update(nonlinear_method="P_DELTA", pdelta_formulation="CONSISTENT", tolerance=1e-9, max_iterations=50)
    lc0 = a["load_cases"][0]
    cases = []
    for i, n_kn in enumerate(AXIAL_KN):
        loads = [{"id": 2 * i + 1, "node": 2, "magnitude": F_H, "direction": [1, 0, 0], "load_type": "force"}]
        if n_kn > 0:
            loads.append({"id": 2 * i + 2, "node": 2, "magnitude": n_kn * 1000.0, "direction": [0, 0, -1], "load_type": "force"})
        cases.append(dict(lc0, id=i + 1, name=f"N={n_kn}kN", nodal_loads=loads))
    a["load_cases"] = cases
    a["load_combinations"] = []
    return doc


def test_parser_and_interpolation():
    m = wire.parse_document(_doc())
    assert len(m.support_curves) == 1
    dof, drv, pts, signed = m.support_curves[0]
    assert (dof, drv, signed) == (4, 2, False) and pts.shape == (9, 2)
    assert m.bc_code[0, 4] == wire.BC_SPRING and m.bc_k[0, 4] == 2_500_000.0
    assert not [w for w in m.warnings if w[0] == "W_SUPPORT_CURVE"]
    r = np.zeros(12)
    for n_kn in (0, 15, 30, 55, 80, 120):
        for sgn in (1.0, -1.0):
            r[2] = sgn * n_kn * 1000.0
            assert hostlogic.curve_stiffness(m, r)[0] == pytest.approx(_interpolate_k(n_kn), rel=1e-15)
    bad = _doc()
    bad["model"]["nodal_supports"][0]["RY"]["stiffness_curve"]["depends_on"] = "N"
    with pytest.raises(wire.WireError):
        wire.parse_document(bad)


def test_oracle_reproduces_example_131_first_order():
    res = fo.analyze(wire.parse_document(_doc()))
    for col, n_kn in enumerate(AXIAL_KN):
        want = F_H * H ** 2 / _interpolate_k(n_kn) + F_H * H ** 3 / (3 * E_MOD * I_BEND)
        assert res.U[col, 1, 0] == pytest.approx(want, rel=1e-9), n_kn
        assert res.R[col, 0, 4] == pytest.approx(-F_H * H, rel=1e-9)          # base moment carried by the spring
        assert res.R[col, 0, 2] == pytest.approx(n_kn * 1000.0, rel=1e-9, abs=1e-6)


def test_oracle_second_order_curve_is_stiffer_than_the_constant_spring():
    a = fo.analyze(wire.parse_document(_doc(order="NONLINEAR")))
    b = fo.analyze(wire.parse_document(_doc(order="NONLINEAR", curve=False)))
    assert a.converged.all() and b.converged.all()
    for col, n_kn in enumerate(AXIAL_KN):
        if n_kn <= 20:
            assert a.U[col, 1, 0] == pytest.approx(b.U[col, 1, 0], rel=1e-9)
        else:
            assert a.U[col, 1, 0] < b.U[col, 1, 0]
    assert a.U[-1, 1, 0] > F_H * H ** 2 / 6_250_000 + F_H * H ** 3 / (3 * E_MOD * I_BEND)    # P-delta amplifies


@pytest.mark.gpu
@pytest.mark.parametrize("order", ["LINEAR", "NONLINEAR"])
def test_gpu_matches_oracle(engine_mod, order):
    doc = _doc(order=order)
    want = fo.analyze(wire.parse_document(doc))
    got = engine_mod.analyze(wire.parse_document(doc))
    helpers.assert_results_close(got, want, 1e-9 if order == "LINEAR" else 1e-8, label="support curve")
    if order == "LINEAR":
        for col, n_kn in enumerate(AXIAL_KN):
            want_d = F_H * H ** 2 / _interpolate_k(n_kn) + F_H * H ** 3 / (3 * E_MOD * I_BEND)
            assert got.U[col, 1, 0] == pytest.approx(want_d, rel=1e-9)
