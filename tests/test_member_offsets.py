"""Member end offsets (MemberEndOffset, pydantic_models.py:525-531; Member.start_offset 1541, end_offset 1518): the
elastic axis runs between (node + offset) points, each tied to its node by a rigid arm.  Wire parser, oracle against
beam theory, host instantiation of the element code against the oracle, GPU path against the oracle."""
import copy
import ctypes as C

import numpy as np
import pytest

from fers_core_b200 import wire
from oracle import fers_oracle as fo
from tests import helpers
from tests.test_elem_host import CASES, _p, _rel

E_MOD, IZ, AREA = 210e9, 1.321e-5, 0.00196
P_TIP, L_NODES = 1000.0, 5.0


@pytest.fixture(scope="module")
def engine_mod():
    from fers_core_b200 import engine
    return engine


def _doc(start=None, end=None, load_dir=(0, -1, 0)):
    doc = copy.deepcopy(helpers.golden("cantilever_end_load")["doc"])
    m = doc["model"]["members"][0]
    if start is not None:
        m["start_offset"] = dict(zip("XYZ", start))
    if end is not None:
        m["end_offset"] = dict(zip("XYZ", end))
    doc["analysis"]["options"]["order"] = "LINEAR"
    doc["analysis"]["load_cases"][0]["nodal_loads"][0]["direction"] = list(load_dir)
    return doc


def test_parser_builds_offset_records():
    m = wire.parse_document(_doc(start=(0.5, 0.0, 0.0), end=(0.0, 0.1, -0.2)))
    assert m.rel.shape == (1, wire.REL_STRIDE) and m.m_rel.tolist() == [0]
    assert m.rel[0, wire.REL_O1:wire.REL_O1 + 6].tolist() == [0.5, 0.0, 0.0, 0.0, 0.1, -0.2] and m.rel[0, wire.REL_EIYZ] == 0.0
    assert (m.rel[0, :12] == wire.RIGID_RELEASE).all() and m.rel[0, 12:14].tolist() == [0.0, 0.0]
    assert not [w for w in m.warnings if "OFFSET" in w[0]]
    assert wire.parse_document(_doc()).m_rel.tolist() == [-1]
    with pytest.raises(wire.WireError):             # the offsets close the member axis to a point
        wire.parse_document(_doc(start=(L_NODES, 0.0, 0.0)))


def test_oracle_rigid_arm_along_the_axis():
    """Clamp at x = 0, rigid arm to x = 0.5, beam 0.5 .. 5: tip = P (4.5)^3 / 3 E I, clamp moment = P * 5."""
    res = fo.analyze(wire.parse_document(_doc(start=(0.5, 0.0, 0.0))))
    Le = L_NODES - 0.5
    assert res.U[0, 1, 1] == pytest.approx(-P_TIP * Le ** 3 / (3 * E_MOD * IZ), rel=1e-9)
    assert res.U[0, 1, 5] == pytest.approx(-P_TIP * Le ** 2 / (2 * E_MOD * IZ), rel=1e-9)
    assert res.R[0, 0, 1] == pytest.approx(P_TIP, rel=1e-9) and res.R[0, 0, 5] == pytest.approx(P_TIP * L_NODES, rel=1e-9)
    # section forces are stations of the ELASTIC axis: M(x) = P (Le - x)
    for s, x in enumerate((0.25, 0.5, 0.75)):
        assert res.sec[0, 0, s, 5] == pytest.approx(P_TIP * Le * (1.0 - x), rel=1e-9)


def test_oracle_eccentric_axial_load():
    """Both ends offset by e in +Y: the axial load P at the tip node acts e below the beam axis: the beam sees
    N = P and a constant moment Mz = -P e... the tip node rotates by rz = M L / E I and shortens by P L / E A + e rz."""
    e = 0.15
    res = fo.analyze(wire.parse_document(_doc(start=(0.0, e, 0.0), end=(0.0, e, 0.0), load_dir=(1, 0, 0))))
    Mz = P_TIP * e                       # force +X applied e below the axis: sagging rotation +rz
    rz = Mz * L_NODES / (E_MOD * IZ)
    assert res.U[0, 1, 5] == pytest.approx(rz, rel=1e-9)
    assert res.U[0, 1, 0] == pytest.approx(P_TIP * L_NODES / (E_MOD * AREA) + e * rz, rel=1e-9)
    assert res.U[0, 1, 1] == pytest.approx(Mz * L_NODES ** 2 / (2 * E_MOD * IZ), rel=1e-9)
    assert res.R[0, 0, 0] == pytest.approx(-P_TIP, rel=1e-9) and abs(res.R[0, 0, 5]) < 1e-6


@pytest.mark.parametrize("case", CASES[:8])
def test_element_code_matches_oracle(elem_host, case):
    a, b, roll, Pp, mtype = case
    rec = np.full(wire.REL_STRIDE, -1.0)
    rec[wire.REL_EIYZ] = 0.3 * np.sqrt(Pp[1] * Pp[2])      # product of inertia on top of everything else
    rec[4] = 0.0
    rec[11] = 7.0e3
    rec[12:14] = (0.02, -0.03)
    rec[14:20] = (0.3, -0.2, 0.1, -0.15, 0.05, 0.25)
    off = rec[14:20].reshape(1, 2, 3)
    xyz = np.stack([a, b])
    R, L = fo.member_frames(xyz, np.array([0]), np.array([1]), np.array([roll]), off)
    k = fo.local_stiffness(Pp[None, :], L, np.array([mtype]))[0]
    T = fo.expand_T(R)[0]
    p = np.arange(1.0, 13.0)[None, :] * 13.0
    kc, pc = fo.member_modifiers(k, p, rec, L[0])
    Em = fo.rigid_arm_matrix(R[0] @ off[0, 0], R[0] @ off[0, 1])
    want = T.T @ (Em.T @ kc @ Em) @ T
    out = np.zeros(144)
    elem_host.h_beam_ke(_p(a), _p(b), C.c_double(roll), _p(Pp), int(mtype), C.c_double(1.0), _p(rec), 1, _p(out))
    assert _rel(out.reshape(12, 12), want) < 1e-12
    # rigid-body motion of the two NODES costs nothing when nothing is released
    rec[:12] = -1.0
    elem_host.h_beam_ke(_p(a), _p(b), C.c_double(roll), _p(Pp), int(mtype), C.c_double(1.0), _p(rec), 1, _p(out))
    kg = out.reshape(12, 12)
    th = np.array([0.3, -0.2, 0.5])
    rb = np.concatenate([np.cross(th, a), th, np.cross(th, b), th])
    assert np.abs(kg @ rb).max() <= 1e-9 * np.abs(kg).max() * np.abs(rb).max()


def test_element_loads_match_oracle(elem_host):
    a, b, roll, Pp, _ = CASES[3]
    mtype = 0
    rec = np.full(wire.REL_STRIDE, -1.0)
    rec[wire.REL_EIYZ] = 0.0
    rec[5] = 0.0
    rec[12:14] = 0.0
    rec[14:20] = (0.3, -0.2, 0.1, -0.15, 0.05, 0.25)
    off = rec[14:20].reshape(1, 2, 3)
    R, L = fo.member_frames(np.stack([a, b]), np.array([0]), np.array([1]), np.array([roll]), off)
    k = fo.local_stiffness(Pp[None, :], L, np.array([mtype]))[0]
    dl = np.array([[0.0, -1200.0, 300.0, 0.0, -400.0, 100.0, 0.1, 0.9]])
    plain = rec.copy()
    plain[:14] = [-1.0] * 12 + [0.0, 0.0]
    # equivalent loads of the offset axis before any modifier: record with offsets only, then undo the arms
    p_axis = np.zeros(12)
    elem_host.h_beam_loads(_p(a + off[0, 0]), _p(b + off[0, 1]), C.c_double(roll), _p(Pp), mtype, _p(dl), 1, None, 0, None, _p(p_axis))
    got = np.zeros(12)
    elem_host.h_beam_loads(_p(a), _p(b), C.c_double(roll), _p(Pp), mtype, _p(dl), 1, None, 0, _p(rec), _p(got))
    _, pc = fo.member_modifiers(k, p_axis[None, :], rec)
    Em = fo.rigid_arm_matrix(R[0] @ off[0, 0], R[0] @ off[0, 1])
    want = Em.T @ pc[0]
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


def _frame_doc(second_order):
    doc = copy.deepcopy(helpers.golden("space_frame_3d")["doc"])
    members = doc["model"]["members"]
    members[0]["start_offset"] = {"X": 0.0, "Y": 0.3, "Z": 0.0}
    members[1]["end_offset"] = {"X": -0.2, "Y": 0.0, "Z": 0.1}
    members[2]["start_offset"] = {"X": 0.1, "Y": 0.1, "Z": -0.1}
    members[2]["end_offset"] = {"X": 0.0, "Y": -0.25, "Z": 0.0}
    o = doc["analysis"]["options"]
    if second_order:
        o["order"] = "NONLINEAR"; o["nonlinear_method"] = "P_DELTA"; o["pdelta_formulation"] = "CONSISTENT"
        o["tolerance"] = 1e-10; o["max_iterations"] = 30
    else:
        o["order"] = "LINEAR"
    return doc


@pytest.mark.parametrize("second_order", [False, True])
def test_oracle_frame_with_offsets_solves(second_order):
    model = wire.parse_document(_frame_doc(second_order))
    res = fo.analyze(model)
    assert np.isfinite(res.U).all() and np.abs(res.U).max() > 0.0
    plain = fo.analyze(wire.parse_document(helpers.golden("space_frame_3d")["doc"]))
    if not second_order:
        assert np.abs(res.U - plain.U).max() > 1e-9           # the offsets do change the answer


@pytest.mark.gpu
@pytest.mark.parametrize("second_order", [False, True])
def test_gpu_matches_oracle(engine_mod, second_order):
    doc = _frame_doc(second_order)
    want = fo.analyze(wire.parse_document(doc))
    got = engine_mod.analyze(wire.parse_document(doc))
    helpers.assert_results_close(got, want, 1e-9 if not second_order else 1e-8, label="end offsets")
    assert np.array_equal(got.iterations, want.iterations)


@pytest.mark.gpu
def test_gpu_cantilever_closed_forms(engine_mod):
    got = engine_mod.analyze(wire.parse_document(_doc(start=(0.5, 0.0, 0.0))))
    Le = L_NODES - 0.5
    assert got.U[0, 1, 1] == pytest.approx(-P_TIP * Le ** 3 / (3 * E_MOD * IZ), rel=1e-9)
    assert got.R[0, 0, 5] == pytest.approx(P_TIP * L_NODES, rel=1e-9)
    for s, x in enumerate((0.25, 0.5, 0.75)):
        assert got.sec[0, 0, s, 5] == pytest.approx(P_TIP * Le * (1.0 - x), rel=1e-9)
