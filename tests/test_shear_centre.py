"""Shear-centre eccentricity coupling (Section.y_s / z_s, pydantic_models.py:1106-1111; AnalysisOptions
.include_shear_center_coupling, 1358-1361; Member.mirror, 1521-1524): wire parser, oracle against beam theory,
host instantiation of the element code against the oracle, GPU path against the oracle.

Closed form used: a cantilever along +X loaded at the tip by P in -Y through the CENTROID of a section whose shear
centre sits at z_s: the load is statically equivalent to P through the shear centre plus the torque
(-z_s e_z) x (-P e_y) = -z_s P e_x about the shear-centre axis, so  rx_tip = -z_s P L / GJ,
v_sc = -P L^3 / (3 E Iz),  v_centroid = v_sc + z_s * rx_tip,  w_centroid = -y_s * rx_tip."""
import copy
import ctypes as C

import numpy as np
import pytest

from fers_core_b200 import wire
from oracle import fers_oracle as fo
from tests import helpers
from tests.test_elem_host import CASES, _oracle_member, _p, _rel

P_TIP, YS, ZS = 1000.0, 0.021, -0.034


@pytest.fixture(scope="module")
def engine_mod():
    from fers_core_b200 import engine
    return engine


def _doc(ys=YS, zs=ZS, coupling=True, mirror=False):
    doc = copy.deepcopy(helpers.golden("cantilever_end_load")["doc"])
    sec = doc["model"]["sections"][0]
    sec["y_s"], sec["z_s"] = ys, zs
    doc["model"]["members"][0]["mirror"] = mirror
    doc["analysis"]["options"]["order"] = "LINEAR"
    doc["analysis"]["options"]["include_shear_center_coupling"] = coupling
    return doc


def _section_numbers(doc):
    sec = doc["model"]["sections"][0]
    mat = sec["material"] if isinstance(sec["material"], dict) else \
        next(m for m in doc["model"]["materials"] if m["id"] == sec["material"])
    nodes = doc["model"]["nodes"]
    L = abs(nodes[1]["X"] - nodes[0]["X"])
    return mat["e_mod"], mat["g_mod"], sec["i_z"], sec["j"], L


def test_parser_builds_modifier_records():
    m = wire.parse_document(_doc())
    assert m.rel.shape == (1, wire.REL_STRIDE) and m.m_rel.tolist() == [0]
    assert (m.rel[0, :12] == wire.RIGID_RELEASE).all() and m.rel[0, 12:14].tolist() == [YS, ZS]
    assert wire.parse_document(_doc(mirror=True)).rel[0, 12:14].tolist() == [-YS, -ZS]
    off = wire.parse_document(_doc(coupling=False))
    assert off.rel.shape == (0, wire.REL_STRIDE) and off.m_rel.tolist() == [-1]
    assert not [w for w in m.warnings if w[0] == "W_SHEAR_CENTRE_IGNORED"]


@pytest.mark.parametrize("mirror", [False, True])
def test_oracle_cantilever_twists_about_shear_centre(mirror):
    doc = _doc(mirror=mirror)
    E, G, Iz, J, L = _section_numbers(doc)
    sgn = -1.0 if mirror else 1.0
    res = fo.analyze(wire.parse_document(doc))
    u = res.U[0, 1]                                   # tip node, load case 0: [dx dy dz rx ry rz]
    rx = -sgn * ZS * P_TIP * L / (G * J)
    v_sc = -P_TIP * L ** 3 / (3 * E * Iz)
    assert u[3] == pytest.approx(rx, rel=1e-9)
    assert u[1] == pytest.approx(v_sc + sgn * ZS * rx, rel=1e-9)
    assert u[2] == pytest.approx(-sgn * YS * rx, rel=1e-9, abs=1e-15)   # w_sc = w + y_s rx = 0
    # reactions at the clamp: equilibrium about the centroid axis (no torque: the load passes through the centroid)
    r = res.R[0, 0]
    assert r[1] == pytest.approx(P_TIP, rel=1e-9) and r[5] == pytest.approx(P_TIP * L, rel=1e-9)
    assert abs(r[3]) < 1e-6
    # coupling off: plain beam
    plain = fo.analyze(wire.parse_document(_doc(coupling=False)))
    assert plain.U[0, 1, 3] == 0.0 and plain.U[0, 1, 1] == pytest.approx(v_sc, rel=1e-9)


@pytest.mark.parametrize("case", CASES[:8])
def test_element_code_matches_oracle(elem_host, case):
    a, b, roll, Pp, mtype = case
    R, L, k, T = _oracle_member(a, b, roll, Pp, mtype)
    rec = np.full(wire.REL_STRIDE, -1.0)
    rec[wire.REL_EIYZ] = 0.0
    rec[5] = 0.0
    rec[9] = 4.0e3
    rec[12:14] = (YS, ZS)
    rec[14:20] = 0.0
    p = np.arange(1.0, 13.0)[None, :] * 17.0
    kc, pc = fo.member_modifiers(k, p, rec)
    out = np.zeros(144)
    elem_host.h_beam_ke(_p(a), _p(b), C.c_double(roll), _p(Pp), int(mtype), C.c_double(1.0), _p(rec), 1, _p(out))
    assert _rel(out.reshape(12, 12), T.T @ kc @ T) < 1e-12
    # symmetric, and a rigid-body twist about the SHEAR-CENTRE axis costs nothing when nothing is released
    rec[:12] = -1.0
    ks, _ = fo.member_modifiers(k, np.zeros((0, 12)), rec)
    assert np.abs(ks - ks.T).max() <= 1e-9 * np.abs(ks).max()
    twist = np.zeros(12)
    twist[[3, 9]] = 1.0
    twist[[1, 7]] = ZS       # v = v_sc + z_s rx with v_sc = 0
    twist[[2, 8]] = -YS
    assert np.abs(ks @ twist).max() <= 1e-9 * np.abs(ks).max()


def test_element_loads_match_oracle(elem_host):
    a, b, roll, Pp, mtype = CASES[1]
    mtype = 0
    R, L, k, T = _oracle_member(a, b, roll, Pp, mtype)
    rec = np.full(wire.REL_STRIDE, -1.0)
    rec[wire.REL_EIYZ] = 0.0
    rec[11] = 0.0
    rec[12:14] = (YS, ZS)
    rec[14:20] = 0.0
    dl = np.array([[0.0, -1200.0, 300.0, 0.0, -400.0, 100.0, 0.1, 0.9]])
    p0 = np.zeros(12)
    elem_host.h_beam_loads(_p(a), _p(b), C.c_double(roll), _p(Pp), mtype, _p(dl), 1, None, 0, None, _p(p0))
    p1 = np.zeros(12)
    elem_host.h_beam_loads(_p(a), _p(b), C.c_double(roll), _p(Pp), mtype, _p(dl), 1, None, 0, _p(rec), _p(p1))
    _, want = fo.member_modifiers(k, p0[None, :], rec)
    assert np.abs(p1 - want[0]).max() <= 1e-12 * np.abs(want).max()


@pytest.mark.gpu
@pytest.mark.parametrize("mirror", [False, True])
def test_gpu_matches_oracle(engine_mod, mirror):
    doc = _doc(mirror=mirror)
    # a distributed load and a second, skew member so that loads, rotation and recovery all see the record
    m = doc["model"]
    m["nodes"].append({"id": 3, "X": 7.0, "Y": 1.5, "Z": -2.0, "nodal_support": None})
    m["members"].append(dict(m["members"][0], id=2, start_node_id=2, end_node_id=3, rotation_angle=25.0))
    doc["analysis"]["load_cases"][0]["distributed_loads"] = [
        {"id": 1, "member": 2, "magnitude": 800.0, "end_magnitude": 200.0, "direction": [0, -1, 0], "start_frac": 0.1, "end_frac": 0.8}]
    model = wire.parse_document(doc)
    want = fo.analyze(model)
    got = engine_mod.analyze(wire.parse_document(doc))
    helpers.assert_results_close(got, want, 1e-9, label="shear centre")
    assert abs(got.U[0, 1, 3]) > 0.0
