"""Product of inertia / principal axes of a section (Section.i_yz, principal_axis_angle, pydantic_models.py:1078-1090):
wire parser (wire.section_bending_inertias), oracle against beam theory, GPU path against the oracle.

Closed form used: Euler-Bernoulli cantilever of length L loaded at the tip by P_y; with the bending energy
E/2 int (I_z v''^2 + I_y w''^2 + 2 I_yz v'' w'') dx the curvatures solve [[EI_z, EI_yz], [EI_yz, EI_y]] (v'', w'') =
(P_y (L - x), 0), so  v_tip = P_y L^3 / 3 * EI_y / D,  w_tip = -P_y L^3 / 3 * EI_yz / D,  D = EI_y EI_z - EI_yz^2.
The engine that defines the sign of i_yz is closed: the convention is the one stated in wire.section_bending_inertias,
pinned here only for self-consistency (a principal-axis angle is the same thing as a member rotation_angle)."""
import copy

import numpy as np
import pytest

from fers_core_b200 import wire
from oracle import fers_oracle as fo
from tests import helpers

P_TIP = 1000.0


@pytest.fixture(scope="module")
def engine_mod():
    from fers_core_b200 import engine
    return engine


def _doc(i_yz=None, angle=None, roll=0.0, i_y=None, i_z=None, load_dir=(0, -1, 0)):
    doc = copy.deepcopy(helpers.golden("cantilever_end_load")["doc"])
    sec = doc["model"]["sections"][0]
    if i_yz is not None:
        sec["i_yz"] = i_yz
    if angle is not None:
        sec["principal_axis_angle"] = angle
    if i_y is not None:
        sec["i_y"], sec["i_z"] = i_y, i_z
    doc["model"]["members"][0]["rotation_angle"] = roll
    doc["analysis"]["options"]["order"] = "LINEAR"
    doc["analysis"]["load_cases"][0]["nodal_loads"][0]["direction"] = list(load_dir)
    return doc


def _numbers(doc):
    sec = doc["model"]["sections"][0]
    mat = next(m for m in doc["model"]["materials"] if m["id"] == sec["material"])
    L = abs(doc["model"]["nodes"][1]["X"] - doc["model"]["nodes"][0]["X"])
    return mat["e_mod"], sec["i_y"], sec["i_z"], L


def test_section_bending_inertias():
    f = wire.section_bending_inertias
    assert f(2.0, 5.0, None, None) == (2.0, 5.0, 0.0, None)
    assert f(2.0, 5.0, 1.0, 0.0) == (2.0, 5.0, 1.0, None)
    # principal values turned by 30 degrees: invariants kept, tan 2a = 2 I_yz / (I_z - I_y)
    iy, iz, iyz, note = f(2.0, 5.0, None, 30.0)
    assert note is None and iy + iz == pytest.approx(7.0) and iy * iz - iyz * iyz == pytest.approx(10.0)
    assert np.degrees(0.5 * np.arctan2(2 * iyz, iz - iy)) == pytest.approx(30.0)
    # a centroidal triple with its own (consistent) angle is used as it is; 90 degrees off is the same principal frame
    assert f(iy, iz, iyz, 30.0) == (iy, iz, iyz, None)
    assert f(iy, iz, iyz, 120.0) == (iy, iz, iyz, None)
    # an angle that contradicts the triple: the angle wins, with a note
    a, b, c, note = f(2.0, 5.0, 0.4, 30.0)
    assert (a, b, c) == pytest.approx((iy, iz, iyz)) and "does not match" in note
    with pytest.raises(wire.WireError):
        f(2.0, 5.0, float("nan"), None)


def test_parser_builds_coupling_record():
    E, Iy, Iz, _ = _numbers(_doc())
    m = wire.parse_document(_doc(i_yz=1.5e-6))
    assert m.rel.shape == (1, wire.REL_STRIDE) and m.m_rel.tolist() == [0]
    assert m.rel[0, wire.REL_EIYZ] == pytest.approx(E * 1.5e-6) and (m.rel[0, :12] == wire.RIGID_RELEASE).all()
    assert m.prop[0, 1] == pytest.approx(E * Iy) and m.prop[0, 2] == pytest.approx(E * Iz)
    assert not [w for w in m.warnings if w[0] == "W_PRINCIPAL_AXES_IGNORED"]
    assert [w[0] for w in m.warnings].count("W_IYZ_UNITS") == 1     # m^4 like i_y / i_z, the schema says mm^4
    # the schema's unit for this one field (mm^4, pydantic_models.py:1080): same record, no note
    mm = wire.parse_document(_doc(i_yz=1.5e6))
    assert mm.rel[0, wire.REL_EIYZ] == pytest.approx(E * 1.5e-6, rel=1e-12) and "W_IYZ_UNITS" not in [w[0] for w in mm.warnings]
    assert wire.i_yz_metres4(Iy, Iz, None) == (0.0, None) and wire.i_yz_metres4(Iy, Iz, -2.0e6)[0] == pytest.approx(-2.0e-6)
    assert wire.parse_document(_doc()).m_rel.tolist() == [-1]
    with pytest.raises(wire.WireError):                  # I_yz^2 < I_y I_z for any real section, in either unit
        wire.parse_document(_doc(i_yz=1.0e7))
    truss = _doc(i_yz=1.5e-6)
    truss["model"]["members"][0]["member_type"] = "Truss"
    assert wire.parse_document(truss).m_rel.tolist() == [-1]      # no bending, no coupling


@pytest.mark.parametrize("i_yz", [1.5e-6, -2.5e-6])
def test_oracle_cantilever_deflects_out_of_the_load_plane(i_yz):
    doc = _doc(i_yz=i_yz)
    E, Iy, Iz, L = _numbers(doc)
    res = fo.analyze(wire.parse_document(doc))
    D = E * E * (Iy * Iz - i_yz * i_yz)
    v = -P_TIP * L ** 3 / 3.0 * E * Iy / D
    w = P_TIP * L ** 3 / 3.0 * E * i_yz / D
    assert res.U[0, 1, 1] == pytest.approx(v, rel=1e-9) and res.U[0, 1, 2] == pytest.approx(w, rel=1e-9)
    # statics do not see the section: clamp force and moment
    assert res.R[0, 0, 1] == pytest.approx(P_TIP, rel=1e-9) and res.R[0, 0, 5] == pytest.approx(P_TIP * L, rel=1e-9)
    assert abs(res.R[0, 0, 4]) <= 1e-6 * P_TIP * L


@pytest.mark.parametrize("angle", [30.0, -20.0, 75.0])
@pytest.mark.parametrize("load_dir", [(0, -1, 0), (0, 0.6, 0.8)])
def test_principal_angle_is_a_member_rotation(angle, load_dir):
    """Principal values + principal_axis_angle on an un-rotated member = the same values about the local axes of a
    member turned by rotation_angle (nodal displacements; the local frames differ, the global response does not)."""
    a = fo.analyze(wire.parse_document(_doc(angle=angle, load_dir=load_dir)))
    b = fo.analyze(wire.parse_document(_doc(roll=angle, load_dir=load_dir)))
    assert np.abs(a.U - b.U).max() <= 1e-9 * np.abs(b.U).max()
    assert np.abs(a.R - b.R).max() <= 1e-9 * np.abs(b.R).max()
    assert abs(a.U[0, 1, 2]) > 1e-3 * abs(a.U[0, 1, 1]) or load_dir[2]


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(i_yz=1.5e-6), dict(angle=35.0)])
def test_gpu_matches_oracle(engine_mod, kw):
    doc = _doc(**kw)
    # a distributed load, a release and a second, skew member so that loads, rotation and recovery all see the record
    m = doc["model"]
    m["nodes"].append({"id": 3, "X": 7.0, "Y": 1.5, "Z": -2.0, "nodal_support": None})
    m["members"].append(dict(m["members"][0], id=2, start_node_id=2, end_node_id=3, rotation_angle=25.0))
    doc["analysis"]["load_cases"][0]["distributed_loads"] = [
        {"id": 1, "member": 2, "magnitude": 800.0, "end_magnitude": 200.0, "direction": [0, -1, 0], "start_frac": 0.1, "end_frac": 0.8}]
    want = fo.analyze(wire.parse_document(doc))
    got = engine_mod.analyze(wire.parse_document(doc))
    helpers.assert_results_close(got, want, 1e-9, label="product of inertia")
    assert abs(got.U[0, 1, 2]) > 0.0


@pytest.mark.gpu
def test_gpu_second_order_matches_oracle(engine_mod):
    doc = _doc(i_yz=1.5e-6, load_dir=(-0.9, -0.4, 0.0))       # compression + bending: K_g on top of the coupled K_e
    doc["analysis"]["options"]["order"] = "NONLINEAR"
    doc["analysis"]["load_cases"][0]["nodal_loads"][0]["magnitude"] = 2.0e3
    want = fo.analyze(wire.parse_document(doc))
    got = engine_mod.analyze(wire.parse_document(doc))
    helpers.assert_results_close(got, want, 1e-7, label="product of inertia, second order")
