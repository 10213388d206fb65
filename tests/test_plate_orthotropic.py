"""Orthotropic plate material (OrthotropicPlateMaterial, pydantic_models.py:663-681; material axes = the plate's local x, y).
Closed forms: a membrane strip in uniform tension has strain sigma / E_x (or E_y) along the load and -nu_xy sigma / E_x
(or -nu_yx sigma / E_y, nu_yx = nu_xy E_y / E_x) across it -- CST elements reproduce a constant strain exactly; a square
plate bent into a cylinder by edge moments has curvature 12 M / (E_x t^3) with an anticlastic -nu_xy share.  An orthotropic
material with isotropic constants must reproduce the isotropic analysis.  The engine that defines the defaults of G_xz / G_yz
is closed (stated in wire.py: they default to G_xy)."""
import copy
import json

import numpy as np
import pytest

from fers_core_b200 import wire
from oracle import fers_oracle
from tests import helpers

EX, EY, GXY, NXY = 11.0e9, 0.7e9, 0.65e9, 0.35
T, LX, LY = 0.02, 2.0, 0.5


def _doc(direction="x", ortho=True, behavior="Membrane", n=(4, 2)):
    """Rectangular strip LX x LY in the XY plane, nx x ny quads split into triangles; edge X = 0 (or Y = 0) held against the
    load direction, one node held against the transverse direction; uniform edge traction on the opposite edge."""
    doc = copy.deepcopy(helpers.golden("plate_strip_surface_load")["doc"])
    m, a = doc["model"], doc["analysis"]
    nx, ny = n
    nid = lambda i, j: 1 + i + (nx + 1) * j
    free = {"condition_type": "Free", "stiffness": None}
    fixed = {"condition_type": "Fixed", "stiffness": None}
    along, across = ("X", "Y") if direction == "x" else ("Y", "X")
    sup = lambda sid, **kw: dict({"id": sid, "classification": None, "warping_condition": None, "X": free, "Y": free, "Z": fixed,
                                  "RX": fixed, "RY": fixed, "RZ": fixed}, **kw)
    m["nodal_supports"] = [sup(1, **{along: fixed}), sup(2, **{along: fixed, across: fixed}), sup(3)]
    m["nodes"] = []
    for j in range(ny + 1):
        for i in range(nx + 1):
            on_edge = (i == 0) if direction == "x" else (j == 0)
            corner = i == 0 and j == 0
            m["nodes"].append({"id": nid(i, j), "classification": "", "X": LX * i / nx, "Y": LY * j / ny, "Z": 0.0,
                               "nodal_support": 2 if corner else (1 if on_edge else 3)})
    mat = {"id": 1, "name": "Ply", "e_mod": EX, "g_mod": EX / (2 * (1 + 0.3)), "density": 500.0, "yield_stress": 30e6}
    if ortho:
        mat["orthotropic_plate"] = {"e_x": EX, "e_y": EY, "g_xy": GXY, "nu_xy": NXY, "g_xz": None, "g_yz": None}
    m["materials"] = [mat]
    m["members"], m["member_sets"], m["sections"], m["plate_surfaces"] = [], [], [], []
    m["plate_elements"] = []
    pid = 1
    for j in range(ny):
        for i in range(nx):
            for tri in ([nid(i, j), nid(i + 1, j), nid(i + 1, j + 1)], [nid(i, j), nid(i + 1, j + 1), nid(i, j + 1)]):
                m["plate_elements"].append({"id": pid, "node_ids": tri, "material": 1, "thickness": T, "behavior": behavior,
                                            "local_x_direction": [1, 0, 0]})
                pid += 1
    sigma = 2.0e6
    loads = []
    if direction == "x":
        edge = [(nid(nx, j), LY / ny) for j in range(ny + 1)]
        vec = [1, 0, 0]
    else:
        edge = [(nid(i, ny), LX / nx) for i in range(nx + 1)]
        vec = [0, 1, 0]
    for k, (node, h) in enumerate(edge):
        w = 0.5 if k in (0, len(edge) - 1) else 1.0
        loads.append({"id": k + 1, "node": node, "magnitude": sigma * T * h * w, "direction": vec, "load_type": "force"})
    a["options"].update(order="LINEAR", solve_loadcases=True)
    a["load_cases"] = [dict(a["load_cases"][0], id=1, name="Tension", nodal_loads=loads, surface_loads=[], plate_pressures=[])]
    a["load_combinations"], a["imperfection_cases"], a["unity_checks"] = [], [], []
    return doc, sigma, nid


@pytest.mark.parametrize("direction", ["x", "y"])
def test_uniaxial_strip_strains(direction):
    doc, sigma, nid = _doc(direction)
    model = wire.parse_document(doc)
    assert model.p_data.shape[1] == wire.NPDATA and (model.p_data[:, wire.PD_ORTHO] == 1.0).all()
    assert not [w for w in model.warnings if "ORTHOTROPIC" in w[0]]
    res = fers_oracle.analyze(model)
    u = res.U[0]
    far = nid(4, 2) - 1
    if direction == "x":
        assert u[far, 0] == pytest.approx(sigma / EX * LX, rel=1e-10)
        assert u[far, 1] == pytest.approx(-NXY * sigma / EX * LY, rel=1e-10)
    else:
        assert u[far, 1] == pytest.approx(sigma / EY * LY, rel=1e-10)
        assert u[far, 0] == pytest.approx(-(NXY * EY / EX) * sigma / EY * LX, rel=1e-10)


def test_isotropic_constants_reproduce_the_isotropic_plate():
    E, nu = EX, 0.3
    iso, _, _ = _doc("x", ortho=False, behavior="Shell")
    ort, _, _ = _doc("x", ortho=True, behavior="Shell")
    g = E / (2 * (1 + nu))
    ort["model"]["materials"][0]["orthotropic_plate"] = {"e_x": E, "e_y": E, "g_xy": g, "nu_xy": nu, "g_xz": g, "g_yz": g}
    for d in (iso, ort):        # a transverse load as well, so that bending and shear take part
        d["analysis"]["load_cases"][0]["nodal_loads"].append({"id": 99, "node": 15, "magnitude": 50.0, "direction": [0, 0, -1], "load_type": "force"})
        for s in d["model"]["nodal_supports"]:
            if s["id"] == 3:
                s.update(Z={"condition_type": "Free", "stiffness": None}, RX={"condition_type": "Free", "stiffness": None},
                         RY={"condition_type": "Free", "stiffness": None})
    a, b = fers_oracle.analyze(wire.parse_document(iso)), fers_oracle.analyze(wire.parse_document(ort))
    assert np.abs(a.U - b.U).max() <= 1e-10 * np.abs(a.U).max()
    assert np.abs(a.p_res - b.p_res).max() <= 1e-9 * np.abs(a.p_res).max()


def test_bad_orthotropic_constants_are_refused_and_plane_strain_warns():
    doc, _, _ = _doc("x")
    doc["model"]["materials"][0]["orthotropic_plate"]["nu_xy"] = 5.0          # nu_xy^2 E_y / E_x >= 1: not positive definite
    with pytest.raises(wire.WireError):
        wire.parse_document(doc)
    doc, _, _ = _doc("x")
    for p in doc["model"]["plate_elements"]:
        p["plane_state"] = "PlaneStrain"
    assert "W_ORTHOTROPIC_PLANE_STRAIN" in [w[0] for w in wire.parse_document(doc).warnings]


@pytest.mark.gpu
@pytest.mark.parametrize("direction,behavior", [("x", "Membrane"), ("y", "Shell")])
def test_gpu_matches_oracle(direction, behavior):
    import fers_calculations
    doc, sigma, nid = _doc(direction, behavior=behavior, n=(6, 3))
    if behavior == "Shell":
        doc["analysis"]["load_cases"][0]["nodal_loads"].append({"id": 99, "node": 20, "magnitude": 50.0, "direction": [0, 0, -1], "load_type": "force"})
        doc["model"]["materials"][0]["orthotropic_plate"].update(g_xz=0.3e9, g_yz=0.04e9)
        for s in doc["model"]["nodal_supports"]:
            if s["id"] == 3:
                s.update(Z={"condition_type": "Free", "stiffness": None}, RX={"condition_type": "Free", "stiffness": None},
                         RY={"condition_type": "Free", "stiffness": None})
    model = wire.parse_document(doc)
    want = fers_oracle.analyze(model)
    from fers_core_b200 import engine
    got = engine.analyze(model)
    for name in ("U", "R", "p_res", "p_nodal"):
        x, y = getattr(got, name), getattr(want, name)
        assert np.abs(x - y).max() <= 1e-9 * np.abs(y).max(), name
    out = json.loads(fers_calculations.calculate_from_json(json.dumps(doc)))["results"]
    assert "Tension" in out["loadcases"]
