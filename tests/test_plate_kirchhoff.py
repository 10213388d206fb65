"""PlateTheory.Kirchhoff (pydantic_models.py:816-819, PlateElement.theory 1723): the discrete Kirchhoff triangle (DKT).

The reference calls its own DKT element "currently unreliable" (tests/nafems/reference_examples.py:16-18) and pins nothing
on it, so this element is pinned on thin-plate theory: constant-curvature patch states (exact), the simply supported square
plate under uniform load (Navier: w = 0.00406235 q a^4 / D, M = 0.0479 q a^2 at the centre, nu = 0.3) on a THICK plate
(t / a = 0.1, where the Mindlin element answers 4-5 % more: the switch is visible), a cantilever strip (w = P L^3 / 3 E I,
M = P (L - x) / b).  The CUDA element (published coefficient lists of Batoz, Bathe, Ho 1980) is held against the oracle's element
(built from the definition of the DKT constraints) through the host build of fers_elem.cuh."""
import copy
import ctypes

import numpy as np
import pytest

from fers_core_b200 import wire
from oracle import fers_oracle as fo
from tests import helpers

C = ctypes
E_MOD, NU = 210e9, 0.3


@pytest.fixture(scope="module")
def engine_mod():
    from fers_core_b200 import engine
    return engine


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def plate_doc(lx, ly, nx, ny, t, theory, nu=NU, support="simple", pressure=None, tip_load=None):
    """lx x ly plate in the XY plane, nx x ny quads cut into triangles by alternating diagonals, bending only
    (behavior Plate).  support: "simple" = w held on all four edges, "cantilever" = edge X = 0 clamped.  pressure: uniform
    load per area, lumped to the nodes by tributary area; tip_load: total force spread over the edge X = lx."""
    doc = copy.deepcopy(helpers.golden("plate_strip_surface_load")["doc"])
    m, a = doc["model"], doc["analysis"]
    nid = lambda i, j: 1 + i + (nx + 1) * j
    free = {"condition_type": "Free", "stiffness": None}
    fixed = {"condition_type": "Fixed", "stiffness": None}
    sup = lambda sid, **kw: dict({"id": sid, "classification": None, "warping_condition": None, "X": fixed, "Y": fixed, "Z": free,
                                  "RX": free, "RY": free, "RZ": fixed}, **kw)
    m["nodal_supports"] = [sup(1), sup(2, Z=fixed), sup(3, Z=fixed, RX=fixed, RY=fixed)]
    m["nodes"] = []
    for j in range(ny + 1):
        for i in range(nx + 1):
            if support == "simple":
                s = 2 if (i in (0, nx) or j in (0, ny)) else 1
            else:
                s = 3 if i == 0 else 1
            m["nodes"].append({"id": nid(i, j), "classification": "", "X": lx * i / nx, "Y": ly * j / ny, "Z": 0.0, "nodal_support": s})
    m["materials"] = [{"id": 1, "name": "Steel", "e_mod": E_MOD, "g_mod": E_MOD / (2 * (1 + nu)), "density": 7850.0, "yield_stress": 235e6}]
    m["members"], m["member_sets"], m["sections"], m["plate_surfaces"] = [], [], [], []
    m["plate_elements"] = []
    pid = 1
    for j in range(ny):
        for i in range(nx):
            q = [nid(i, j), nid(i + 1, j), nid(i + 1, j + 1), nid(i, j + 1)]
            tris = ([q[0], q[1], q[2]], [q[0], q[2], q[3]]) if (i + j) % 2 == 0 else ([q[0], q[1], q[3]], [q[1], q[2], q[3]])
            for tri in tris:
                m["plate_elements"].append({"id": pid, "node_ids": tri, "material": 1, "thickness": t, "behavior": "Plate",
                                            "theory": theory})
                pid += 1
    loads = []
    if pressure is not None:
        for j in range(ny + 1):
            for i in range(nx + 1):
                wx = 0.5 if i in (0, nx) else 1.0
                wy = 0.5 if j in (0, ny) else 1.0
                loads.append({"id": len(loads) + 1, "node": nid(i, j), "magnitude": pressure * lx / nx * ly / ny * wx * wy,
                              "direction": [0, 0, -1], "load_type": "force"})
    if tip_load is not None:
        for j in range(ny + 1):
            wy = 0.5 if j in (0, ny) else 1.0
            loads.append({"id": len(loads) + 1, "node": nid(nx, j), "magnitude": tip_load * wy / ny, "direction": [0, 0, -1],
                          "load_type": "force"})
    a["options"].update(order="LINEAR", solve_loadcases=True)
    a["load_cases"] = [dict(a["load_cases"][0], id=1, name="Load", nodal_loads=loads, surface_loads=[], plate_pressures=[])]
    a["load_combinations"], a["imperfection_cases"], a["unity_checks"] = [], [], []
    return doc, nid


# ---- the element ---------------------------------------------------------------------------------------------------------
def test_constant_curvature_states_are_exact():
    xl = np.array([[0.0, 0.0], [2.0, 0.3], [0.7, 1.9]])
    # w = x^2/2 + 0.3 x y + 0.2 y^2:  b = -grad w,  kappa = [bx,x  by,y  bx,y + by,x] = [-1, -0.4, -0.6] everywhere
    u = []
    for x, y in xl:
        u += [0.5 * x * x + 0.3 * x * y + 0.2 * y * y, -(x + 0.3 * y), -(0.3 * x + 0.4 * y)]
    for gp in ((1 / 3, 1 / 3), (0.5, 0.5), (0.1, 0.7), (0.0, 0.0)):
        assert np.allclose(fo.dkt_curvature(xl, *gp) @ np.array(u), [-1.0, -0.4, -0.6], atol=1e-13)
    # rigid motions w = a + b x + c y carry no curvature
    u = []
    for x, y in xl:
        u += [0.3 + 0.1 * x - 0.2 * y, -0.1, 0.2]
    assert np.abs(fo.dkt_curvature(xl, 0.2, 0.3) @ np.array(u)).max() < 1e-14


def test_device_element_equals_the_oracle_element(elem_host):
    rng = np.random.default_rng(21)
    elem_host.h_plate_B.argtypes = [C.POINTER(C.c_double)] * 4 + [C.c_int, C.POINTER(C.c_double)]
    for t in range(10):
        p = rng.normal(size=(3, 3))
        if t < 3:
            p[:, 2] = -0.4
        data = np.zeros(wire.NPDATA)
        data[:7] = (0.05, E_MOD, NU, 1.0, 1.0, 1.0, 0.0)
        if t % 3 == 1:
            data[3:6] = (0.4, 0.7, 1.3)
        ortho = None
        if t % 4 == 2:
            ortho = (11.0e9, 0.7e9, 0.65e9, 0.35, 0.6e9, 0.05e9)
            data[wire.PD_ORTHO] = 1.0
            data[wire.PD_EX:wire.PD_GYZ + 1] = ortho
        if t % 5 == 4:
            data[wire.PD_OFFSET] = 0.03
        for flags in (2 | 8, 3 | 8):
            R = fo.plate_frame(p)
            xl = (p - p[0]) @ R.T
            k, mats = fo.plate_local_matrices(xl[:, :2], data[0], data[1], data[2], flags, tuple(data[3:6]), ortho)
            if data[wire.PD_OFFSET] != 0.0:
                Eo = fo.plate_offset_matrix(data[wire.PD_OFFSET])
                k = Eo.T @ k @ Eo
            T = np.zeros((18, 18))
            for bq in range(6):
                T[3 * bq:3 * bq + 3, 3 * bq:3 * bq + 3] = R
            want = T.T @ k @ T
            out = np.zeros(324)
            elem_host.h_plate_ke(_p(p[0].copy()), _p(p[1].copy()), _p(p[2].copy()), _p(data), flags, _p(out))
            assert np.abs(out.reshape(18, 18) - want).max() < 1e-11 * np.abs(want).max()
            # operators of the results: curvature at the centroid, shear forces from the moments
            _Bm, Bb, Bs, _Dm, _Db, Ds, _A = mats
            S = np.zeros((9, 18))
            for n in range(3):
                S[3 * n, 6 * n + 2] = 1.0; S[3 * n + 1, 6 * n + 4] = 1.0; S[3 * n + 2, 6 * n + 3] = -1.0
            ops = np.zeros(8 * 18)
            elem_host.h_plate_B(_p(p[0].copy()), _p(p[1].copy()), _p(p[2].copy()), _p(data), flags, _p(ops))
            ops = ops.reshape(8, 18)
            assert np.abs(ops[3:6] - Bb @ S).max() < 1e-11 * np.abs(Bb).max()
            assert not ops[6:8].any() and not Bs.any()          # no shear rows (W_KIRCHHOFF_NO_SHEAR_FORCES)


def test_wire_sets_the_theory_flag():
    doc, _ = plate_doc(1.0, 1.0, 2, 2, 0.01, "Kirchhoff")
    m = wire.parse_document(doc)
    assert (m.p_flags & 8).all() and not [w for w in m.warnings if w[0] == "W_PLATE_THEORY_IGNORED"]
    assert [w[0] for w in m.warnings].count("W_KIRCHHOFF_NO_SHEAR_FORCES") == 1
    for theory in ("Mindlin", "Auto", None):
        doc, _ = plate_doc(1.0, 1.0, 2, 2, 0.01, theory)
        assert not (wire.parse_document(doc).p_flags & 8).any()


# ---- thin-plate closed forms, oracle -------------------------------------------------------------------------------------
A_SIDE, T_THICK, Q0 = 2.0, 0.2, 5.0e3


def _centre(res, nid, n):
    return -res.U[0, nid(n // 2, n // 2) - 1, 2]


def test_simply_supported_square_plate_follows_thin_plate_theory():
    D = E_MOD * T_THICK ** 3 / (12 * (1 - NU ** 2))
    w_thin = 0.00406235 * Q0 * A_SIDE ** 4 / D
    n = 12
    doc, nid = plate_doc(A_SIDE, A_SIDE, n, n, T_THICK, "Kirchhoff", pressure=Q0)
    res = fo.analyze(wire.parse_document(doc))
    w_k = _centre(res, nid, n)
    assert w_k == pytest.approx(w_thin, rel=0.01)
    # the same thick plate (t / a = 0.1) with shear deformation deflects visibly more
    doc_m, _ = plate_doc(A_SIDE, A_SIDE, n, n, T_THICK, "Mindlin", pressure=Q0)
    w_m = _centre(fo.analyze(wire.parse_document(doc_m)), nid, n)
    assert w_m > 1.02 * w_k
    # centre moment 0.0479 q a^2 (both directions); the elements around the centre node bracket it
    m = wire.parse_document(doc)
    cen = m.xyz[m.p_nodes].mean(1)
    near = np.argsort(np.linalg.norm(cen[:, :2] - A_SIDE / 2, axis=1))[:8]
    mx = res.p_res[0, near, 3].mean()
    my = res.p_res[0, near, 4].mean()
    assert abs(mx) == pytest.approx(0.0479 * Q0 * A_SIDE ** 2, rel=0.04)
    assert abs(my) == pytest.approx(0.0479 * Q0 * A_SIDE ** 2, rel=0.04)


def test_mesh_refinement_converges_to_the_thin_plate_value():
    D = E_MOD * T_THICK ** 3 / (12 * (1 - NU ** 2))
    w_thin = 0.00406235 * Q0 * A_SIDE ** 4 / D
    errs = []
    for n in (4, 8, 16):
        doc, nid = plate_doc(A_SIDE, A_SIDE, n, n, T_THICK, "Kirchhoff", pressure=Q0)
        errs.append(abs(_centre(fo.analyze(wire.parse_document(doc)), nid, n) / w_thin - 1.0))
    assert errs[2] < errs[1] < errs[0] and errs[2] < 0.005


def test_cantilever_strip_beam_limit():
    L, b, t, P = 2.0, 0.4, 0.05, 100.0
    res = {}
    for theory in ("Kirchhoff", "Mindlin"):
        doc, nid = plate_doc(L, b, 16, 2, t, theory, nu=0.0, support="cantilever", tip_load=P)
        m = wire.parse_document(doc)
        res[theory] = (fo.analyze(m), m, nid)
    r, m, nid = res["Kirchhoff"]
    tip = -r.U[0, [nid(16, j) - 1 for j in range(3)], 2].mean()
    assert tip == pytest.approx(P * L ** 3 / (3 * E_MOD * b * t ** 3 / 12), rel=2e-3)
    # thin-plate elements report no shear forces; the Mindlin element carries Q = P / b on average
    assert not r.p_res[0, :, 6:8].any()
    assert abs(res["Mindlin"][0].p_res[0, :, 6].mean()) == pytest.approx(P / b, rel=0.02)
    # root moment per width P L / b, linear along the strip: element moments against their centroid position
    cen = m.xyz[m.p_nodes].mean(1)[:, 0]
    assert np.allclose(np.abs(r.p_res[0, :, 3]), P * (L - cen) / b, rtol=0.05, atol=0.02 * P * L / b)


# ---- GPU -----------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", ["simple", "cantilever"])
def test_gpu_matches_oracle(engine_mod, case):
    if case == "simple":
        doc, _ = plate_doc(A_SIDE, A_SIDE, 8, 8, T_THICK, "Kirchhoff", pressure=Q0)
    else:
        doc, _ = plate_doc(2.0, 0.4, 12, 2, 0.05, "Kirchhoff", support="cantilever", tip_load=100.0)
    m = wire.parse_document(doc)
    got = engine_mod.analyze(m)
    want = fo.analyze(m)
    helpers.assert_results_close(got, want, rtol=1e-9, label=case)


@pytest.mark.gpu
def test_gpu_thin_plate_value(engine_mod):
    D = E_MOD * T_THICK ** 3 / (12 * (1 - NU ** 2))
    n = 12
    doc, nid = plate_doc(A_SIDE, A_SIDE, n, n, T_THICK, "Kirchhoff", pressure=Q0)
    res = engine_mod.analyze(wire.parse_document(doc))
    assert _centre(res, nid, n) == pytest.approx(0.00406235 * Q0 * A_SIDE ** 4 / D, rel=0.01)
