"""csrc/fers_elem.cuh (compiled for the host by g++) against the numpy oracle.

Every kernel's per-element arithmetic lives in fers_elem.cuh as FERS_HD functions; this pins it
on the CPU before any GPU time is spent.  Tolerance: 1e-13 relative per entry (fp64, same
formulas, different summation order).
"""
import ctypes

import numpy as np
import pytest

from fers_core_b200 import wire
from oracle import fers_oracle as fo

C = ctypes
dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(dp)


def _rel(a, b):
    s = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / s


RNG = np.random.default_rng(11)
CASES = []
for t in range(24):
    a = RNG.normal(size=3) * 3
    b = a + RNG.normal(size=3) * 4
    if t == 0:
        b = a + np.array([0.0, 3.5, 0.0])       # vertical: +Z reference branch
    if t == 1:
        b = a + np.array([0.0, -2.0, 0.0])
    if t == 2:
        b = a + np.array([5.0, 0.0, 0.0])
    roll = 0.0 if t % 3 == 0 else RNG.uniform(-3, 3)
    P = np.array([3.1e9, 1.8e7, 5.2e7, 4e5, 0.0, 0.0, 0, 0])
    if t % 2:
        P[4:6] = [8.3e8, 2.6e8]
    mtype = wire.MT_TRUSS if t % 7 == 6 else wire.MT_NORMAL
    CASES.append((a, b, roll, P, mtype))


def _oracle_member(a, b, roll, P, mtype):
    xyz = np.stack([a, b])
    R, L = fo.member_frames(xyz, np.array([0]), np.array([1]), np.array([roll]))
    k = fo.local_stiffness(P[None, :], L, np.array([mtype]))
    return R[0], L[0], k[0], fo.expand_T(R)[0]


@pytest.mark.parametrize("case", CASES)
def test_beam_global_stiffness(elem_host, case):
    a, b, roll, P, mtype = case
    R, L, k, T = _oracle_member(a, b, roll, P, mtype)
    want = T.T @ k @ T
    for mode in (0, 1):
        out = np.zeros(144)
        elem_host.h_beam_ke(_p(a), _p(b), C.c_double(roll), _p(P), int(mtype), C.c_double(1.0), None, mode, _p(out))
        assert _rel(out.reshape(12, 12), want) < 1e-13, mode


@pytest.mark.parametrize("case", CASES[:8])
def test_beam_releases_and_kg(elem_host, case):
    a, b, roll, P, mtype = case
    R, L, k, T = _oracle_member(a, b, roll, P, mtype)
    rel = np.full(wire.REL_STRIDE, -1.0)
    rel[wire.REL_EIYZ] = 0.0
    rel[5] = 0.0
    rel[10] = 2.5e4
    rel[12:] = 0.0
    kc, _ = fo.condense_releases(k, np.zeros((0, 12)), rel)
    out = np.zeros(144)
    elem_host.h_beam_ke(_p(a), _p(b), C.c_double(roll), _p(P), int(mtype), C.c_double(1.0), _p(rel), 1, _p(out))
    assert _rel(out.reshape(12, 12), T.T @ kc @ T) < 1e-12
    for consistent, name in ((1, "CONSISTENT"), (0, "SIMPLIFIED")):
        N = -3.3e5
        kg = fo.geometric_stiffness(np.array([N]), np.array([L]), name, np.array([mtype]))[0]
        kk = np.zeros(144); Rr = np.zeros(9); LL = C.c_double()
        elem_host.h_beam_local(_p(a), _p(b), C.c_double(roll), _p(P), int(mtype), C.c_double(1.0), C.c_double(N),
                               consistent, _p(kk), _p(Rr), C.byref(LL))
        assert _rel(kk.reshape(12, 12), k + kg) < 1e-13
        assert _rel(Rr.reshape(3, 3), R) < 1e-15
        assert abs(LL.value - L) < 1e-15 * L


@pytest.mark.parametrize("case", CASES[:12])
def test_beam_equivalent_loads_and_span(elem_host, case):
    a, b, roll, P, mtype = case
    model_like = type("M", (), {})()
    dl = np.array([[0.0, -1e3, 200.0, 50.0, -4e3, 0.0, 0.2, 0.9],
                   [300.0, 0.0, 0.0, 300.0, 0.0, 0.0, 0.0, 1.0]])
    pl = np.array([[100.0, -2e3, 30.0, 0.0, 0.0, 0.0, 0.35, 0.0],
                   [0.0, 0.0, 0.0, 50.0, -80.0, 900.0, 0.6, 1.0]])
    # oracle
    xyz = np.stack([a, b])
    m = wire.ModelArrays(
        node_ids=np.array([1, 2]), xyz=xyz, node_support_id=np.array([-1, -1]), bc_code=np.zeros((2, 6), np.uint8),
        bc_k=np.zeros((2, 6)), m_ids=np.array([1]), m_n1=np.array([0], np.int32), m_n2=np.array([1], np.int32),
        m_type=np.array([mtype], np.int32), m_roll=np.array([roll]), m_prop=np.array([0], np.int32),
        m_rel=np.array([-1], np.int32), prop=P[None, :].copy(), prop_rho_a=np.zeros(1), rel=np.zeros((0, wire.REL_STRIDE)),
        p_ids=np.zeros(0, np.int64), p_nodes=np.zeros((0, 3), np.int32), p_data=np.zeros((0, wire.NPDATA)),
        p_flags=np.zeros(0, np.int32), load_cases=[], combo_ids=np.zeros(0, np.int64), combo_names=[],
        combo_factors=np.zeros((0, 0)), imperfections=[], member_sets={}, options=wire.Options())
    lc = wire.LoadCaseArrays(id=1, name="x")
    lc.dl_member = np.array([0, 0], np.int32); lc.dl_data = dl
    lc.pl_member = np.array([0, 0], np.int32); lc.pl_data = pl
    m.load_cases = [lc]
    asm = fo.Assembly(m)
    want = fo.equivalent_loads(asm, lc)[0]
    got = np.zeros(12)
    elem_host.h_beam_loads(_p(a), _p(b), C.c_double(roll), _p(P), int(mtype), _p(dl), 2, _p(pl), 2, None, _p(got))
    assert _rel(got, want) < 1e-13
    tab = fo.span_load_tables(asm)[0][0]
    for x in (0.25, 0.5, 0.75, 1.0):
        f, mo = fo.span_resultant(tab, asm.L[0], x)
        s = np.zeros(6)
        elem_host.h_span(_p(a), _p(b), C.c_double(roll), _p(dl), 2, _p(pl), 2, C.c_double(x), _p(s))
        assert _rel(s, np.concatenate([f, mo])) < 1e-13


def test_plate_global_stiffness(elem_host):
    rng = np.random.default_rng(5)
    for t in range(12):
        p = rng.normal(size=(3, 3))
        if t < 4:
            p[:, 2] = 0.3           # XY-plane plate: local = global (every pinned case)
        if t == 4:
            p[:, 0] = 1.0           # normal along X: alternate reference branch
        # plate record: thickness, E, nu, stiffness modifiers (membrane, bending, shear), 2 spare
        data = np.zeros(wire.NPDATA)
        data[:7] = (0.02, 210e9, 0.3, 1.0, 1.0, 1.0, 0.0)
        if t % 3 == 1:
            data[3:6] = (0.35, 0.6, 1.7)
        ortho = None
        if t % 4 == 2:              # orthotropic plate material: E_x, E_y, G_xy, nu_xy, G_xz, G_yz
            ortho = (11.0e9, 0.7e9, 0.65e9, 0.35, 0.6e9, 0.05e9)
            data[wire.PD_ORTHO] = 1.0
            data[wire.PD_EX:wire.PD_GYZ + 1] = ortho
        if t % 5 == 3:
            data[wire.PD_OFFSET] = 0.07 * (-1) ** t          # mid-surface off the nodes
        for flags in (1, 2, 3, 5, 7):           # bit2 = plane strain
            if ortho is not None and flags & 4:
                continue                        # orthotropic: plane-stress law only (wire warns)
            R = fo.plate_frame(p)
            xl = (p - p[0]) @ R.T
            k, _ = fo.plate_local_matrices(xl[:, :2], data[0], data[1], data[2], flags, tuple(data[3:6]), ortho)
            if data[wire.PD_OFFSET] != 0.0:
                Eo = fo.plate_offset_matrix(data[wire.PD_OFFSET])
                k = Eo.T @ k @ Eo
            T = np.zeros((18, 18))
            for bq in range(6):
                T[3 * bq:3 * bq + 3, 3 * bq:3 * bq + 3] = R
            want = T.T @ k @ T
            out = np.zeros(324)
            elem_host.h_plate_ke(_p(p[0].copy()), _p(p[1].copy()), _p(p[2].copy()), _p(data), flags, _p(out))
            assert _rel(out.reshape(18, 18), want) < 1e-12
