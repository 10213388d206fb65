"""CPU ORACLE for the FERS analysis hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg may import this module.  The product path (``fers_calculations`` ->
``fers_core_b200.engine`` -> ``libfers_b200.so``) never does.

What it restates.  The reference (``/root/reference``, Jeroen124/FERS_core) holds no solver
source: the arithmetic lives in the third-party Rust wheel ``fers_calculations==0.2.43``
(pyproject.toml:26, requirements.txt:6), absent from the reference tree and not installable
here.  This file therefore restates the *published* algorithm that the reference's call sites
(fers_core/fers/fers.py:260,300; tests/nafems/nafems_harness.py:38), schema docstrings
(fers_core/types/pydantic_models.py), tests and golden numbers pin:

* 3D Euler-Bernoulli/Timoshenko frame element (pydantic_models.py:1052-1114, shear pairing per
  SURVEY.md A.5), local axes of fers_core/members/member.py:309-354, T^T k T assembly;
* supports by elimination, springs on the diagonal (pydantic_models.py:2104-2115, 1956-1959);
* consistent equivalent nodal loads for trapezoid partial member loads and member point
  loads/moments (pydantic_models.py:1438-1445, 1562-1591);
* combinations as factored load vectors (pydantic_models.py:491-505);
* P-Delta iteration with consistent (Przemieniecki) / simplified geometric stiffness and sway /
  translation imperfections (pydantic_models.py:1370-1398, 1970-1991);
* CST membrane + DSG3 Mindlin TRI3 plate (pinned by
  tests/nafems/benchmarks/LE1_result.json and LE6_result.json);
* result recovery in the reference's conventions (golden
  fers_core/examples/json_input_solver/805_JSON_Import_Export_with_results.json).

Pinning status: pinned against the reference's closed-form test-suite
(tests/beam_equation_checks, tests/functionality) and the NAFEMS LE1/LE6 known-answer numbers
(see tests/test_oracle_*.py).  Parity against the Rust engine at 1e-9 is otherwise UNPINNED
(no runnable reference; see DESIGN.md "Oracle").
"""
from __future__ import annotations

import dataclasses
import time

import os

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from fers_core_b200 import wire
from fers_core_b200.results import ResultArrays, SECTION_FRACS
# the host-side rules (result columns, imperfect geometry, rigid members, active sets, load-dependent springs) in the oracle's
# own restatement -- not the product's hostlogic.py (a shared helper hides its bugs from every parity test)
from oracle import cable_oracle, host_oracle
from oracle.host_oracle import (
    rhs_plan, imperfect_coordinates, rigid_property_row, pdelta_suppressed_axes,
    unilateral_dofs, effective_bc, unilateral_update, curve_stiffness, curves_converged,
    hinge_curve_stiffness, rel_with_hinge_curves,
)

INACTIVE_SCALE = 1.0e-9      # stiffness scale of a tension- / compression-only member that is switched off

GAUSS3_X = np.array([-np.sqrt(3.0 / 5.0), 0.0, np.sqrt(3.0 / 5.0)])
GAUSS3_W = np.array([5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0])


# --------------------------------------------------------------------------------------
# element geometry  (fers_core/members/member.py:309-354)
# --------------------------------------------------------------------------------------
def member_offsets(model):
    """Per-member end offsets [Nm,2,3] in global axes (MemberEndOffset, pydantic_models.py:525-531; Member.start_offset /
    end_offset 1541 / 1518), taken from the member modifier records; zero for members without a record."""
    off = np.zeros((model.n_members, 2, 3))
    rows = np.nonzero(model.m_rel >= 0)[0]
    if len(rows) and model.rel.shape[1] >= 20:
        off[rows] = model.rel[model.m_rel[rows], 14:20].reshape(-1, 2, 3)
    return off


def rigid_arm_matrix(o1, o2):
    """E [12,12] of  u_end = u_node + theta_node x o  at both member ends (o in LOCAL axes):  the elastic axis runs
    between the offset points, each tied to its node by a rigid arm.  k_node = E^T k E,  p_node = E^T p.
    Semantics stated here, not pinned by any reference test (no solver source in the reference tree)."""
    E = np.eye(12)
    for e, o in ((0, o1), (6, o2)):
        # (theta x o) = [ty oz - tz oy,  tz ox - tx oz,  tx oy - ty ox]
        E[e + 0, e + 4] = o[2]; E[e + 0, e + 5] = -o[1]
        E[e + 1, e + 5] = o[0]; E[e + 1, e + 3] = -o[2]
        E[e + 2, e + 3] = o[1]; E[e + 2, e + 4] = -o[0]
    return E


def member_frames(xyz, n1, n2, roll, offsets=None):
    """Rotation matrices R [Nm,3,3] (rows = local x,y,z in global coordinates) and lengths.  With end offsets the
    axis runs from (start node + start offset) to (end node + end offset)."""
    d = xyz[n2] - xyz[n1]
    if offsets is not None:
        d = d + offsets[:, 1] - offsets[:, 0]
    L = np.sqrt((d * d).sum(1))
    if np.any(L < 1e-12):
        raise ValueError("zero-length member")
    lx = d / L[:, None]
    ref = np.tile(np.array([0.0, 1.0, 0.0]), (len(L), 1))
    vertical = np.abs(lx[:, 1]) > 1.0 - 1e-6
    ref[vertical] = np.array([0.0, 0.0, 1.0])
    lz = np.cross(lx, ref)
    lz /= np.sqrt((lz * lz).sum(1))[:, None]
    ly = np.cross(lz, lx)
    ly /= np.sqrt((ly * ly).sum(1))[:, None]
    c, s = np.cos(roll)[:, None], np.sin(roll)[:, None]
    rolled = roll != 0.0
    ly2 = np.where(rolled[:, None], c * ly + s * lz, ly)
    lz2 = np.where(rolled[:, None], -s * ly + c * lz, lz)
    return np.stack([lx, ly2, lz2], 1), L


def member_props(model: wire.ModelArrays):
    """Per-member [EA EIy EIz GJ GAsy GAsz]; rigid members get the stiff stand-in row."""
    prop = model.prop
    rigid = rigid_property_row(prop)
    table = np.vstack([prop, rigid[None, :]]) if prop.size else rigid[None, :]
    idx = np.where(model.m_prop >= 0, model.m_prop, table.shape[0] - 1)
    idx = np.where(model.m_type == wire.MT_RIGID, table.shape[0] - 1, idx)
    return table[idx]


def shear_phis(P, L, mtype):
    """(phi_y, phi_z): phi_y couples v/theta_z bending (EIz with G*a_sz)."""
    EIy, EIz, GAsy, GAsz = P[:, 1], P[:, 2], P[:, 4], P[:, 5]
    with np.errstate(divide="ignore", invalid="ignore"):
        phi_y = np.where(GAsz > 0.0, 12.0 * EIz / (GAsz * L * L), 0.0)
        phi_z = np.where(GAsy > 0.0, 12.0 * EIy / (GAsy * L * L), 0.0)
    return phi_y, phi_z


def local_stiffness(P, L, mtype):
    """12x12 local stiffness per member, DOF order [u v w rx ry rz] x 2 nodes."""
    nm = len(L)
    k = np.zeros((nm, 12, 12))
    EA, EIy, EIz, GJ = P[:, 0], P[:, 1], P[:, 2], P[:, 3]
    phi_y, phi_z = shear_phis(P, L, mtype)
    axial_only = mtype == wire.MT_TRUSS
    a = EA / L
    # linear St-Venant torsion GJ/L plus the theta_x part 12 EIw / L^3 of the cubic-Hermite warping matrix
    EIw = P[:, 7] if P.shape[1] > 7 else np.zeros(nm)
    t = np.where(axial_only, 0.0, GJ / L + 12.0 * EIw / L**3)
    for (i, j, s) in ((0, 0, 1), (0, 6, -1), (6, 0, -1), (6, 6, 1)):
        k[:, i, j] = s * a
    for (i, j, s) in ((3, 3, 1), (3, 9, -1), (9, 3, -1), (9, 9, 1)):
        k[:, i, j] = s * t
    # bending about local z: v (1,7), theta_z (5,11)
    bz = np.where(axial_only, 0.0, EIz / (1.0 + phi_y))
    k12, k6, k4, k2 = 12 * bz / L**3, 6 * bz / L**2, (4 + phi_y) * bz / L, (2 - phi_y) * bz / L
    v1, r1, v2, r2 = 1, 5, 7, 11
    blk = ((v1, v1, k12), (v1, r1, k6), (v1, v2, -k12), (v1, r2, k6),
           (r1, r1, k4), (r1, v2, -k6), (r1, r2, k2),
           (v2, v2, k12), (v2, r2, -k6), (r2, r2, k4))
    for (i, j, val) in blk:
        k[:, i, j] = val
        k[:, j, i] = val
    # bending about local y: w (2,8), theta_y (4,10)
    by = np.where(axial_only, 0.0, EIy / (1.0 + phi_z))
    k12, k6, k4, k2 = 12 * by / L**3, 6 * by / L**2, (4 + phi_z) * by / L, (2 - phi_z) * by / L
    w1, q1, w2, q2 = 2, 4, 8, 10
    blk = ((w1, w1, k12), (w1, q1, -k6), (w1, w2, -k12), (w1, q2, -k6),
           (q1, q1, k4), (q1, w2, k6), (q1, q2, k2),
           (w2, w2, k12), (w2, q2, k6), (q2, q2, k4))
    for (i, j, val) in blk:
        k[:, i, j] = val
        k[:, j, i] = val
    return k


def geometric_stiffness(N, L, formulation, mtype=None):
    """Local geometric stiffness for axial force N (tension positive).

    CONSISTENT = Przemieniecki string+bending terms; SIMPLIFIED = N/L on the transverse
    translations only (pydantic_models.py:1377-1380).
    """
    nm = len(L)
    g = np.zeros((nm, 12, 12))
    if formulation == "SIMPLIFIED":
        c = N / L
        for (i, j) in ((1, 7), (2, 8)):
            g[:, i, i] = c; g[:, j, j] = c; g[:, i, j] = -c; g[:, j, i] = -c
        return g
    c = N / (30.0 * L)
    v1, r1, v2, r2 = 1, 5, 7, 11
    for (i, j, val) in ((v1, v1, 36.0), (v1, r1, 3 * L), (v1, v2, -36.0), (v1, r2, 3 * L),
                        (r1, r1, 4 * L * L), (r1, v2, -3 * L), (r1, r2, -L * L),
                        (v2, v2, 36.0), (v2, r2, -3 * L), (r2, r2, 4 * L * L)):
        g[:, i, j] = c * val
        g[:, j, i] = c * val
    w1, q1, w2, q2 = 2, 4, 8, 10
    for (i, j, val) in ((w1, w1, 36.0), (w1, q1, -3 * L), (w1, w2, -36.0), (w1, q2, -3 * L),
                        (q1, q1, 4 * L * L), (q1, w2, 3 * L), (q1, q2, -L * L),
                        (w2, w2, 36.0), (w2, q2, 3 * L), (q2, q2, 4 * L * L)):
        g[:, i, j] = c * val
        g[:, j, i] = c * val
    if mtype is not None:
        # truss members keep only the string terms (no rotational stiffness to couple to)
        tr = mtype == wire.MT_TRUSS
        if np.any(tr):
            gs = geometric_stiffness(N[tr], L[tr], "SIMPLIFIED")
            g[tr] = gs
    return g


def condense_releases(k, p, rel):
    """Static condensation of end-release springs (pydantic_models.py:2062-2087).

    For each local end DOF d with release spring ks >= 0 (ks = 0: full release) the member end
    is connected to the node through ks:  k <- k - k[:,d] k[d,:] / (k[d,d] + ks), same for p.
    ``p`` is [nrhs,12] equivalent nodal load rows (may be empty).
    """
    k = k.copy()
    p = p.copy()
    for d in range(12):
        ks = rel[d]
        if ks < 0.0:
            continue
        den = k[d, d] + ks
        if den <= 0.0:
            continue
        col = k[:, d].copy()
        if p.size:
            p -= np.outer(p[:, d] / den, col)
        k -= np.outer(col, col) / den
    return k, p


def shear_centre_transform(k, p, ys, zs):
    """Shear-centre eccentricity coupling (pydantic_models.py:1106-1111 `y_s`/`z_s`, toggle 1358-1361:
    "T^T.K.T transformation").  The beam's transverse displacements and twist are those of its shear-centre
    axis; the nodes are on the centroid axis.  For a rigid cross-section rotating by theta_x the point
    (y_s, z_s) moves by theta_x * (-z_s, +y_s), so per end  v_sc = v - z_s rx,  w_sc = w + y_s rx:
    k <- S^T k S and p <- S^T p (rows of ``p`` [nrhs,12]).  Parity unpinned (no reference number)."""
    S = np.eye(12)
    for e in (0, 6):
        S[e + 1, e + 3] = -zs
        S[e + 2, e + 3] = ys
    return S.T @ k @ S, (p @ S if p.size else p.copy())


def product_inertia_coupling(eiyz, L):
    """E I_yz int v'' w'' dx between the bending planes of an Euler-Bernoulli member (Section.i_yz,
    pydantic_models.py:1078-1082; wire.section_bending_inertias): v over (v1, rz1, v2, rz2) with the Hermite cubics,
    w over (w1, ry1, w2, ry2) with ry = -w'."""
    B = np.array([[12.0, 6 * L, -12.0, 6 * L], [6 * L, 4 * L * L, -6 * L, 2 * L * L],
                  [-12.0, -6 * L, 12.0, -6 * L], [6 * L, 2 * L * L, -6 * L, 4 * L * L]]) * (eiyz / L ** 3)
    vs, ws, sw = (1, 5, 7, 11), (2, 4, 8, 10), (1.0, -1.0, 1.0, -1.0)
    k = np.zeros((12, 12))
    for i in range(4):
        for j in range(4):
            k[vs[i], ws[j]] += B[i, j] * sw[j]
            k[ws[j], vs[i]] += B[i, j] * sw[j]
    return k


def member_modifiers(k, p, rec, L=None):
    """One row of ``ModelArrays.rel``: product of inertia, eccentricity, then the end-release condensation."""
    if len(rec) > 20 and rec[20] != 0.0:
        k = k + product_inertia_coupling(rec[20], L)
    if len(rec) > 12 and (rec[12] != 0.0 or rec[13] != 0.0):
        k, p = shear_centre_transform(k, p, rec[12], rec[13])
    return condense_releases(k, p, rec[:12])


def expand_T(R):
    """[Nm,12,12] block-diagonal transformation from R [Nm,3,3]."""
    nm = R.shape[0]
    T = np.zeros((nm, 12, 12))
    for b in range(4):
        T[:, 3 * b:3 * b + 3, 3 * b:3 * b + 3] = R
    return T


# --------------------------------------------------------------------------------------
# plate element: CST membrane + DSG3 Mindlin bending
# --------------------------------------------------------------------------------------
def plate_frame(x3, theta=0.0):
    """Local frame of a flat triangle: rows = local x,y,z.  x3 [3,3] corner coordinates.  ``theta`` turns local x, y
    about the normal from the default frame (PlateElement.local_x_direction, resolved to an angle by the parser).

    local z = normal (right-hand rule over the wire node order); local x = global X projected
    into the plane (global Y when the normal is parallel to X), so an XY-plane plate has local =
    global as in every pinned case.
    """
    n = np.cross(x3[1] - x3[0], x3[2] - x3[0])
    nn = np.linalg.norm(n)
    if nn < 1e-300:
        raise ValueError("degenerate plate element")
    ez = n / nn
    gx = np.array([1.0, 0.0, 0.0])
    if abs(ez @ gx) > 1.0 - 1e-6:
        gx = np.array([0.0, 1.0, 0.0])
    ex = gx - (gx @ ez) * ez
    ex /= np.linalg.norm(ex)
    ey = np.cross(ez, ex)
    if theta != 0.0:
        c, s = np.cos(theta), np.sin(theta)
        ex, ey = c * ex + s * ey, -s * ex + c * ey
    return np.stack([ex, ey, ez])


def plate_offset_matrix(offset: float) -> np.ndarray:
    """E with u_mid = E u_node for a plate whose mid-surface sits `offset` above its nodes along the local normal
    (PlateElement.offset, pydantic_models.py:1711-1714): a point at height z of a rotating normal moves by theta x (0,0,z),
    i.e. (+z ry, -z rx, 0) in local axes."""
    E = np.eye(18)
    for n in range(3):
        arm = np.array([0.0, 0.0, offset])
        for c in range(3):                                   # unit rotation about local axis c at node n
            rot = np.zeros(3); rot[c] = 1.0
            E[6 * n:6 * n + 3, 6 * n + 3 + c] = np.cross(rot, arm)
    return E


def dkt_curvature(xl, xi: float, eta: float) -> np.ndarray:
    """Curvature operator [3,9] of the discrete Kirchhoff triangle (PlateTheory.Kirchhoff, pydantic_models.py:816-819) at the
    area coordinates (xi, eta) = (L2, L3), DOFs (w, bx, by) per corner, built from the element's DEFINITION rather than
    from the published coefficient lists the CUDA path uses: the rotations of the normal are quadratic (corner + mid-side
    values); at a corner they are the nodal ones; at the middle of a side the tangential one is minus the slope of the
    cubic w along the side, -w,s = 3/(2l) (w_i - w_j) - (b_s,i + b_s,j)/4, the normal one the mean of the two ends."""
    X = np.asarray(xl, float)
    (x1, y1), (x2, y2), (x3, y3) = X
    twoA = (x2 - x1) * (y3 - y1) - (x3 - x1) * (y2 - y1)
    dLx = np.array([y2 - y3, y3 - y1, y1 - y2]) / twoA
    dLy = np.array([x3 - x2, x1 - x3, x2 - x1]) / twoA
    H = np.zeros((6, 2, 9))                                  # rotation (bx, by) of the six nodes in the nine DOFs
    for i in range(3):
        H[i, 0, 3 * i + 1] = 1.0
        H[i, 1, 3 * i + 2] = 1.0
    sides = [(1, 2), (2, 0), (0, 1)]
    for k, (i, j) in enumerate(sides):
        d = X[j] - X[i]
        l = float(np.hypot(*d))
        s_ = d / l
        n_ = np.array([s_[1], -s_[0]])
        bs = np.zeros(9); bn = np.zeros(9)
        bs[3 * i] += 1.5 / l; bs[3 * j] -= 1.5 / l
        for nd in (i, j):
            bs[3 * nd + 1] -= 0.25 * s_[0]; bs[3 * nd + 2] -= 0.25 * s_[1]
            bn[3 * nd + 1] += 0.5 * n_[0]; bn[3 * nd + 2] += 0.5 * n_[1]
        H[3 + k, 0] = bs * s_[0] + bn * n_[0]
        H[3 + k, 1] = bs * s_[1] + bn * n_[1]
    L = np.array([1.0 - xi - eta, xi, eta])
    dN = np.zeros((6, 3))                                    # dN_a / dL_i of the quadratic shape functions
    for i in range(3):
        dN[i, i] = 4.0 * L[i] - 1.0
    for k, (i, j) in enumerate(sides):
        dN[3 + k, i] = 4.0 * L[j]; dN[3 + k, j] = 4.0 * L[i]
    nx, ny = dN @ dLx, dN @ dLy
    B = np.zeros((3, 9))
    for a in range(6):
        B[0] += nx[a] * H[a, 0]
        B[1] += ny[a] * H[a, 1]
        B[2] += ny[a] * H[a, 0] + nx[a] * H[a, 1]
    return B


def plate_local_matrices(xl, t, E, nu, flags, mods=(1.0, 1.0, 1.0), ortho=None):
    """18x18 local stiffness of a TRI3 shell, plus (Bm, Bb, Bs, Dm, Db, Ds, area).

    ortho = (E_x, E_y, G_xy, nu_xy, G_xz, G_yz): OrthotropicPlateMaterial (pydantic_models.py:663-681), material axes =
    local x, y; the compliance form is inverted here (the CUDA path writes the stiffness form out).

    Local DOFs per node [u v w rx ry rz]; DSG3 uses (w, bx, by) with bx = +ry, by = -rx
    (SURVEY.md B.3 note).  Drilling rz gets a small stabilising stiffness.
    """
    (x1, y1), (x2, y2), (x3, y3) = xl
    a_, b_, c_, d_ = x2 - x1, y2 - y1, y3 - y1, x3 - x1
    A = 0.5 * (a_ * c_ - b_ * d_)
    twoA = 2.0 * A
    k = np.zeros((18, 18))
    # ---- CST membrane (B.1)
    bb = np.array([y2 - y3, y3 - y1, y1 - y2])
    cc = np.array([x3 - x2, x1 - x3, x2 - x1])
    Bm = np.zeros((3, 6))
    Bm[0, 0::2] = bb
    Bm[1, 1::2] = cc
    Bm[2, 0::2] = cc
    Bm[2, 1::2] = bb
    Bm /= twoA
    if flags & 4:      # plane strain (PlateElement.plane_state, pydantic_models.py:1712-1715)
        Dm = E / ((1 + nu) * (1 - 2 * nu)) * np.array([[1 - nu, nu, 0], [nu, 1 - nu, 0], [0, 0, (1 - 2 * nu) / 2]])
    else:
        Dm = E / (1.0 - nu * nu) * np.array([[1, nu, 0], [nu, 1, 0], [0, 0, (1 - nu) / 2]])
    D_ortho = None
    if ortho is not None:
        ex, ey, gxy, nxy, gxz, gyz = ortho
        compliance = np.array([[1.0 / ex, -nxy / ex, 0.0], [-nxy / ex, 1.0 / ey, 0.0], [0.0, 0.0, 1.0 / gxy]])   # eps = C sigma
        D_ortho = np.linalg.inv(compliance)
        Dm = D_ortho.copy()
    Dm = mods[0] * Dm                                  # PlateElement.stiffness_modifiers (801-807)
    mdof = np.array([0, 1, 6, 7, 12, 13])
    if flags & 1:
        km = t * abs(A) * Bm.T @ Dm @ Bm
        k[np.ix_(mdof, mdof)] += km
    # ---- DSG3 bending (B.3), dofs (w, bx, by) per node
    Bb = np.array([[0, b_ - c_, 0, 0, c_, 0, 0, -b_, 0],
                   [0, 0, d_ - a_, 0, 0, -d_, 0, 0, a_],
                   [0, d_ - a_, b_ - c_, 0, -d_, c_, 0, a_, -b_]]) / twoA
    Bs = np.array([[b_ - c_, A, 0, c_, a_ * c_ / 2, b_ * c_ / 2, -b_, -b_ * d_ / 2, -b_ * c_ / 2],
                   [d_ - a_, 0, A, -d_, -a_ * d_ / 2, -b_ * d_ / 2, a_, a_ * d_ / 2, a_ * c_ / 2]]) / twoA
    Db = mods[1] * E * t**3 / (12.0 * (1.0 - nu * nu)) * np.array([[1, nu, 0], [nu, 1, 0], [0, 0, (1 - nu) / 2]])
    G = E / (2.0 * (1.0 + nu))
    Ds = mods[2] * (5.0 / 6.0) * G * t * np.eye(2)
    if ortho is not None:
        Db = mods[1] * t**3 / 12.0 * D_ortho
        Ds = mods[2] * (5.0 / 6.0) * t * np.diag([ortho[4], ortho[5]])
    if flags & 2 and flags & 8:
        # thin plate (Kirchhoff): DKT bending, three mid-side integration points, no shear energy.  The results take the
        # curvature at the centroid; shear forces are not formed (they are derivatives of the moments, which one DKT element
        # does not resolve: W_KIRCHHOFF_NO_SHEAR_FORCES)
        kb = sum(abs(A) / 3.0 * (B.T @ Db @ B) for B in (dkt_curvature(xl, *gp) for gp in ((0.5, 0.5), (0.0, 0.5), (0.5, 0.0))))
        Bb = dkt_curvature(xl, 1.0 / 3.0, 1.0 / 3.0)
        Bs = np.zeros((2, 9))
    elif flags & 2:
        kb = abs(A) * (Bb.T @ Db @ Bb + Bs.T @ Ds @ Bs)
    if flags & 2:
        # map (w, bx, by) -> (w, rx, ry): bx = ry, by = -rx
        S = np.zeros((9, 18))
        for n in range(3):
            S[3 * n + 0, 6 * n + 2] = 1.0
            S[3 * n + 1, 6 * n + 4] = 1.0
            S[3 * n + 2, 6 * n + 3] = -1.0
        k += S.T @ kb @ S
    # ---- drilling stabilisation: small fraction of the membrane shear stiffness
    kd = 1e-6 * E * t * abs(A)
    for n in range(3):
        k[6 * n + 5, 6 * n + 5] += kd
    return k, (Bm, Bb, Bs, Dm, Db, Ds, A)


def plate_stiffness_global(model, xyz):
    """Per-plate (18x18 global stiffness, frame R, local corner coords)."""
    out = []
    for e in range(model.n_plates):
        x3 = xyz[model.p_nodes[e]]
        R = plate_frame(x3, float(model.p_data[e, wire.PD_THETA]))
        xl = (x3 - x3[0]) @ R.T
        t, E, nu = model.p_data[e, :3]
        ortho = tuple(model.p_data[e, 8:14]) if model.p_data.shape[1] > 8 and model.p_data[e, 7] != 0.0 else None
        k, mats = plate_local_matrices(xl[:, :2], t, E, nu, int(model.p_flags[e]), tuple(model.p_data[e, 3:6]), ortho)
        off = float(model.p_data[e, 14]) if model.p_data.shape[1] > 14 else 0.0
        if off != 0.0:                                       # nodes on the reference surface, element on the mid-surface
            Eo = plate_offset_matrix(off)
            k = Eo.T @ k @ Eo
        T = np.zeros((18, 18))
        for b in range(6):
            T[3 * b:3 * b + 3, 3 * b:3 * b + 3] = R
        out.append((T.T @ k @ T, R, xl[:, :2], mats, T))
    return out


# --------------------------------------------------------------------------------------
# symbolic pattern (bit-exact contract with the CUDA path)
# --------------------------------------------------------------------------------------
def bsr_pattern(model: wire.ModelArrays):
    """Node-block sparsity: sorted unique (row,col) node pairs, every diagonal present.

    Returns (rowptr int32 [Nn+1], colidx int32 [nnzb]).
    """
    nn = model.n_nodes
    rows = [np.arange(nn, dtype=np.int64)]
    cols = [np.arange(nn, dtype=np.int64)]
    a, b = model.m_n1.astype(np.int64), model.m_n2.astype(np.int64)
    rows += [a, b]
    cols += [b, a]
    if model.n_plates:
        pn = model.p_nodes.astype(np.int64)
        for i in range(3):
            for j in range(3):
                if i != j:
                    rows.append(pn[:, i]); cols.append(pn[:, j])
    key = np.unique(np.concatenate(rows) * nn + np.concatenate(cols))
    r, c = key // nn, key % nn
    rowptr = np.zeros(nn + 1, np.int64)
    np.add.at(rowptr, r + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr.astype(np.int32), c.astype(np.int32)


def bc_masks(model: wire.ModelArrays):
    """(fixed mask uint8 [Ndof], free-DOF numbering int32 [Ndof] (-1 constrained))."""
    code = model.bc_code.reshape(-1)
    fixed = (code == wire.BC_FIXED).astype(np.uint8)
    free_index = np.where(fixed == 0, np.cumsum(1 - fixed.astype(np.int64)) - 1, -1).astype(np.int32)
    return fixed, free_index


# --------------------------------------------------------------------------------------
# assembly
# --------------------------------------------------------------------------------------
class Assembly:
    """Everything geometry-dependent for one coordinate set (one imperfection case)."""

    def __init__(self, model: wire.ModelArrays, xyz=None, m_scale=None):
        self.model = model
        self.xyz = model.xyz if xyz is None else xyz
        nm = model.n_members
        self.m_scale = np.ones(nm) if m_scale is None else m_scale
        if nm:
            self.offsets = member_offsets(model)
            self.R, self.L = member_frames(self.xyz, model.m_n1, model.m_n2, model.m_roll, self.offsets)
            self.P = member_props(model)
            # the per-member multiplier (active-set switch) scales the stiffness BEFORE release condensation
            self.k_local = local_stiffness(self.P, self.L, model.m_type) * self.m_scale[:, None, None]
            self.phi_y, self.phi_z = shear_phis(self.P, self.L, model.m_type)
        else:
            self.R = np.zeros((0, 3, 3)); self.L = np.zeros(0); self.P = np.zeros((0, 8))
            self.k_local = np.zeros((0, 12, 12)); self.phi_y = self.phi_z = np.zeros(0)
        self.T = expand_T(self.R)
        self.k_cond = self.k_local.copy()       # after release condensation
        for i in np.nonzero(model.m_rel >= 0)[0]:
            self.k_cond[i], _ = member_modifiers(self.k_local[i], np.zeros((0, 12)), model.rel[model.m_rel[i]], self.L[i])
        # rigid arms of the end offsets: E per member (local axes), applied after the release condensation
        self.arms = {}
        if nm:
            for i in np.nonzero(np.abs(self.offsets).reshape(nm, -1).max(1) > 0.0)[0]:
                E = rigid_arm_matrix(self.R[i] @ self.offsets[i, 0], self.R[i] @ self.offsets[i, 1])
                self.arms[int(i)] = E
                self.k_cond[i] = E.T @ self.k_cond[i] @ E
        self.plates = plate_stiffness_global(model, self.xyz) if model.n_plates else []
        self.edof = np.concatenate([6 * model.m_n1[:, None] + np.arange(6), 6 * model.m_n2[:, None] + np.arange(6)], 1) if nm else np.zeros((0, 12), np.int64)

    def warping_coefficients(self):
        """Per warping member: (theta_x direction lx [Nw,3], c6 = 6 EIw / L^2, c4 = 4 EIw / L, c2 = 2 EIw / L), scaled."""
        model = self.model
        w = model.w_member
        L = self.L[w]
        eiw = self.P[w, 7] * self.m_scale[w]
        return self.R[w, 0, :], 6.0 * eiw / L**2, 4.0 * eiw / L, 2.0 * eiw / L

    def warping_triplets(self):
        """Coupling of the warping DOFs psi (DOF 0 of the auxiliary nodes) with the twist of the member ends:
        cubic-Hermite matrix EIw/L^3 [[12,6L,-12,6L],[6L,4L^2,-6L,2L^2],[-12,-6L,12,-6L],[6L,2L^2,-6L,4L^2]] on
        (theta_x1, psi1, theta_x2, psi2); its theta-theta part sits in the 12x12 torsion term."""
        model = self.model
        w = model.w_member
        lx, c6, c4, c2 = self.warping_coefficients()
        s, e = model.m_n1[w], model.m_n2[w]
        p1, p2 = 6 * model.w_aux[:, 0].astype(np.int64), 6 * model.w_aux[:, 1].astype(np.int64)
        rows, cols, vals = [], [], []

        def add(r, c, v):
            rows.extend([r, c]); cols.extend([c, r]); vals.extend([v, v])
        for i in range(3):
            t1, t2 = 6 * s.astype(np.int64) + 3 + i, 6 * e.astype(np.int64) + 3 + i
            add(t1, p1, c6 * lx[:, i]); add(t1, p2, c6 * lx[:, i])
            add(t2, p1, -c6 * lx[:, i]); add(t2, p2, -c6 * lx[:, i])
        add(p1, p2, c2)
        rows.extend([p1, p2]); cols.extend([p1, p2]); vals.extend([c4, c4])
        return np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)

    def K_elastic(self, kg_local=None):
        """Global stiffness (no supports, no springs) as CSR [Ndof,Ndof]."""
        model = self.model
        nd = model.n_dof
        rows, cols, vals = [], [], []
        if model.n_members:
            kl = self.k_cond
            if kg_local is not None:
                kl = kl + kg_local
            kg = np.einsum("mji,mjk,mkl->mil", self.T, kl, self.T)
            rows.append(np.repeat(self.edof, 12, axis=1).reshape(-1))
            cols.append(np.tile(self.edof, (1, 12)).reshape(-1))
            vals.append(kg.reshape(-1))
        for e, (kge, *_rest) in enumerate(self.plates):
            ed = (6 * model.p_nodes[e][:, None] + np.arange(6)).reshape(-1)
            rows.append(np.repeat(ed, 18)); cols.append(np.tile(ed, 18)); vals.append(kge.reshape(-1))
        if len(model.w_member):
            r_, c_, v_ = self.warping_triplets()
            rows.append(r_); cols.append(c_); vals.append(v_)
        if not rows:
            return sp.csr_matrix((nd, nd))
        K = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nd, nd))
        return K.tocsr()


def shape_integrals(a, b, q1, q2, phi, L):
    """Integrals of a linear load q(xi) (q1 at a, q2 at b) against the Timoshenko shapes.

    Returns (S1, S2, S3, S4, lin0, lin1): int N_i q L dxi and int (1-xi) q L dxi, int xi q L dxi.
    3-point Gauss is exact (integrand degree 4).
    """
    half = 0.5 * (b - a)
    mid = 0.5 * (a + b)
    out = np.zeros(6)
    if half <= 0.0:
        return out
    for gx, gw in zip(GAUSS3_X, GAUSS3_W):
        xi = mid + half * gx
        q = q1 + (q2 - q1) * (xi - a) / (b - a)
        w = gw * half * L * q
        n1, n2, n3, n4 = hermite(xi, phi, L)
        out += w * np.array([n1, n2, n3, n4, 1.0 - xi, xi])
    return out


def hermite(xi, phi, L):
    d = 1.0 + phi
    n1 = (1 - 3 * xi**2 + 2 * xi**3 + phi * (1 - xi)) / d
    n2 = L * (xi - 2 * xi**2 + xi**3 + phi * (xi - xi**2) / 2) / d
    n3 = (3 * xi**2 - 2 * xi**3 + phi * xi) / d
    n4 = L * (-xi**2 + xi**3 - phi * (xi - xi**2) / 2) / d
    return n1, n2, n3, n4


def hermite_rot(xi, phi, L):
    d = 1.0 + phi
    g1 = 6 * (xi**2 - xi) / (L * d)
    g2 = (1 - 4 * xi + 3 * xi**2 + phi * (1 - xi)) / d
    g3 = -g1
    g4 = (3 * xi**2 - 2 * xi + phi * xi) / d
    return g1, g2, g3, g4


def equivalent_loads(asm: Assembly, lc: wire.LoadCaseArrays):
    """Local equivalent nodal loads p_eq [Nm,12] of one load case (before release condensation)."""
    model = asm.model
    p = np.zeros((model.n_members, 12))
    for mi, row in zip(lc.dl_member, lc.dl_data):
        R, L = asm.R[mi], asm.L[mi]
        q1, q2, a, b = R @ row[0:3], R @ row[3:6], row[6], row[7]
        truss = model.m_type[mi] == wire.MT_TRUSS
        # axial
        s = shape_integrals(a, b, q1[0], q2[0], 0.0, L)
        p[mi, 0] += s[4]; p[mi, 6] += s[5]
        sy = shape_integrals(a, b, q1[1], q2[1], asm.phi_y[mi], L)
        sz = shape_integrals(a, b, q1[2], q2[2], asm.phi_z[mi], L)
        if truss:
            p[mi, 1] += sy[4]; p[mi, 7] += sy[5]; p[mi, 2] += sz[4]; p[mi, 8] += sz[5]
        else:
            p[mi, 1] += sy[0]; p[mi, 5] += sy[1]; p[mi, 7] += sy[2]; p[mi, 11] += sy[3]
            p[mi, 2] += sz[0]; p[mi, 4] -= sz[1]; p[mi, 8] += sz[2]; p[mi, 10] -= sz[3]
    for mi, row in zip(lc.pl_member, lc.pl_data):
        R, L = asm.R[mi], asm.L[mi]
        F, M, xi, is_local = row[0:3], row[3:6], row[6], row[7] != 0.0
        if not is_local:
            F, M = R @ F, R @ M
        truss = model.m_type[mi] == wire.MT_TRUSS
        p[mi, 0] += (1 - xi) * F[0]; p[mi, 6] += xi * F[0]
        p[mi, 3] += (1 - xi) * M[0]; p[mi, 9] += xi * M[0]
        if truss:
            p[mi, 1] += (1 - xi) * F[1]; p[mi, 7] += xi * F[1]
            p[mi, 2] += (1 - xi) * F[2]; p[mi, 8] += xi * F[2]
            continue
        n = hermite(xi, asm.phi_y[mi], L); g = hermite_rot(xi, asm.phi_y[mi], L)
        for dof, nv, gv in zip((1, 5, 7, 11), n, g):
            p[mi, dof] += nv * F[1] + gv * M[2]
        n = hermite(xi, asm.phi_z[mi], L); g = hermite_rot(xi, asm.phi_z[mi], L)
        for dof, nv, gv, sg in zip((2, 4, 8, 10), n, g, (1.0, -1.0, 1.0, -1.0)):
            # w = N1 w1 - N2 ry1 + N3 w2 - N4 ry2 ; ry = -G1 w1 + G2 ry1 - G3 w2 + G4 ry2
            p[mi, dof] += sg * nv * F[2] + (-sg) * gv * M[1]
    return p


def plate_loads(asm: Assembly, lc: wire.LoadCaseArrays, F):
    """Lump plate pressures: q*A/3 on each corner's translations."""
    model = asm.model
    for e, vec in zip(lc.pp_plate, lc.pp_vec):
        A = abs(asm.plates[e][3][6])
        for n in model.p_nodes[e]:
            F[6 * n:6 * n + 3] += vec * A / 3.0


def load_vectors(asm: Assembly):
    """Per load case: global nodal load vector F [Nlc,Ndof] and local p_eq [Nlc,Nm,12]
    (p_eq already condensed for end releases)."""
    model = asm.model
    nlc = len(model.load_cases)
    F = np.zeros((nlc, model.n_dof))
    peq = np.zeros((nlc, model.n_members, 12))
    for li, lc in enumerate(model.load_cases):
        for n, v in zip(lc.nod_node, lc.nod_vec):
            F[li, 6 * n:6 * n + 6] += v
        p = equivalent_loads(asm, lc)
        for i in np.nonzero(model.m_rel >= 0)[0]:
            _, pc = member_modifiers(asm.k_local[i], p[i][None, :], model.rel[model.m_rel[i]], asm.L[i])
            p[i] = pc[0]
        for i, E in asm.arms.items():
            p[i] = E.T @ p[i]
        p *= 1.0  # member scale does not apply to loads
        peq[li] = p
        if model.n_members:
            pg = np.einsum("mji,mj->mi", asm.T, p)
            np.add.at(F[li], asm.edof.reshape(-1), pg.reshape(-1))
        if len(lc.pp_plate):
            plate_loads(asm, lc, F[li])
    return F, peq


# --------------------------------------------------------------------------------------
# solve + recovery
# --------------------------------------------------------------------------------------
def constrain(K, model, extra_fixed=None, bc=None):
    """Add support springs, eliminate fixed DOFs. Returns (K_ff factor-ready csc, free idx).
    bc = (code, k) overrides the model's support codes (active set of the unilateral supports)."""
    nd = model.n_dof
    code = (model.bc_code if bc is None else bc[0]).reshape(-1)
    kspring = np.where(code == wire.BC_SPRING, (model.bc_k if bc is None else bc[1]).reshape(-1), 0.0)
    fixed = code == wire.BC_FIXED
    if extra_fixed is not None:
        fixed = fixed | extra_fixed
    Ks = K + sp.diags(kspring)
    # zero-stiffness free DOFs (e.g. rotations of truss-only nodes) are restrained automatically
    diag = Ks.diagonal()
    dead = (~fixed) & (diag == 0.0)
    fixed = fixed | dead
    free = np.nonzero(~fixed)[0]
    return Ks[free][:, free].tocsc(), free, dead


SOLVE_THREADS = max(1, os.cpu_count() or 1)      # host threads of the multi-right-hand-side substitution


class NotPositiveDefinite(RuntimeError):
    """The tangent matrix K + K_g of a second-order iteration is not positive definite: the structure has buckled
    under the applied load (the engine raises EngineError "... not positive definite ..." at the same point)."""


def solve_sym(Kff, B, require_pd=False):
    """Direct solve of K_ff X = B (B [n,nrhs]).  require_pd: refuse an indefinite matrix (second-order iterations)
    instead of solving it by LU."""
    if Kff.shape[0] == 0:
        return np.zeros_like(B)
    if Kff.shape[0] <= 3000:
        import scipy.linalg as sla
        A = Kff.toarray()
        try:
            c = sla.cho_factor(A, lower=True, check_finite=False)
            return sla.cho_solve(c, B, check_finite=False)
        except np.linalg.LinAlgError:
            if require_pd:
                raise NotPositiveDefinite("stiffness matrix is not positive definite: buckled under the applied load") from None
            return np.linalg.solve(A, B)
    lu = spla.splu(Kff, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    if require_pd and not (lu.U.diagonal() > 0.0).all():      # no pivoting: U carries the pivots of the LDL^T
        raise NotPositiveDefinite("stiffness matrix is not positive definite: buckled under the applied load")
    nrhs = B.shape[1] if B.ndim == 2 else 1
    threads = min(SOLVE_THREADS, nrhs)
    if threads <= 1 or B.ndim != 2:
        return lu.solve(B)
    # SuperLU substitutes one column at a time and releases the GIL: independent column chunks on host threads
    # (the factorisation itself is sequential).  Same arithmetic per column as the single call.
    from concurrent.futures import ThreadPoolExecutor
    bounds = [nrhs * k // threads for k in range(threads + 1)]
    X = np.empty_like(B)

    def work(k):
        X[:, bounds[k]:bounds[k + 1]] = lu.solve(np.ascontiguousarray(B[:, bounds[k]:bounds[k + 1]]))
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(threads)))
    return X


def recover_members(asm: Assembly, U, peq, kg_local=None):
    """Member forces for displacement columns U [nrhs,Ndof], p_eq [nrhs,Nm,12].

    Returns dict of arrays (see fers_core_b200.results.ResultArrays).
    """
    model = asm.model
    nrhs, nm = U.shape[0], model.n_members
    out = {}
    if nm == 0:
        z = lambda *s: np.zeros((nrhs, 0) + s)
        return dict(f_local=z(12), f_global=z(12), u_local=z(12), sec=z(len(SECTION_FRACS), 6),
                    lmax=z(6), lmin=z(6), gmax=z(6), gmin=z(6), bw=z(2))
    ue = U[:, asm.edof]                                   # [nrhs,Nm,12]
    ul = np.einsum("mij,rmj->rmi", asm.T, ue)
    kl = asm.k_cond
    if kg_local is not None:
        kl = kl[None] + kg_local                          # [nrhs,Nm,12,12]
        fl = np.einsum("rmij,rmj->rmi", kl, ul) - peq
    else:
        fl = np.einsum("mij,rmj->rmi", kl, ul) - peq
    bw = np.zeros((nrhs, nm, 2))
    if len(model.w_member):
        # warping members: psi-coupling of the end torques and the bimoments (nodal force k u at the psi DOF)
        w = model.w_member
        _, c6, c4, c2 = asm.warping_coefficients()
        psi1, psi2 = U[:, 6 * model.w_aux[:, 0]], U[:, 6 * model.w_aux[:, 1]]      # [nrhs,Nw]
        th1, th2 = ul[:, w, 3], ul[:, w, 9]
        fl[:, w, 3] += c6 * (psi1 + psi2)
        fl[:, w, 9] -= c6 * (psi1 + psi2)
        bw[:, w, 0] = c6 * (th1 - th2) + c4 * psi1 + c2 * psi2
        bw[:, w, 1] = c6 * (th1 - th2) + c2 * psi1 + c4 * psi2
    fg = np.einsum("mji,rmj->rmi", asm.T, fl)
    out["f_local"], out["f_global"], out["u_local"], out["bw"] = fl, fg, ul, bw
    return out


def span_load_tables(asm: Assembly):
    """Per load case, per member: list of local span loads for section-force transport."""
    model = asm.model
    tabs = []
    for lc in model.load_cases:
        per = {}
        for mi, row in zip(lc.dl_member, lc.dl_data):
            R = asm.R[mi]
            per.setdefault(int(mi), []).append(("d", R @ row[0:3], R @ row[3:6], row[6], row[7]))
        for mi, row in zip(lc.pl_member, lc.pl_data):
            R = asm.R[mi]
            F, M = row[0:3], row[3:6]
            if row[7] == 0.0:
                F, M = R @ F, R @ M
            per.setdefault(int(mi), []).append(("p", F, M, row[6]))
        tabs.append(per)
    return tabs


def span_resultant(loads, L, x):
    """Resultant (force[3], moment[3] about the section point) of span loads on [0, x*L]."""
    f = np.zeros(3); m = np.zeros(3)
    for ld in loads:
        if ld[0] == "d":
            _, q1, q2, a, b = ld
            hi = min(b, x)
            if hi <= a:
                continue
            half, mid = 0.5 * (hi - a), 0.5 * (hi + a)
            for gx, gw in zip(GAUSS3_X, GAUSS3_W):
                xi = mid + half * gx
                q = q1 + (q2 - q1) * ((xi - a) / (b - a))
                w = gw * half * L
                d = (x - xi) * L
                f += w * q
                m += w * np.array([0.0, d * q[2], -d * q[1]])
        else:
            _, F, M, pos = ld
            if pos <= x:
                d = (x - pos) * L
                f += F
                m += M + np.array([0.0, d * F[2], -d * F[1]])
    return f, m


def section_forces(asm: Assembly, fl, span_by_rhs):
    """Internal forces at SECTION_FRACS: start force transported along the member.

    fl [nrhs,Nm,12]; span_by_rhs: per rhs a list of (factor, span table) pairs.
    """
    nrhs, nm = fl.shape[:2]
    ns = len(SECTION_FRACS)
    sec = np.zeros((nrhs, nm, ns, 6))
    if asm.arms:
        # forces at the start of the elastic axis from the forces at the node: m_axis = m_node - o x f
        fl = fl.copy()
        for i in asm.arms:
            o = asm.R[i] @ asm.offsets[i, 0]
            fl[:, i, 3:6] -= np.cross(o[None, :], fl[:, i, 0:3])
    for si, x in enumerate(SECTION_FRACS):
        d = x * asm.L
        sec[:, :, si, 0:3] = fl[:, :, 0:3]
        sec[:, :, si, 3] = fl[:, :, 3]
        sec[:, :, si, 4] = fl[:, :, 4] + d[None, :] * fl[:, :, 2]
        sec[:, :, si, 5] = fl[:, :, 5] - d[None, :] * fl[:, :, 1]
    for r in range(nrhs):
        for fac, tab in span_by_rhs[r]:
            if fac == 0.0:
                continue
            for mi, loads in tab.items():
                for si, x in enumerate(SECTION_FRACS):
                    f, m = span_resultant(loads, asm.L[mi], x)
                    sec[r, mi, si, 0:3] += fac * f
                    sec[r, mi, si, 3:6] += fac * m
    return sec


def min_max(asm: Assembly, fl, fg, sec):
    """Per-component extrema over {start, end (nodal convention), interior sections}."""
    loc = np.concatenate([fl[:, :, None, 0:6], fl[:, :, None, 6:12], sec], axis=2)
    secg = np.concatenate([np.einsum("mji,rmsj->rmsi", asm.R, sec[..., 0:3]),
                           np.einsum("mji,rmsj->rmsi", asm.R, sec[..., 3:6])], axis=-1)
    glo = np.concatenate([fg[:, :, None, 0:6], fg[:, :, None, 6:12], secg], axis=2)
    return loc.max(2), loc.min(2), glo.max(2), glo.min(2)


def recover_plates(asm: Assembly, U):
    """Plate resultants at the centroid + corner forces.  Returns dict of arrays."""
    model = asm.model
    nrhs, npl = U.shape[0], model.n_plates
    res = np.zeros((nrhs, npl, 8))        # nx ny nxy mx my mxy qx qy
    nod = np.zeros((nrhs, npl, 3, 6))
    cdg = np.zeros((nrhs, npl, 6)); cdl = np.zeros((nrhs, npl, 6))
    cen = np.zeros((npl, 3))
    for e, (kge, R, xl, mats, T) in enumerate(asm.plates):
        Bm, Bb, Bs, Dm, Db, Ds, A = mats
        ed = (6 * model.p_nodes[e][:, None] + np.arange(6)).reshape(-1)
        cen[e] = asm.xyz[model.p_nodes[e]].mean(0)
        t = model.p_data[e, 0]
        for r in range(nrhs):
            ug = U[r, ed]
            ul = T @ ug
            off = float(model.p_data[e, 14]) if model.p_data.shape[1] > 14 else 0.0
            ue = plate_offset_matrix(off) @ ul if off != 0.0 else ul      # mid-surface displacements
            ul_nodes, ul = ul, ue
            um = ul[[0, 1, 6, 7, 12, 13]]
            ub = np.zeros(9)
            for n in range(3):
                ub[3 * n] = ul[6 * n + 2]; ub[3 * n + 1] = ul[6 * n + 4]; ub[3 * n + 2] = -ul[6 * n + 3]
            flags = int(model.p_flags[e])
            Nm_ = t * Dm @ (Bm @ um) if flags & 1 else np.zeros(3)
            Mb = -Db @ (Bb @ ub) if flags & 2 else np.zeros(3)
            Qs = Ds @ (Bs @ ub) if flags & 2 else np.zeros(2)
            res[r, e] = np.concatenate([Nm_, Mb, Qs])
            nod[r, e] = (kge @ ug).reshape(3, 6)
            cdg[r, e] = ug.reshape(3, 6).mean(0)
            cdl[r, e] = ul_nodes.reshape(3, 6).mean(0)
    return dict(p_res=res, p_nodal=nod, p_cdg=cdg, p_cdl=cdl, p_centroid=cen)


def analyze(model: wire.ModelArrays, timings: dict | None = None, detail: bool = True, exact_mpc: bool | None = None) -> ResultArrays:
    """Full analysis of one document: the oracle twin of fers_core_b200.engine.analyze.
    detail=False skips interior section forces / extrema (as the engine's detail=False)."""
    if len(model.c_member):
        # catenary cables: outside their own element the member rows are pin-ended bars without stiffness
        model = dataclasses.replace(model, m_type=np.where(model.m_type == wire.MT_CABLE, wire.MT_TRUSS, model.m_type).astype(np.int32))
    opts = model.options
    plan = rhs_plan(model)              # list of RhsGroup(imperfection case, rhs columns ...)
    nn, nm, nd = model.n_nodes, model.n_members, model.n_dof
    nrhs = plan.n_rhs
    ns = len(SECTION_FRACS)
    out = ResultArrays.empty(model, plan)
    second_order = opts.order == "NONLINEAR"
    unilateral_members = np.isin(model.m_type, (wire.MT_TENSION, wire.MT_COMPRESSION)) | (len(model.c_member) > 0)
    t0 = time.perf_counter()
    # RigidStrategy.LinearMpc (pydantic_models.py:1047-1049): exact constraints where mpc_applies, else stiff members + a note;
    # exact_mpc=True / False forces either (tests measure the stiff member against the constraint)
    wants_mpc = opts.rigid_strategy == "LinearMpc" and bool(np.any(model.m_type == wire.MT_RIGID))
    use_mpc = mpc_applies(model) if exact_mpc is None else (bool(exact_mpc) and mpc_applies(model))
    if exact_mpc is None and not wants_mpc:
        use_mpc = False
    if wants_mpc and not use_mpc:
        out.warnings.append(("W_LINEAR_MPC_AS_RIGID_MEMBER", "rigid members solved as stiff members (RigidMember strategy)"))
    for grp in plan.groups:
        xyz = imperfect_coordinates(model, grp.imperfection) if grp.imperfection is not None else model.xyz
        asm = Assembly(model, xyz)
        F_lc, peq_lc = load_vectors(asm)
        spans = span_load_tables(asm)
        fac = plan.factors[grp.columns]                         # [ng,Nlc]
        F = fac @ F_lc
        peq = np.einsum("rl,lmk->rmk", fac, peq_lc)
        cols = list(grp.columns)
        needs_own = (second_order or np.any(unilateral_members) or len(unilateral_dofs(model)[0]) > 0
                     or len(model.support_curves) > 0 or len(model.hinge_curves) > 0)
        if not needs_own and use_mpc:
            # ---- linear with RigidStrategy.LinearMpc: the rigid members as exact constraints (master-slave elimination)
            U, links, asm0, K0 = solve_linear_mpc(model, xyz, F)
            _store(out, asm0, model, cols, U, F, peq, K0, None, spans, fac, iters=None, detail=detail, links=links)
        elif not needs_own:
            # ---- linear: one factorisation, all right-hand sides
            K = asm.K_elastic()
            Kff, free, dead = constrain(K, model)
            Uf = solve_sym(Kff, F[:, free].T).T
            U = np.zeros((len(cols), nd)); U[:, free] = Uf
            _store(out, asm, model, cols, U, F, peq, K, None, spans, fac, iters=None, detail=detail)
        else:
            for ci, col in enumerate(cols):
                def reload(asm_it, row=fac[ci]):
                    # equivalent loads of released members depend on the hinge stiffness of the iteration
                    F_it, peq_it = load_vectors(asm_it)
                    return row @ F_it, np.einsum("l,lmk->mk", row, peq_it)
                cab_q = None
                if len(model.c_member):
                    # load per unstretched length of every cable in this column: own weight + the uniform member loads
                    cab_q = model.c_w + sum(fac[ci, li] * lc.cab_q for li, lc in enumerate(model.load_cases))
                U1, K1, kg1, scale1, it, conv, resn, asm_c, F1, peq1, cab = _solve_nonlinear(asm, model, F[ci], peq[ci], second_order,
                                                                                           reload, cab_q)
                _store(out, asm_c, model, [col], U1[None], F1[None], peq1[None], K1,
                       None if kg1 is None else kg1[None], spans, fac[ci:ci + 1], iters=(it, conv, resn), detail=detail, cables=cab)
    if model.hinge_caps and model.n_members:
        out.warnings += host_oracle.hinge_capacity_notes(model, out.f_local, plan.names)
    if timings is not None:
        timings["total_s"] = time.perf_counter() - t0
    return out


def mpc_applies(model) -> bool:
    """Where the rigid members can be exact constraints: a first-order analysis on the perfect geometry without unilateral
    parts or load-dependent springs; rigid members free of member loads, forming trees with at most one supported node each;
    slave nodes (every node of a tree but its root) without supports, plates or warping members; no elastic member with
    both ends in one tree.  (The product states the same conditions in hostlogic.mpc_transform; tests hold the two together.)"""
    o = model.options
    if not np.any(model.m_type == wire.MT_RIGID) or o.order != "LINEAR" or model.imperfections:
        return False
    if len(model.support_curves) or len(model.hinge_curves) or len(unilateral_dofs(model)[0]):
        return False
    if np.isin(model.m_type, (wire.MT_TENSION, wire.MT_COMPRESSION, wire.MT_CABLE)).any() or len(model.c_member):
        return False
    rigid = set(np.nonzero(model.m_type == wire.MT_RIGID)[0].tolist())
    for lc in model.load_cases:
        if rigid & set(np.asarray(lc.dl_member).tolist()) or rigid & set(np.asarray(lc.pl_member).tolist()):
            return False
    try:
        _parent, _via, root = mpc_forest(model)
    except ValueError:
        return False
    slave = root != np.arange(model.n_nodes)
    if (model.bc_code.reshape(model.n_nodes, 6)[slave] != wire.BC_FREE).any():
        return False
    if model.n_plates and slave[model.p_nodes].any():
        return False
    if len(model.w_member) and (slave[model.m_n1[model.w_member]].any() or slave[model.m_n2[model.w_member]].any()):
        return False
    elastic = model.m_type != wire.MT_RIGID
    return not bool((elastic & (root[model.m_n1] == root[model.m_n2])).any())


def mpc_forest(model):
    """The rigid members as a forest over the nodes: (parent node or -1, rigid member to the parent or -1, root) per node.
    The root of a cluster is its supported node when it has one, else its lowest node index."""
    nn = model.n_nodes
    rigid = np.nonzero(model.m_type == wire.MT_RIGID)[0]
    adj = {}
    for m in rigid:
        a, b = int(model.m_n1[m]), int(model.m_n2[m])
        adj.setdefault(a, []).append((b, int(m))); adj.setdefault(b, []).append((a, int(m)))
    parent = np.full(nn, -1); via = np.full(nn, -1); root = np.arange(nn)
    seen = set()
    for start in sorted(adj):
        if start in seen:
            continue
        comp, stack = [], [start]
        seen.add(start)
        while stack:
            v = stack.pop(); comp.append(v)
            for w, _m in adj[v]:
                if w not in seen:
                    seen.add(w); stack.append(w)
        supported = [v for v in sorted(comp) if model.node_support_id[v] >= 0]
        if len(supported) > 1:
            raise ValueError("LinearMpc: a rigid cluster with more than one supported node is over-constrained")
        r0 = supported[0] if supported else min(comp)
        order, stack, done = [], [r0], {r0}
        while stack:
            v = stack.pop(); order.append(v)
            for w, m in adj[v]:
                if w not in done:
                    done.add(w); parent[w] = v; via[w] = m; stack.append(w)
        if len(done) != len(comp):
            raise ValueError("LinearMpc: rigid members form a loop")
        for v in comp:
            root[v] = r0
    return parent, via, root


def solve_linear_mpc(model, xyz, F):
    """First-order solve with the rigid members as exact kinematic constraints  u_s = u_m + theta_m x (x_s - x_m),
    theta_s = theta_m  (master m = root of the rigid cluster): u = T u_red, (T^T K0 T) u_red = T^T f with K0 the structure
    WITHOUT its rigid members.  Returns U [ncols, nd], (reactions carried through the links [ncols, nd], rigid member rows,
    their global end forces [ncols, n_rigid, 12] in the stiffness sign), the assembly without the links and K0."""
    nd, nn = model.n_dof, model.n_nodes
    rigid = np.nonzero(model.m_type == wire.MT_RIGID)[0]
    asm0 = Assembly(model, xyz, m_scale=np.where(model.m_type == wire.MT_RIGID, 0.0, 1.0))
    K0 = asm0.K_elastic()
    parent, via, root = mpc_forest(model)
    slaves = np.nonzero(root != np.arange(nn))[0]
    code = model.bc_code.reshape(nn, 6)
    if np.any(code[slaves] != wire.BC_FREE):
        raise ValueError("LinearMpc: a support on a slave node of a rigid cluster")
    rows, colsT, vals = list(range(nd)), list(range(nd)), [1.0] * nd
    T = sp.lil_matrix((nd, nd))
    T.setdiag(1.0)
    for s_ in slaves:
        m_ = int(root[s_])
        r = xyz[s_] - xyz[m_]
        T[6 * s_:6 * s_ + 6, 6 * s_:6 * s_ + 6] = 0.0
        for c in range(3):
            T[6 * s_ + c, 6 * m_ + c] = 1.0
            T[6 * s_ + 3 + c, 6 * m_ + 3 + c] = 1.0
        # u_s = u_m + theta x r  ->  -[r]_x theta
        T[6 * s_ + 0, 6 * m_ + 4] = r[2];  T[6 * s_ + 0, 6 * m_ + 5] = -r[1]
        T[6 * s_ + 1, 6 * m_ + 3] = -r[2]; T[6 * s_ + 1, 6 * m_ + 5] = r[0]
        T[6 * s_ + 2, 6 * m_ + 3] = r[1];  T[6 * s_ + 2, 6 * m_ + 4] = -r[0]
    T = T.tocsr()
    kspring = np.where(model.bc_code.reshape(-1) == wire.BC_SPRING, model.bc_k.reshape(-1), 0.0)
    Kr = (T.T @ (K0 + sp.diags(kspring)) @ T).tocsr()
    Fr = (T.T @ F.T).T
    fixed = model.bc_code.reshape(-1) == wire.BC_FIXED
    slave_dof = np.zeros(nd, bool)
    slave_dof[(6 * slaves[:, None] + np.arange(6)).reshape(-1)] = True
    diag = Kr.diagonal()
    dead = (~fixed) & ((diag == 0.0) | slave_dof)
    free = np.nonzero(~(fixed | dead))[0]
    Ur = np.zeros((F.shape[0], nd))
    Ur[:, free] = solve_sym(Kr[free][:, free].tocsc(), Fr[:, free].T).T
    U = (T @ Ur.T).T
    # what the links must supply at every node: rho = K0 u - f (springs are supports, excluded as everywhere)
    rho = (K0 @ U.T).T - F                                       # [ncols, nd]
    R_links = (T.T @ rho.T).T                                    # the whole cluster's demand arrives at its root
    # end forces of the links: the demand of the subtree beyond a link, carried to its far node
    depth = np.zeros(nn, int)
    for v in range(nn):
        d, w = 0, v
        while parent[w] >= 0:
            w = parent[w]; d += 1
        depth[v] = d
    sub = rho.reshape(F.shape[0], nn, 6).copy()                  # accumulated demand at each node, about that node
    fg = np.zeros((F.shape[0], len(rigid), 12))
    row_of = {int(m): i for i, m in enumerate(rigid)}
    for v in sorted(np.nonzero(parent >= 0)[0], key=lambda n: -depth[n]):
        p_, m_ = int(parent[v]), int(via[v])
        Fv, Mv = sub[:, v, :3], sub[:, v, 3:]
        # K_rigid u = f - K0 u = -rho: in the stiffness sign (forces of the nodes on the member ends) the far node acts on the
        # link with minus the demand of its subtree, the near node with that demand carried over the arm (link in equilibrium)
        arm = xyz[v] - xyz[p_]
        at_v = np.concatenate([-Fv, -Mv], 1)
        at_p = np.concatenate([Fv, Mv + np.cross(arm[None, :], Fv)], 1)
        i = row_of[m_]
        if int(model.m_n2[m_]) == v:
            fg[:, i, 6:] = at_v; fg[:, i, :6] = at_p
        else:
            fg[:, i, :6] = at_v; fg[:, i, 6:] = at_p
        sub[:, p_, :3] += Fv
        sub[:, p_, 3:] += Mv + np.cross(arm[None, :], Fv)
    return U, (R_links, rigid, fg), asm0, K0


CABLE_TOL = 1.0e-9         # relative change of the displacements between two Newton steps of the cable state
CABLE_MAX_ITER = 60


def _axial_forces(asm, U, peq, kg_local=None):
    rec = recover_members(asm, U[None], peq[None], None if kg_local is None else kg_local[None])
    return rec["f_local"][0][:, 6]


def _cable_linearisation(model, xyz, U, cab_q):
    """Catenary cables at the displacement state U (oracle/cable_oracle.py): sparse tangent Kc, the right-hand-side
    correction Kc U - p (p the internal forces of the cables at U) and the state (a, q) the results are reported from."""
    nd = model.n_dof
    cm = model.c_member
    ni, nj = model.m_n1[cm], model.m_n2[cm]
    u = U.reshape(-1, 6)
    d = (xyz[nj] + u[nj, :3]) - (xyz[ni] + u[ni, :3])
    a, G = cable_oracle.end_tension(d, cab_q, model.c_l0, model.c_ea)
    Kc = sp.lil_matrix((nd, nd))
    p = np.zeros(nd)
    for c in range(len(cm)):
        di = np.arange(6 * ni[c], 6 * ni[c] + 3); dj = np.arange(6 * nj[c], 6 * nj[c] + 3)
        for r_, c_, sg in ((di, di, 1.0), (di, dj, -1.0), (dj, di, -1.0), (dj, dj, 1.0)):
            Kc[np.ix_(r_, c_)] = Kc[np.ix_(r_, c_)] + sg * G[c]
        p[di] += -a[c]
        p[dj] += a[c] - cab_q[c] * model.c_l0[c]
    Kc = Kc.tocsr()
    return Kc, Kc @ U - p, (a, np.array(cab_q, float))


def _solve_nonlinear(asm0: Assembly, model, F, peq, second_order, reload=None, cab_q=None):
    """Active-set (tension/compression-only members, unilateral supports) and load-dependent springs (support and hinge
    stiffness curves) in one outer loop around the P-Delta fixed point.  ``reload(asm) -> (F, peq)`` re-forms the
    column's loads for a new state of the hinge springs.  Returns the assembly, loads of the final state as well."""
    opts = model.options
    nd = model.n_dof
    uni = np.nonzero(np.isin(model.m_type, (wire.MT_TENSION, wire.MT_COMPRESSION)))[0]
    scale = np.ones(model.n_members)
    asm = asm0
    n_cab = 0 if cab_q is None else len(model.c_member)
    if n_cab:
        # cables are catenary elements outside the beam assembly: their member rows carry no stiffness
        scale[model.c_member] = 0.0
        asm = Assembly(model, asm0.xyz, scale)
    F_ext = F
    U = np.zeros(nd)
    Kc, cab = None, None
    total_it, conv, resn = 0, True, 0.0
    suppressed = pdelta_suppressed_axes(model)
    sup_idx = unilateral_dofs(model)[0]
    sup_active = np.ones(len(sup_idx), bool)            # unilateral supports start closed
    n_curves = len(model.support_curves)
    curve_k = curve_stiffness(model, None)              # load-dependent springs start at their zero-force stiffness
    n_hinge = len(model.hinge_curves)
    hinge_k = hinge_curve_stiffness(model, None)        # = the values model.rel carries
    model_it = model
    for _outer in range(max(1, 2 * (len(uni) + len(sup_idx)) + 2 + (max(opts.max_iterations, 20) if n_curves or n_hinge else 0)
                            + (CABLE_MAX_ITER if n_cab else 0))):
        bc = effective_bc(model, sup_active, curve_k) if (len(sup_idx) or n_curves) else None
        K0 = asm.K_elastic()
        U_prev = U
        if n_cab:
            # Newton step in total form: (K + Kc) u = f - p(u_k) + Kc u_k
            Kc, corr, cab = _cable_linearisation(model, asm0.xyz, U_prev, cab_q)
            K0 = (K0 + Kc).tocsr()
            F = F_ext + corr
        Kff, free, _ = constrain(K0, model, bc=bc)
        U = np.zeros(nd); U[free] = solve_sym(Kff, F[free][:, None])[:, 0]
        kg = None
        K = K0
        if second_order and model.n_members:
            conv = False
            for it in range(max(1, opts.max_iterations)):
                N = _axial_forces(asm, U, peq, kg)
                kg = geometric_stiffness(N, asm.L, opts.pdelta_formulation, model.m_type)
                for i, E in asm.arms.items():
                    kg[i] = E.T @ kg[i] @ E
                kg = _suppress_kg(asm, kg, suppressed)
                K = asm.K_elastic(kg)
                if n_cab:
                    K = (K + Kc).tocsr()
                Kff, free, _ = constrain(K, model, bc=bc)
                Un = np.zeros(nd); Un[free] = solve_sym(Kff, F[free][:, None], require_pd=True)[:, 0]
                total_it += 1
                den = np.linalg.norm(Un)
                resn = np.linalg.norm(Un - U) / den if den > 0 else 0.0
                U = Un
                if resn < opts.tolerance:
                    conv = True
                    break
            # geometric stiffness consistent with the converged displacement
            N = _axial_forces(asm, U, peq, kg)
        changed = False
        if n_cab:
            den = np.linalg.norm(U)
            changed = bool(den > 0.0 and np.linalg.norm(U - U_prev) > CABLE_TOL * den)
        if len(sup_idx):
            new_active = unilateral_update(model, sup_active, U, K @ U - F)
            changed = not np.array_equal(new_active, sup_active)
            sup_active = new_active
        if n_curves:
            new_k = curve_stiffness(model, K @ U - F)
            changed = changed or not curves_converged(new_k, curve_k, opts.tolerance)
            curve_k = new_k
        hinges_moved = False
        if n_hinge:
            f_local = recover_members(asm, U[None], peq[None], None if kg is None else kg[None])["f_local"][0]
            new_k = hinge_curve_stiffness(model, f_local)
            hinges_moved = not curves_converged(new_k, hinge_k, opts.tolerance)
            changed = changed or hinges_moved
            hinge_k = new_k
        if len(uni) == 0:
            if changed:
                if hinges_moved:
                    model_it = dataclasses.replace(model, rel=rel_with_hinge_curves(model, hinge_k))
                    asm = Assembly(model_it, asm0.xyz, scale)
                    F_ext, peq = reload(asm)
                    F = F_ext
                continue
            break
        N = _axial_forces(asm, U, peq, kg)
        # elongation decides re-activation of inactive members
        ue = U[asm.edof[uni]]
        ul = np.einsum("mij,mj->mi", asm.T[uni], ue)
        elong = ul[:, 6] - ul[:, 0]
        tens = np.isin(model.m_type[uni], (wire.MT_TENSION, wire.MT_CABLE))
        want_active = np.where(tens, elong >= 0.0, elong <= 0.0)
        new_scale = scale.copy()
        new_scale[uni] = np.where(want_active, 1.0, INACTIVE_SCALE)
        if np.array_equal(new_scale, scale) and not changed:
            break
        scale = new_scale
        if hinges_moved:
            model_it = dataclasses.replace(model, rel=rel_with_hinge_curves(model, hinge_k))
        asm = Assembly(model_it, asm0.xyz, scale)
        if hinges_moved:
            F_ext, peq = reload(asm)
            F = F_ext
    else:
        conv = False        # outer bound reached with the active set / stiffness curves still changing (engine: W_ACTIVE_SET_NOT_CONVERGED)
    if n_cab:
        # report from the cable state AT the final displacements: reactions K u - F then hold the cables' true end forces
        Kc_f, corr, cab = _cable_linearisation(model, asm0.xyz, U, cab_q)
        K = (K - Kc + Kc_f).tocsr()
        F = F_ext + corr
    return U, K, kg, (scale if len(uni) else None), (total_it if second_order else None), conv, resn, asm, F, peq, cab


def _suppress_kg(asm, kg, suppressed):
    """pdelta_suppress_axes: zero the global translational rows/cols of the element K_g
    (pydantic_models.py:1381-1384)."""
    if not suppressed:
        return kg
    kgg = np.einsum("mji,mjk,mkl->mil", asm.T, kg, asm.T)
    for ax in suppressed:
        for d in (ax, ax + 6):
            kgg[:, d, :] = 0.0
            kgg[:, :, d] = 0.0
    return np.einsum("mij,mjk,mlk->mil", asm.T, kgg, asm.T)


def _store(out: ResultArrays, asm, model, cols, U, F, peq, K, kg, spans, fac, iters, detail=True, links=None, cables=None):
    """Fill result arrays for the given rhs columns.  links = (R_full [ncols, nd], member rows, f_global [ncols, n, 12]) of a
    LinearMpc solve: the reactions carried through the rigid links and the end forces of the links themselves."""
    nn = model.n_nodes
    cols = np.asarray(cols)
    out.U[cols] = U.reshape(len(cols), nn, 6)
    # reactions R = K u - F at supported nodes; springs excluded from K
    Rfull = (K @ U.T).T - F if links is None else links[0]
    sup = model.node_support_id >= 0
    Rn = Rfull.reshape(len(cols), nn, 6) * sup[None, :, None]
    out.R[cols] = Rn
    rec = recover_members(asm, U, peq, kg)
    if links is not None:
        rows, fg = links[1], links[2]
        rec["f_global"][:, rows] += fg
        rec["f_local"][:, rows] += np.einsum("mij,cmj->cmi", asm.T[rows], fg)
    if cables is not None:
        # end forces of the catenary cables (one column): the nodes act on the cable with -a at the start, a - q L0 at the end
        a, q = cables
        rows = model.c_member
        fg = np.zeros((len(rows), 12))
        fg[:, 0:3] = -a
        fg[:, 6:9] = a - q * model.c_l0[:, None]
        rec["f_global"][0, rows] = fg
        rec["f_local"][0, rows] = np.einsum("mij,mj->mi", asm.T[rows], fg)
    out.f_local[cols] = rec["f_local"]; out.f_global[cols] = rec["f_global"]; out.u_local[cols] = rec["u_local"]
    out.bw[cols] = rec["bw"]
    if model.n_members and detail:
        span_by_rhs = [[(fac[r, li], spans[li]) for li in range(fac.shape[1])] for r in range(len(cols))]
        sec = section_forces(asm, rec["f_local"], span_by_rhs)
        if cables is not None:
            # along a cable: the start force carried along the unstretched length, -t(s) = q s - a, no moments
            a, q = cables
            rows = model.c_member
            for si, x in enumerate(SECTION_FRACS):
                tg = q * (x * model.c_l0)[:, None] - a
                sec[0, rows, si, 0:3] = np.einsum("mij,mj->mi", asm.R[rows], tg)
                sec[0, rows, si, 3:6] = 0.0
        out.sec[cols] = sec
        lmax, lmin, gmax, gmin = min_max(asm, rec["f_local"], rec["f_global"], sec)
        out.lmax[cols], out.lmin[cols], out.gmax[cols], out.gmin[cols] = lmax, lmin, gmax, gmin
    if model.n_plates:
        pr = recover_plates(asm, U)
        out.p_res[cols] = pr["p_res"]; out.p_nodal[cols] = pr["p_nodal"]
        out.p_cdg[cols] = pr["p_cdg"]; out.p_cdl[cols] = pr["p_cdl"]; out.p_centroid = pr["p_centroid"]
    if iters is not None:
        it, conv, resn = iters
        for c in cols:
            out.iterations[c] = -1 if it is None else it
            out.converged[c] = bool(conv)
            out.residual[c] = resn
