"""CPU ORACLE -- elastic catenary cable element.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

`MemberType.Cable` (pydantic_models.py:592-598) with `pretension` / `unstretched_length` (1525-1528, 1542-1545) and intrinsic
self weight (1337-1340).  The reference's engine is closed, no cable example or test ships with it: PARITY UNPINNED against the
reference; the element is pinned on the classical catenary instead (tests/test_cable.py: inextensible sag / tension closed
form, straight bar limit, equilibrium, work-conjugate tangent).

The product (fers_core_b200/cable.py) solves the two closed-form span equations in the cable's plane.  This restatement does
NOT use them: it integrates the defining relation

    dx/ds = t/|t| (1 + |t|/EA),   t(s) = a - q s,   s in [0, L0]

numerically (composite Gauss-Legendre) and finds the start tension vector a [3] by a Newton iteration in three unknowns with
the Jacobian  dd/da = int (I - tt^T/|t|^2)/|t| ds + (L0/EA) I  from the same quadrature.
"""
from __future__ import annotations

import numpy as np

_GL_X, _GL_W = np.polynomial.legendre.leggauss(12)
PANELS = 400
SLACK_STIFFNESS = 1.0e-9


def _stations(l0: float):
    edges = np.linspace(0.0, l0, PANELS + 1)
    half = 0.5 * (edges[1:] - edges[:-1])
    mid = 0.5 * (edges[1:] + edges[:-1])
    s = (mid[:, None] + half[:, None] * _GL_X[None, :]).reshape(-1)
    wt = (half[:, None] * _GL_W[None, :]).reshape(-1)
    return s, wt


def span_of(a, q, l0: float, ea: float):
    """End-to-end vector d(a) and its Jacobian for one cable."""
    s, wt = _stations(l0)
    t = a[None, :] - q[None, :] * s[:, None]
    tn = np.linalg.norm(t, axis=1)
    that = t / tn[:, None]
    d = (that * (wt * (1.0 + tn / ea))[:, None]).sum(0)
    J = np.eye(3) * ((wt / tn).sum() + l0 / ea) - np.einsum("s,si,sj->ij", wt / tn, that, that)
    return d, J


def end_tension_one(d, q, l0: float, ea: float):
    """(a, G): tension vector at the start and the tangent dp_j/dx_j of one cable."""
    d = np.asarray(d, float); q = np.asarray(q, float)
    L = float(np.linalg.norm(d))
    w = float(np.linalg.norm(q))
    if w * l0 <= 1e-14 * ea:
        n = d / L
        T = ea * (L - l0) / l0
        kax = ea / l0 if T > 0.0 else SLACK_STIFFNESS * ea / l0
        T = max(T, 0.0)
        nn = np.outer(n, n)
        return T * n, kax * nn + (T / L) * (np.eye(3) - nn)
    # start: straight bar tension (or a fraction of the weight when slack) along the chord, half of the load at each end
    T0 = max(ea * (L - l0) / l0, 0.5 * w * l0)
    a = T0 * d / L + 0.5 * q * l0
    for _ in range(400):
        dc, J = span_of(a, q, l0, ea)
        r = d - dc
        err = np.linalg.norm(r) / l0
        if err < 1e-13:
            break
        da = np.linalg.solve(J, r)
        step = 1.0
        for _bt in range(60):
            an = a + step * da
            rn = np.linalg.norm(d - span_of(an, q, l0, ea)[0]) / l0
            if rn <= err:
                break
            step *= 0.5
        a = an
    else:
        raise ArithmeticError("cable oracle: Newton iteration did not converge")
    _, J = span_of(a, q, l0, ea)
    return a, np.linalg.inv(J)


def end_tension(d, q, l0, ea):
    d = np.asarray(d, float).reshape(-1, 3)
    a = np.zeros_like(d); G = np.zeros((d.shape[0], 3, 3))
    for i in range(d.shape[0]):
        a[i], G[i] = end_tension_one(d[i], np.asarray(q, float).reshape(-1, 3)[i], float(l0[i]), float(ea[i]))
    return a, G
