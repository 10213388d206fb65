"""Elastic catenary of `MemberType.Cable` members (pydantic_models.py:592-598; `pretension` 1525-1528, `unstretched_length`
1542-1545; self weight "carried intrinsically in the catenary formulation", 1337-1340).

A cable of unstretched length L0 and axial stiffness EA under a uniform load q per unit of UNSTRETCHED length (q [3], global
axes: own weight along the gravity direction plus the uniform member loads of the combination) hangs between its end nodes.
With a the tension vector at the start (pointing into the cable), the tension vector at the unstretched arc coordinate s is
t(s) = a - q s and the end-to-end vector is

    d(a) = int_0^L0  t/|t| ds  +  (a L0 - q L0^2 / 2) / EA .

In the plane of d and q (e_q = q/|q|, e_h the unit vector of the part of d normal to q) a = H e_h + p0 e_q and, p1 = p0 - w L0,

    l = H [ L0/EA + (asinh(p0/H) - asinh(p1/H)) / w ]               (normal to q)
    h = (p0 L0 - w L0^2/2)/EA + (sqrt(H^2+p0^2) - sqrt(H^2+p1^2)) / w   (along q)

which `end_tension` solves for (H, p0) by a damped Newton iteration, all cables of the model at once.  The internal forces
of the element (forces of the end nodes on the cable, the sign of K u = f) are  p_i = -a,  p_j = a - q L0,  and its tangent
is [[G, -G], [-G, G]] with G = (dd/da)^-1, formed in closed form (`_flex2`).  Without a load (|q| = 0) the cable is a
straight bar that only pulls: T = EA (|d| - L0)/L0 when positive, else slack.  The unstretched length itself comes from the
wire (`unstretched_length`, else L_chord / (1 + pretension/EA), else the chord: wire._parse_members).

Everything here is host-side bookkeeping of a handful of cables per model (numpy, vectorised over the cables); the global
tangent blocks and the right-hand-side corrections it produces are assembled and solved on the device (engine.py).
"""
from __future__ import annotations

import numpy as np

class CableError(RuntimeError):
    """The catenary equations of a cable have no solution the iteration can reach (the boundary raises RuntimeError for
    every analysis failure, as the reference's engine does: fers_core/fers/fers.py:301-302)."""


SLACK_STIFFNESS = 1.0e-9          # a slack weightless cable keeps this fraction of EA/L0 along its chord (K stays regular)
NEWTON_MAX = 200
NEWTON_TOL = 1.0e-13


def _plane(d: np.ndarray, q: np.ndarray):
    """Unit vectors (e_h, e_q), spans (l, h) and w = |q| per cable; e_h is any normal of q where d is parallel to it."""
    w = np.linalg.norm(q, axis=1)
    e_q = q / np.where(w > 0.0, w, 1.0)[:, None]
    h = (d * e_q).sum(1)
    dn = d - h[:, None] * e_q
    l = np.linalg.norm(dn, axis=1)
    tiny = l <= 1e-12 * np.maximum(np.linalg.norm(d, axis=1), 1e-300)
    if tiny.any():
        # a normal of e_q: cross with the axis it is least aligned with
        ax = np.eye(3)[np.argmin(np.abs(e_q[tiny]), axis=1)]
        n = np.cross(e_q[tiny], ax)
        dn[tiny] = n / np.linalg.norm(n, axis=1)[:, None]
        l_safe = np.where(tiny, 1.0, l)
    else:
        l_safe = l
    e_h = dn / l_safe[:, None]
    return e_h, e_q, l, h, w


def _spans(H, p0, w, l0, ea):
    p1 = p0 - w * l0
    t0, t1 = np.hypot(H, p0), np.hypot(H, p1)
    # differences written without cancellation (a taut cable has t0 ~ t1 >> w l0):
    #   t0 - t1 = (p0^2 - p1^2) / (t0 + t1);  asinh(x0) - asinh(x1) = asinh(x0 sqrt(1 + x1^2) - x1 sqrt(1 + x0^2)), whose
    #   argument is H^2 (p0^2 - p1^2) / (p0 t1 + p1 t0) / H^2 when p0 and p1 have the same sign (no cancellation in the sum)
    dt = w * l0 * (p0 + p1) / (t0 + t1)
    same = p0 * p1 > 0.0
    den = np.where(same, p0 * t1 + p1 * t0, 1.0)
    A = np.where(same, np.arcsinh(w * l0 * (p0 + p1) / den), np.arcsinh(p0 / H) - np.arcsinh(p1 / H)) / w
    l = H * (l0 / ea + A)
    h = (p0 * l0 - 0.5 * w * l0 * l0) / ea + dt / w
    return l, h, p1, t0, t1, A


def _flex2(H, p0, w, l0, ea):
    """In-plane flexibility d(l, h)/d(H, p0) (symmetric) and the out-of-plane term A + L0/EA."""
    l, h, p1, t0, t1, A = _spans(H, p0, w, l0, ea)
    c = l0 / ea
    fhh = c + A - (p0 / t0 - p1 / t1) / w          # int p^2/|t|^3 ds = [asinh(p/H) - p/|t|] / w
    fhq = (H / t0 - H / t1) / w                    # int -pH/|t|^3 ds
    fqq = c + (p0 / t0 - p1 / t1) / w              # int H^2/|t|^3 ds
    return l, h, fhh, fhq, fqq, c + A


def end_tension(d: np.ndarray, q: np.ndarray, l0: np.ndarray, ea: np.ndarray):
    """Tension vector a [nc,3] at the start of every cable and the 3x3 tangent G [nc,3,3] (dp_j/dx_j) for end-to-end vectors
    d [nc,3], loads q [nc,3] per unstretched length, unstretched lengths l0 and axial stiffnesses ea."""
    d = np.asarray(d, float).reshape(-1, 3)
    q = np.asarray(q, float).reshape(-1, 3)
    nc = d.shape[0]
    a = np.zeros((nc, 3))
    G = np.zeros((nc, 3, 3))
    L = np.linalg.norm(d, axis=1)
    w = np.linalg.norm(q, axis=1)
    # ---- weightless cables: a straight bar that only pulls
    bar = w * l0 <= 1e-14 * ea
    if bar.any():
        n = d[bar] / L[bar][:, None]
        T = ea[bar] * (L[bar] - l0[bar]) / l0[bar]
        taut = T > 0.0
        kax = np.where(taut, 1.0, SLACK_STIFFNESS) * ea[bar] / l0[bar]
        T = np.where(taut, T, 0.0)
        nn = n[:, :, None] * n[:, None, :]
        a[bar] = T[:, None] * n
        G[bar] = kax[:, None, None] * nn + (T / L[bar])[:, None, None] * (np.eye(3)[None] - nn)
    cat = ~bar
    if cat.any():
        e_h, e_q, l, h, wc = _plane(d[cat], q[cat])
        l0c, eac = l0[cat], ea[cat]
        # a cable cannot be shorter than a vertical fold allows: keep the horizontal span away from zero
        l = np.maximum(l, 1e-9 * l0c)
        H, p0 = _initial_guess(l, h, wc, l0c, eac)
        for _ in range(NEWTON_MAX):
            lc, hc, fhh, fhq, fqq, _ = _flex2(H, p0, wc, l0c, eac)
            r1, r2 = l - lc, h - hc
            err = np.hypot(r1, r2) / l0c
            if (err < NEWTON_TOL).all():
                break
            det = fhh * fqq - fhq * fhq
            dH = (fqq * r1 - fhq * r2) / det
            dp = (fhh * r2 - fhq * r1) / det
            if (np.maximum(np.abs(dH), np.abs(dp)) <= 1e-14 * np.hypot(H, p0)).all():
                break                                  # the step is below the resolution of the unknowns
            # damping: H stays positive and the residual must not grow
            step = np.ones_like(H)
            neg = H + dH <= 0.0
            step[neg] = np.minimum(1.0, 0.9 * H[neg] / -dH[neg])
            for _bt in range(40):
                Hn, pn = H + step * dH, p0 + step * dp
                ln, hn, *_ = _spans(Hn, pn, wc, l0c, eac)
                worse = (np.hypot(l - ln, h - hn) / l0c > err) & (err >= NEWTON_TOL)
                if not worse.any():
                    break
                step[worse] *= 0.5
            H, p0 = Hn, pn
        else:
            raise CableError("cable: the catenary equations did not converge")
        _, _, fhh, fhq, fqq, fnn = _flex2(H, p0, wc, l0c, eac)
        det = fhh * fqq - fhq * fhq
        ghh, ghq, gqq = fqq / det, -fhq / det, fhh / det
        e_n = np.cross(e_h, e_q)
        a[cat] = H[:, None] * e_h + p0[:, None] * e_q
        outer = lambda u, v: u[:, :, None] * v[:, None, :]
        G[cat] = (ghh[:, None, None] * outer(e_h, e_h) + gqq[:, None, None] * outer(e_q, e_q)
                  + ghq[:, None, None] * (outer(e_h, e_q) + outer(e_q, e_h)) + (1.0 / fnn)[:, None, None] * outer(e_n, e_n))
    return a, G


def _initial_guess(l, h, w, l0, ea):
    """Start of the Newton iteration: the parabola-like estimate for a slack cable, the straight bar for a taut one; each
    end carries half of the load."""
    chord = np.hypot(l, h)
    slack = l0 > chord
    with np.errstate(invalid="ignore", divide="ignore"):
        lam = np.where(slack, np.sqrt(np.maximum(3.0 * ((l0 * l0 - h * h) / (l * l) - 1.0), 1e-12)), 0.2)
    H_slack = w * l / (2.0 * np.maximum(lam, 1e-6))
    T_bar = ea * np.maximum(chord - l0, 0.0) / l0
    H = np.where(slack, H_slack, np.maximum(T_bar * l / chord, H_slack))
    H = np.maximum(H, 1e-12 * w * l0)
    p0 = H * h / l + 0.5 * w * l0
    return H, p0


def element_state(xyz_i, xyz_j, q, l0, ea):
    """Internal forces p [nc,6] (start xyz, end xyz; forces of the nodes on the cable), tangent G [nc,3,3] and the start
    tension vector a for the CURRENT end positions."""
    a, G = end_tension(np.asarray(xyz_j) - np.asarray(xyz_i), q, l0, ea)
    p = np.concatenate([-a, a - np.asarray(q) * np.asarray(l0)[:, None]], axis=1)
    return p, G, a


def tension_along(a, q, l0, fracs):
    """Tension vectors t(s) = a - q s at s = frac * L0: [nc, len(fracs), 3]."""
    s = np.asarray(fracs, float)[None, :, None] * np.asarray(l0, float)[:, None, None]
    return np.asarray(a)[:, None, :] - np.asarray(q)[:, None, :] * s
