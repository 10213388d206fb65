"""MemberType.Cable as an elastic catenary (pydantic_models.py:592-598; `pretension` 1525-1528, `unstretched_length`
1542-1545; own weight "carried intrinsically in the catenary formulation", 1337-1340).

The reference ships no cable example, test or number (its engine is closed): the element is pinned on the classical
catenary instead -- the inextensible closed form  L0 = (2H/w) sinh(w l / 2H),  V = w L0 / 2,  the straight-bar limit, a
pretensioned string under a point load (scalar equilibrium solved by bisection), overall equilibrium and the work-conjugate
tangent -- with the product's closed-form element (fers_core_b200/cable.py) held against the oracle's quadrature
restatement (oracle/cable_oracle.py), and the GPU analysis against the oracle's."""
import copy

import numpy as np
import pytest
from scipy.optimize import brentq

from fers_core_b200 import cable, wire
from oracle import cable_oracle, fers_oracle
from tests import helpers

E_MOD = 210e9


@pytest.fixture(scope="module")
def engine_mod():
    from fers_core_b200 import engine
    return engine


def cable_doc(nodes, members, loads=(), dls=(), order="LINEAR", self_weight=False, combos=()):
    """nodes [(id, x, y, z, support id or None)]: support 1 clamps, 2 pins; members: dicts over the golden cantilever's
    member record; loads [(case, node, magnitude, direction)]; dls [(case, member, magnitude, direction, start, end)]."""
    doc = copy.deepcopy(helpers.golden("cantilever_end_load")["doc"])
    m = doc["model"]
    m["nodes"] = [{"id": i, "X": x, "Y": y, "Z": z, "nodal_support": s} for i, x, y, z, s in nodes]
    fixed = {"condition_type": "Fixed", "stiffness": None}
    free = {"condition_type": "Free", "stiffness": None}
    m["nodal_supports"] = [dict(id=1, classification=None, warping_condition=None, X=fixed, Y=fixed, Z=fixed, RX=fixed, RY=fixed, RZ=fixed),
                           dict(id=2, classification=None, warping_condition=None, X=fixed, Y=fixed, Z=fixed, RX=free, RY=free, RZ=free)]
    base = m["members"][0]
    m["members"] = [dict(base, **mm) for mm in members]
    m["member_sets"] = [{"id": 1, "classification": None, "member_ids": [mm["id"] for mm in members], "buckling_restraints": []}]
    a = doc["analysis"]
    a["options"]["order"] = order
    a["options"]["enable_self_weight"] = self_weight
    lc = a["load_cases"][0]
    cases = sorted({c for c, *_ in loads} | {c for c, *_ in dls}) or [1]
    a["load_cases"] = [
        dict(lc, id=c, name=f"LC{c}",
             nodal_loads=[{"id": 100 * c + k, "node": n, "magnitude": mag, "direction": list(dr), "load_type": "force"}
                          for k, (cc, n, mag, dr) in enumerate(loads) if cc == c],
             distributed_loads=[{"id": 200 * c + k, "member": mm, "magnitude": mag, "end_magnitude": mag, "direction": list(dr),
                                 "start_frac": s0, "end_frac": s1, "load_type": "force"}
                                for k, (cc, mm, mag, dr, s0, s1) in enumerate(dls) if cc == c])
        for c in cases]
    a["load_combinations"] = [{"id": 10 + k, "name": f"C{k}", "load_case_factors": {str(c): f for c, f in fac.items()},
                               "limit_state": "ULS"} for k, fac in enumerate(combos)]
    return doc


def _cab(i, a, b, **kw):
    return dict(id=i, start_node_id=a, end_node_id=b, member_type="Cable", **kw)


def _ea(doc):
    return float(wire.parse_document(doc).prop[0][0])


# ---- the element ---------------------------------------------------------------------------------------------------------
def test_element_equals_the_quadrature_restatement():
    rng = np.random.default_rng(7)
    n = 40
    d = rng.normal(size=(n, 3)) * 10.0
    q = rng.normal(size=(n, 3)) * 5.0
    L = np.linalg.norm(d, axis=1)
    l0 = L * rng.uniform(0.9, 1.3, n)                      # taut and slack
    ea = 10.0 ** rng.uniform(4, 8, n)
    a, G = cable.end_tension(d, q, l0, ea)
    a2, G2 = cable_oracle.end_tension(d, q, l0, ea)
    assert (np.abs(a - a2).max(1) / np.abs(a2).max(1)).max() < 1e-10
    assert (np.abs(G - G2).max((1, 2)) / np.abs(G2).max((1, 2))).max() < 1e-9
    assert np.abs(G - G.transpose(0, 2, 1)).max() <= 1e-12 * np.abs(G).max()
    assert (np.linalg.eigvalsh(G) > 0.0).all()             # a hanging cable is stable in every direction


def test_tangent_is_the_derivative_of_the_end_force():
    d = np.array([[30.0, -4.0, 7.0]]); q = np.array([[0.5, -12.0, 1.0]]); l0 = np.array([33.0]); ea = np.array([3.0e6])
    _, G = cable.end_tension(d, q, l0, ea)
    h = 1e-6
    fd = np.zeros((3, 3))
    for c in range(3):
        e = np.zeros((1, 3)); e[0, c] = h
        fd[:, c] = (cable.end_tension(d + e, q, l0, ea)[0] - cable.end_tension(d - e, q, l0, ea)[0])[0] / (2 * h)
    assert np.abs(fd - G[0]).max() < 1e-7 * np.abs(G).max()


def test_inextensible_catenary_closed_form():
    w, span, H = 10.0, 100.0, 2000.0
    l0 = 2 * H / w * np.sinh(w * span / (2 * H))
    a, _ = cable.end_tension(np.array([[span, 0.0, 0.0]]), np.array([[0.0, -w, 0.0]]), np.array([l0]), np.array([1e15]))
    assert a[0, 0] == pytest.approx(H, rel=1e-8)
    assert a[0, 1] == pytest.approx(-w * l0 / 2, rel=1e-12)     # the start tension points down into the sagging cable
    # sag from the tension at mid-length: T(L0/2) = H, and from the tension at the support T0 = H cosh(w l / 2H)
    assert np.linalg.norm(a[0]) == pytest.approx(H * np.cosh(w * span / (2 * H)), rel=1e-8)


@pytest.mark.parametrize("rise", [-30.0, 12.0, 45.0])
def test_inclined_inextensible_catenary_closed_form(rise):
    """Supports at different levels: an inextensible catenary of horizontal tension H over the span l with the end `rise`
    above the start has the length  L0^2 = rise^2 + (2H/w sinh(w l / 2H))^2  (Irvine, Cable Structures, 1.2); the end forces
    carry the weight (force balance) and balance its moment about the start (centre of gravity from the integrated shape)."""
    w, span, H = 8.0, 120.0, 3500.0
    l0 = np.sqrt(rise ** 2 + (2 * H / w * np.sinh(w * span / (2 * H))) ** 2)
    d = np.array([[span, rise, 0.0]])
    q = np.array([[0.0, -w, 0.0]])
    a, _ = cable.end_tension(d, q, np.array([l0]), np.array([1e16]))
    assert a[0, 0] == pytest.approx(H, rel=1e-7)
    a2, _ = cable_oracle.end_tension(d, q, np.array([l0]), np.array([1e16]))
    assert np.allclose(a, a2, rtol=1e-8)
    # forces of the cable on its end nodes: +a at the start, -(a - q L0) at the end; their sum is the weight
    f_start, f_end = a[0], -(a[0] - q[0] * l0)
    assert np.allclose(f_start + f_end, q[0] * l0, atol=1e-9 * w * l0)
    # moment balance about the start: the end force at (span, rise) against the weight acting at the centre of gravity, whose
    # abscissa follows from the shape x(s) -- integrate the tangent direction t/|t| numerically
    s = np.linspace(0.0, l0, 20001)
    t = a[0][None, :] - q * s[:, None]
    tx = t[:, 0] / np.linalg.norm(t, axis=1)
    x = np.concatenate([[0.0], np.cumsum(0.5 * (tx[1:] + tx[:-1]) * np.diff(s))])
    assert x[-1] == pytest.approx(span, rel=1e-6)
    x_cg = float(np.sum(0.5 * (x[1:] + x[:-1]) * np.diff(s))) / l0
    # sum of moments about the start (z component): the end node pulls the cable with -f_end at (span, rise); weight -w L0 at x_cg
    on_cable_end = -f_end
    assert span * on_cable_end[1] - rise * on_cable_end[0] - w * l0 * x_cg == pytest.approx(0.0, abs=2e-5 * w * l0 * span)


def test_element_is_robust_over_extreme_cables():
    """400 random cables over six decades of length, load and stiffness, from half to three times the chord in unstretched
    length, vertical ones included: the iteration always ends, the tangent is finite and positive semi-definite, and the
    end-to-end vector of the answer (trapezoidal rule over the defining relation) is the one asked for."""
    rng = np.random.default_rng(123)
    for trial in range(400):
        L = 10 ** rng.uniform(-1, 3)
        u = rng.normal(size=3); u /= np.linalg.norm(u)
        g = rng.normal(size=3); g /= np.linalg.norm(g)
        vertical = rng.random() < 0.1
        if vertical:
            g = np.array([0.0, -1.0, 0.0]); u = np.array([0.0, -1.0 if rng.random() < 0.5 else 1.0, 0.0])
        d = (u * L)[None, :]
        w = 10 ** rng.uniform(-3, 3)
        ratio = [rng.uniform(0.5, 0.999), rng.uniform(0.999, 1.001), rng.uniform(1.001, 3.0)][trial % 3]
        l0 = np.array([L * ratio]); ea = np.array([10 ** rng.uniform(3, 10)])
        q = (g * w)[None, :]
        a, G = cable.end_tension(d, q, l0, ea)
        assert np.isfinite(a).all() and np.isfinite(G).all(), trial
        ev = np.linalg.eigvalsh(G[0])
        assert ev.min() >= -1e-9 * np.abs(ev).max(), (trial, ev)
        s = np.linspace(0.0, l0[0], 4001)
        t = a[0][None, :] - q * s[:, None]
        tn = np.linalg.norm(t, axis=1)
        if tn.min() < 1e-3 * tn.max():
            continue                                       # a vertical cable with a slack part folds: the integrand has a kink
        f = t / tn[:, None] * (1.0 + tn / ea[0])[:, None]
        span = np.sum(0.5 * (f[1:] + f[:-1]) * np.diff(s)[:, None], axis=0)
        assert np.linalg.norm(span - d[0]) < 2e-5 * l0[0], (trial, span, d)


def test_weightless_cable_is_a_bar_that_only_pulls():
    d = np.array([[3.0, 4.0, 0.0], [3.0, 4.0, 0.0]]); l0 = np.array([4.9, 5.1]); ea = np.array([1e6, 1e6])
    a, G = cable.end_tension(d, np.zeros((2, 3)), l0, ea)
    assert np.allclose(a[0], 1e6 * 0.1 / 4.9 * np.array([0.6, 0.8, 0.0]))
    assert np.allclose(a[1], 0.0) and np.abs(G[1]).max() <= 1e-8 * np.abs(G[0]).max()
    n = np.array([0.6, 0.8, 0.0])
    assert n @ G[0] @ n == pytest.approx(1e6 / 4.9)
    t = np.array([-0.8, 0.6, 0.0])
    assert t @ G[0] @ t == pytest.approx(np.linalg.norm(a[0]) / 5.0)        # string stiffness T / L across the chord


def test_vertical_cable_hangs_on_its_upper_end():
    w, l0, ea = 10.0, 9.99, 1e6
    a, _ = cable.end_tension(np.array([[0.0, -10.0, 0.0]]), np.array([[0.0, -w, 0.0]]), np.array([l0]), np.array([ea]))
    # stretched by 0.01 in total: mean tension EA * 0.01 / L0, the upper end carries half of the weight more
    assert a[0, 1] == pytest.approx(-(ea * 0.01 / l0 + w * l0 / 2), rel=1e-9)
    assert abs(a[0, 0]) + abs(a[0, 2]) < 1e-3


# ---- the wire ------------------------------------------------------------------------------------------------------------
def test_wire_reads_the_cable_fields():
    nodes = [(1, 0, 0, 0, 2), (2, 10, 0, 0, 2), (3, 20, 0, 0, 2)]
    doc = cable_doc(nodes, [_cab(1, 1, 2, pretension=5000.0), _cab(2, 2, 3, unstretched_length=10.5, pretension=1.0, weight=3.0)],
                    dls=[(1, 1, 7.0, (0, -1, 0), 0.0, 1.0), (1, 2, 7.0, (0, -1, 0), 0.0, 0.5)])
    m = wire.parse_document(doc)
    ea = _ea(doc)
    assert m.c_member.tolist() == [0, 1]
    assert m.c_l0[0] == pytest.approx(10.0 / (1.0 + 5000.0 / ea)) and m.c_l0[1] == 10.5       # unstretched_length wins
    assert np.allclose(m.c_w, [[0, 0, 0], [0, -3.0, 0]])                                      # `weight` override, along gravity
    lc = m.load_cases[0]
    assert np.allclose(lc.cab_q, [[0, -7.0, 0], [0, 0, 0]])               # the full-length uniform load belongs to the catenary
    assert lc.dl_member.tolist() == [1]                                    # the partial one is carried to the end nodes
    assert any(code == "W_CABLE_LOAD_LUMPED" for code, _ in m.warnings)
    assert not any(code == "W_CABLE_AS_TENSION" for code, _ in m.warnings)
    doc["analysis"]["options"]["enable_self_weight"] = True
    m = wire.parse_document(doc)
    assert m.c_w[0, 1] == pytest.approx(-m.prop_rho_a[0] * 9.81) and m.c_w[1, 1] == -3.0
    sw = [lc for lc in m.load_cases if lc.name == "Self weight"]
    assert sw and len(sw[0].dl_member) == 0                                # cables are excluded from the generated self weight


# ---- structures, oracle --------------------------------------------------------------------------------------------------
def _hanging(span=100.0, w=10.0, H=2000.0):
    l0 = 2 * H / w * np.sinh(w * span / (2 * H))
    return cable_doc([(1, 0, 0, 0, 2), (2, span, 0, 0, 2)], [_cab(1, 1, 2, weight=w, unstretched_length=l0)]), l0


def test_hanging_cable_reactions():
    doc, l0 = _hanging()
    res = fers_oracle.analyze(wire.parse_document(doc))
    R = res.R[0].reshape(-1, 6)
    # EA = 4e8: the elastic stretch lowers H by 2.4e-4 of its inextensible value
    assert R[0, 0] == pytest.approx(-2000.0, rel=5e-4) and R[1, 0] == pytest.approx(2000.0, rel=5e-4)
    assert R[0, 1] == pytest.approx(10.0 * l0 / 2, rel=1e-10) and R[1, 1] == pytest.approx(10.0 * l0 / 2, rel=1e-10)
    # tension along the cable: at mid-length it is horizontal and equal to H
    mid = res.sec[0, 0, 1]
    assert abs(mid[0]) == pytest.approx(abs(R[0, 0]), rel=1e-10) and abs(mid[1]) < 1e-6 and np.abs(mid[3:]).max() == 0.0


def _string_doc(b=5.0, n0=20e3, p=3e3, order="LINEAR"):
    nodes = [(1, 0, 0, 0, 2), (2, b, 0, 0, None), (3, 2 * b, 0, 0, 2)]
    return cable_doc(nodes, [_cab(1, 1, 2, pretension=n0), _cab(2, 2, 3, pretension=n0)], loads=[(1, 2, p, (0, -1, 0))], order=order)


@pytest.mark.parametrize("order", ["LINEAR", "NONLINEAR"])
def test_pretensioned_string_under_a_point_load(order):
    b, n0, p = 5.0, 20e3, 3e3
    doc = _string_doc(b, n0, p, order)
    ea = _ea(doc)
    l0 = b / (1.0 + n0 / ea)

    def unbalance(delta):
        L = np.hypot(b, delta)
        return 2.0 * ea * (L - l0) / l0 * delta / L - p
    delta = brentq(unbalance, 0.0, b, xtol=1e-15)
    res = fers_oracle.analyze(wire.parse_document(doc))
    assert res.U[0, 1, 1] == pytest.approx(-delta, rel=1e-8)
    T = ea * (np.hypot(b, delta) - l0) / l0
    assert np.linalg.norm(res.f_global[0, 0, 0:3]) == pytest.approx(T, rel=1e-8)
    assert res.R[0, 0, 1] == pytest.approx(p / 2, rel=1e-8) and res.R[0, 2, 1] == pytest.approx(p / 2, rel=1e-8)
    assert bool(res.converged[0])


def _slack_doc(l0=5.5, p=100.0):
    """Two slack cable segments (10 % longer than their chords) between pinned ends, a point load on the middle node: the
    node drops by 2.25 m on a 10 m span -- far outside small displacements."""
    nodes = [(1, 0, 0, 0, 2), (2, 5, 0, 0, None), (3, 10, 0, 0, 2)]
    return cable_doc(nodes, [_cab(1, 1, 2, unstretched_length=l0, weight=10.0), _cab(2, 2, 3, unstretched_length=l0, weight=10.0)],
                     loads=[(1, 2, p, (0, -1, 0))])


def test_slack_cable_with_a_point_load_large_displacements():
    l0, p, w = 5.5, 100.0, 10.0
    m = wire.parse_document(_slack_doc(l0, p))
    res = fers_oracle.analyze(m)
    assert res.converged.all()
    drop = -res.U[0, 1, 1]
    assert 2.0 < drop < np.sqrt(l0 ** 2 - 5.0 ** 2) and abs(res.U[0, 1, 0]) < 1e-9          # below the straight-segment bound
    # supports: half of the load and of the weight each, equal and opposite horizontal pulls
    assert res.R[0, 0, 1] == pytest.approx((p + 2 * w * l0) / 2, rel=1e-9) and res.R[0, 2, 1] == pytest.approx((p + 2 * w * l0) / 2, rel=1e-9)
    assert res.R[0, 0, 0] == pytest.approx(-res.R[0, 2, 0], rel=1e-9) and res.R[0, 0, 0] < 0.0
    # the left segment in its final position is the catenary the element formulas give: support reaction = -(its pull on node 1)
    a, _ = cable.end_tension(np.array([[5.0, -drop, 0.0]]), np.array([[0.0, -w, 0.0]]), np.array([l0]), m.c_ea[:1])
    assert np.allclose(res.R[0, 0, :3], -a[0], rtol=1e-8, atol=1e-8 * np.abs(a).max())


def test_chain_of_cable_segments_is_the_single_catenary():
    """A heavy cable cut into four elements at free nodes is still ONE catenary: after the nodes have dropped (2.8 m on a 20 m
    span -- form finding from the straight chord) the horizontal reaction equals the horizontal tension of the single elastic
    catenary of the same total unstretched length, the nodes lie on its curve, and the net of four cables on a loaded node,
    a stay that goes slack under uplift and a 30 % slack chain all converge."""
    w, slack = 50.0, 1.05
    nodes = [(1, 0, 0, 0, 2), (2, 5, 0, 0, None), (3, 10, 0, 0, None), (4, 15, 0, 0, None), (5, 20, 0, 0, 2)]
    m = wire.parse_document(cable_doc(nodes, [_cab(i, i, i + 1, unstretched_length=5 * slack, weight=w) for i in (1, 2, 3, 4)]))
    res = fers_oracle.analyze(m)
    assert res.converged.all()
    q = np.array([[0.0, -w, 0.0]])
    a, _ = cable.end_tension(np.array([[20.0, 0.0, 0.0]]), q, np.array([20 * slack]), m.c_ea[:1])
    assert -res.R[0, 0, 0] == pytest.approx(a[0, 0], rel=1e-9)
    assert res.R[0, 0, 1] == pytest.approx(w * 20 * slack / 2, rel=1e-9)
    # the middle node sits at the lowest point of the single catenary: integrate its shape to s = L0 / 2
    s = np.linspace(0.0, 10 * slack, 200001)
    t = a[0][None, :] - q * s[:, None]
    tn = np.linalg.norm(t, axis=1)
    f = t / tn[:, None] * (1.0 + tn / m.c_ea[0])[:, None]
    mid = np.sum(0.5 * (f[1:] + f[:-1]) * np.diff(s)[:, None], axis=0)
    assert mid[0] == pytest.approx(10.0, abs=1e-6)
    assert res.U[0, 2, 1] == pytest.approx(mid[1], rel=1e-6) and res.U[0, 2, 1] < -2.5
    for slack2 in (1.001, 1.3):
        doc = cable_doc(nodes, [_cab(i, i, i + 1, unstretched_length=5 * slack2, weight=w) for i in (1, 2, 3, 4)])
        assert fers_oracle.analyze(wire.parse_document(doc)).converged.all()
    # four cables to a loaded node; a stay whose pretension is overcome by uplift
    b = 6.0
    net = [(1, -b, 0, -b, 2), (2, b, 0, -b, 2), (3, b, 0, b, 2), (4, -b, 0, b, 2), (5, 0, -1.0, 0, None)]
    L = np.sqrt(2 * b * b + 1.0)
    r_net = fers_oracle.analyze(wire.parse_document(cable_doc(net, [_cab(i, i, 5, unstretched_length=1.02 * L, weight=8.0) for i in (1, 2, 3, 4)],
                                                              loads=[(1, 5, 5e3, (0, -1, 0))])))
    assert r_net.converged.all() and r_net.R[0, :4, 1].sum() == pytest.approx(5e3 + 4 * 8.0 * 1.02 * L, rel=1e-9)
    stay = cable_doc([(1, 0, 0, 0, 1), (2, 10, 0, 0, None), (3, 0, 6, 0, 2)],
                     [dict(id=1, start_node_id=1, end_node_id=2), _cab(2, 3, 2, pretension=100.0)], loads=[(1, 2, 8e3, (0, 1, 0))])
    r_stay = fers_oracle.analyze(wire.parse_document(stay))
    assert r_stay.converged.all() and np.abs(r_stay.f_global[0, 1, :3]).max() < 1e-3      # slack: the stay carries nothing


def _mast_doc(order="LINEAR", self_weight=True):
    """A 12 m clamped mast with two guys in its plane and one out of it, a lateral load, a uniform load on one guy; three
    combinations."""
    nodes = [(1, 0, 0, 0, 1), (2, 0, 6, 0, None), (3, 0, 12, 0, None), (4, -8, 0, 0, 2), (5, 8, 0, 0, 2), (6, 0, 0, 9, 2)]
    members = [dict(id=1, start_node_id=1, end_node_id=2), dict(id=2, start_node_id=2, end_node_id=3),
               _cab(3, 4, 3, pretension=1e3, weight=5.0), _cab(4, 5, 3, unstretched_length=14.44, weight=5.0),
               _cab(5, 6, 3, unstretched_length=15.02, weight=5.0 if self_weight else 0.0)]
    # (weak axis of the mast: E I = 2.1e5, Euler load of the guyed 12 m mast ~ 29 kN: the loads stay well below; the cables
    # share the mast's section, EA = 4e8, and weigh 5 N/m through the `weight` override)
    return cable_doc(nodes, members, loads=[(1, 3, 800.0, (1, 0, 0)), (2, 2, 500.0, (0, 0, -1)), (2, 3, 2e3, (0, -1, 0))],
                     dls=[(2, 4, 20.0, (0, -1, 0), 0.0, 1.0)], order=order, self_weight=self_weight,
                     combos=[{1: 1.0}, {1: 1.5, 2: 1.0}, {2: 1.35}])


@pytest.mark.parametrize("order", ["LINEAR", "NONLINEAR"])
def test_guyed_mast_is_in_equilibrium(order):
    m = wire.parse_document(_mast_doc(order))
    res = fers_oracle.analyze(m)
    assert res.converged.all()
    plan_cols = res.U.shape[0]
    for col in range(plan_cols):
        # what the cables hand to their end nodes plus what they weigh is in balance with the rest of the structure:
        # sum of reactions + applied loads (incl. the cables' own weight and their uniform loads) = 0
        fac = fers_oracle.rhs_plan(m).factors[col]
        applied = np.zeros(3)
        for li, lc in enumerate(m.load_cases):
            applied += fac[li] * (lc.nod_vec[:, :3].sum(0) + (lc.cab_q * m.c_l0[:, None]).sum(0))
            for mi, row in zip(lc.dl_member, lc.dl_data):
                L = np.linalg.norm(m.xyz[m.m_n2[mi]] - m.xyz[m.m_n1[mi]])
                applied += fac[li] * 0.5 * (row[0:3] + row[3:6]) * (row[7] - row[6]) * L
        applied += (m.c_w * m.c_l0[:, None]).sum(0)
        total = res.R[col, :, :3].sum(0) + applied
        assert np.abs(total).max() < 1e-6 * max(np.abs(applied).max(), 1.0), (col, total, applied)
    # the guys pull: start tension along the chord
    for c, mi in enumerate(m.c_member):
        chord = m.xyz[m.m_n2[mi]] - m.xyz[m.m_n1[mi]]
        assert (-res.f_global[:, mi, 0:3] @ chord > 0.0).all()


# ---- GPU -----------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["hanging", "string_linear", "string_nonlinear", "mast_linear", "mast_nonlinear", "mast_no_weight", "slack"])
def test_gpu_matches_oracle(engine_mod, name):
    doc = {"hanging": lambda: _hanging()[0], "string_linear": lambda: _string_doc(order="LINEAR"),
           "string_nonlinear": lambda: _string_doc(order="NONLINEAR"), "mast_linear": lambda: _mast_doc("LINEAR"),
           "mast_nonlinear": lambda: _mast_doc("NONLINEAR"), "mast_no_weight": lambda: _mast_doc("LINEAR", False),
           "slack": _slack_doc}[name]()
    m = wire.parse_document(doc)
    got = engine_mod.analyze(m)
    want = fers_oracle.analyze(m)
    assert got.converged.all() and want.converged.all()
    # both sides stop their Newton iteration at a relative step of 1e-9: the converged states agree far below that
    helpers.assert_results_close(got, want, rtol=1e-8, label=name)


@pytest.mark.gpu
def test_gpu_pretensioned_string_closed_form(engine_mod):
    b, n0, p = 5.0, 20e3, 3e3
    doc = _string_doc(b, n0, p)
    ea = _ea(doc)
    l0 = b / (1.0 + n0 / ea)
    delta = brentq(lambda d: 2.0 * ea * (np.hypot(b, d) - l0) / l0 * d / np.hypot(b, d) - p, 0.0, b, xtol=1e-15)
    res = engine_mod.analyze(wire.parse_document(doc))
    assert res.U[0, 1, 1] == pytest.approx(-delta, rel=1e-8)
    assert res.R[0, 0, 1] == pytest.approx(p / 2, rel=1e-8)
