"""Generate tests/golden/wire_cases.json.gz with the reference's OWN front end.

Run in the build container only (needs /root/reference):

    PYTHONPATH=tests/refshim:/root/reference python tests/golden/make_golden.py

Every case is a model built through the unmodified ``fers_core`` objects (Node, Member,
Section, NodalSupport, loads, imperfections, plates ...), serialised by ``FERS.to_dict()``
(fers_core/fers/fers.py:315-361) -- i.e. the exact wire document the reference would hand to
``fers_calculations.calculate_from_json`` -- plus the expectations the reference's tests hold
for that model: closed-form beam theory (tests/common_functions.py:90-246,
tests/beam_equation_checks/test_beam_equations.py, tests/functionality/test_functionality.py)
and the NAFEMS known-answer numbers (tests/nafems/benchmarks/LE1_result.json, LE5_result.json, LE6_result.json).
The GPU box has no /root/reference, so the documents travel as this fixture.
"""
import gzip
import json
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/tests/nafems")
sys.path.insert(0, "/root/reference/tests/nafems/benchmarks")

from fers_core import (  # noqa: E402
    FERS, Member, MemberHinge, MemberSet, Material, NodalLoad, NodalSupport, Node, Section, SupportCondition,
    DistributedLoad,
)
from fers_core.settings.enums import AnalysisOrder  # noqa: E402

E, G = 210e9, 80.769e9
IZ, IY, A, J = 13.21e-6, 1.01e-6, 0.00196, 0.027e-6


def steel():
    return Material(name="Steel", e_mod=E, g_mod=G, density=7850.0, yield_stress=235e6)


def ipe180(mat, shear=False):
    kw = dict(name="IPE 180", material=mat, i_y=IY, i_z=IZ, j=J, area=A)
    s = Section(**kw)
    if shear:
        s.a_sy, s.a_sz = 1.1e-3, 0.9e-3
    return s


def reset():
    for cls in (Node, Member, Section, Material, NodalSupport, MemberSet, MemberHinge):
        for name in dir(cls):
            if name.startswith("reset_counter"):
                getattr(cls, name)()
    try:
        from fers_core.loads.loadcase import LoadCase
        from fers_core.loads.loadcombination import LoadCombination
        from fers_core.loads.nodalload import NodalLoad as NL
        LoadCase.reset_counter(); LoadCombination.reset_counter()
        if hasattr(NL, "reset_counter"):
            NL.reset_counter()
    except Exception:
        pass


CASES = []


def emit(name, calc, expect, mutate=None):
    d = calc.to_dict()
    d["results"] = None
    if mutate:
        mutate(d)
    CASES.append({"name": name, "doc": d, "expect": expect})


def X(path, value, abs_tol, rel_tol=1e-5):
    return {"path": path, "value": value, "abs": abs_tol, "rel": rel_tol}


def linear(calc):
    calc.settings.analysis_options.order = AnalysisOrder.LINEAR


# 1 cantilever, tip load (default NONLINEAR order, as the reference test runs it)
reset()
c = FERS(); s = ipe180(steel())
n1, n2 = Node(0, 0, 0), Node(5.0, 0, 0); n1.nodal_support = NodalSupport()
m = Member(start_node=n1, end_node=n2, section=s); c.add_member_set(MemberSet(members=[m]))
lc = c.create_load_case(name="End Load")
NodalLoad(node=n2, load_case=lc, magnitude=1000.0, direction=(0, -1, 0))
emit("cantilever_end_load", c, [
    X("loadcases/End Load/displacement_nodes/2/dy", -1000 * 125 / (3 * E * IZ), 1e-6),
    X("loadcases/End Load/reaction_nodes/1/nodal_forces/mz", 5000.0, 1e-3),
    X("loadcases/End Load/reaction_nodes/1/nodal_forces/fy", 1000.0, 1e-6),
    X("loadcases/End Load/member_results/1/section_forces/1/forces/mz", 2500.0, 1e-3),
])

# 2 cantilever, partial UDL 0.4-0.7 L
reset()
c = FERS(); linear(c); s = ipe180(steel())
n1, n2 = Node(0, 0, 0), Node(5.0, 0, 0); n1.nodal_support = NodalSupport()
m = Member(start_node=n1, end_node=n2, section=s); c.add_member_set(MemberSet(members=[m]))
lc = c.create_load_case(name="Partial")
DistributedLoad(member=m, load_case=lc, magnitude=1000.0, direction=(0, -1, 0), start_frac=0.4, end_frac=0.7)
W = 1000.0 * 0.3 * 5.0
emit("cantilever_partial_udl", c, [
    X("loadcases/Partial/reaction_nodes/1/nodal_forces/fy", W, 1e-6),
    X("loadcases/Partial/reaction_nodes/1/nodal_forces/mz", W * 0.55 * 5.0, 1e-3),
])

# 3 cantilever, partial triangular 0 -> 1000 over 0.3-0.7 L
reset()
c = FERS(); linear(c); s = ipe180(steel())
n1, n2 = Node(0, 0, 0), Node(5.0, 0, 0); n1.nodal_support = NodalSupport()
m = Member(start_node=n1, end_node=n2, section=s); c.add_member_set(MemberSet(members=[m]))
lc = c.create_load_case(name="Tri")
DistributedLoad(member=m, load_case=lc, magnitude=0.0, end_magnitude=1000.0, direction=(0, -1, 0),
                start_frac=0.3, end_frac=0.7)
W = 0.5 * 1000.0 * 0.4 * 5.0
xbar = (0.3 + 0.4 * 2.0 / 3.0) * 5.0
emit("cantilever_partial_triangular", c, [
    X("loadcases/Tri/reaction_nodes/1/nodal_forces/fy", W, 1e-6),
    X("loadcases/Tri/reaction_nodes/1/nodal_forces/mz", W * xbar, 1e-3),
])

# 4 cantilever, end moment about +Z
reset()
from fers_core import NodalMoment  # noqa: E402
c = FERS(); linear(c); s = ipe180(steel())
n1, n2 = Node(0, 0, 0), Node(5.0, 0, 0); n1.nodal_support = NodalSupport()
m = Member(start_node=n1, end_node=n2, section=s); c.add_member_set(MemberSet(members=[m]))
lc = c.create_load_case(name="Moment")
NodalMoment(node=n2, load_case=lc, magnitude=500.0, direction=(0, 0, 1))
emit("cantilever_end_moment", c, [
    X("loadcases/Moment/displacement_nodes/2/dy", 500.0 * 25 / (2 * E * IZ), 1e-6),
    X("loadcases/Moment/displacement_nodes/2/rz", 500.0 * 5 / (E * IZ), 1e-9),
])

# 5 simply supported, UDL, two members
reset()
c = FERS(); linear(c); s = ipe180(steel())
L = 6.0
n1, n2, n3 = Node(0, 0, 0), Node(L / 2, 0, 0), Node(L, 0, 0)
n1.nodal_support = NodalSupport(rotation_conditions={"X": SupportCondition.fixed(), "Y": SupportCondition.free(),
                                                     "Z": SupportCondition.free()})
n3.nodal_support = NodalSupport(displacement_conditions={"X": SupportCondition.free(), "Y": SupportCondition.fixed(),
                                                         "Z": SupportCondition.fixed()},
                                rotation_conditions={"X": SupportCondition.free(), "Y": SupportCondition.free(),
                                                     "Z": SupportCondition.free()})
m1 = Member(start_node=n1, end_node=n2, section=s); m2 = Member(start_node=n2, end_node=n3, section=s)
c.add_member_set(MemberSet(members=[m1, m2]))
lc = c.create_load_case(name="UDL")
w = 2000.0
DistributedLoad(member=m1, load_case=lc, magnitude=w, direction=(0, -1, 0))
DistributedLoad(member=m2, load_case=lc, magnitude=w, direction=(0, -1, 0))
emit("simply_supported_udl", c, [
    X("loadcases/UDL/displacement_nodes/2/dy", -5 * w * L**4 / (384 * E * IZ), 1e-6),
    X("loadcases/UDL/member_results/1/end_node_forces/mz", w * L * L / 8, 1e-3),
    X("loadcases/UDL/reaction_nodes/1/nodal_forces/fy", w * L / 2, 1e-6),
])

# 6 fixed-fixed, centre load, Timoshenko section (shear pairing SURVEY.md A.5)
reset()
c = FERS(); linear(c); s = ipe180(steel(), shear=True)
L = 6.0; P = 20000.0
n1, n2, n3 = Node(0, 0, 0), Node(L / 2, 0, 0), Node(L, 0, 0)
n1.nodal_support = NodalSupport(); n3.nodal_support = NodalSupport()
m1 = Member(start_node=n1, end_node=n2, section=s); m2 = Member(start_node=n2, end_node=n3, section=s)
c.add_member_set(MemberSet(members=[m1, m2]))
lc = c.create_load_case(name="Centre")
NodalLoad(node=n2, load_case=lc, magnitude=P, direction=(0, -1, 0))
emit("fixed_fixed_timoshenko", c, [
    X("loadcases/Centre/displacement_nodes/2/dy", -(P * L**3 / (192 * E * IZ) + P * L / (4 * G * 0.9e-3)), 1e-7),
    X("loadcases/Centre/reaction_nodes/1/nodal_forces/mz", P * L / 8, 1e-3),
])

# 7 rigid link (reference test_041)
reset()
c = FERS(); s = ipe180(steel())
n1, n2, n3 = Node(0, 0, 0), Node(5, 0, 0), Node(10, 0, 0); n1.nodal_support = NodalSupport()
me = Member(start_node=n1, end_node=n2, section=s); mr = Member(start_node=n2, end_node=n3, member_type="RIGID")
c.add_member_set(MemberSet(members=[me, mr]))
lc = c.create_load_case(name="End Load")
NodalLoad(node=n2, load_case=lc, magnitude=1000.0, direction=(0, -1, 0))
rz = -1000.0 * 25 / (2 * E * IZ)
emit("rigid_link", c, [
    X("loadcases/End Load/displacement_nodes/2/rz", rz, 1e-9),
    X("loadcases/End Load/displacement_nodes/3/rz", rz, 1e-9),
    X("loadcases/End Load/displacement_nodes/3/dy", -1000 * 125 / (3 * E * IZ) + rz * 5.0, 1e-6),
])

# 8 hinge with rotational spring (reference test_051)
reset()
c = FERS(); s = ipe180(steel())
n1, n2, n3 = Node(0, 0, 0), Node(2.5, 0, 0), Node(5, 0, 0); n1.nodal_support = NodalSupport()
kphi = 1000.0 * 2.5 / 0.1
hinge = MemberHinge(hinge_type="SPRING_Z", rotational_release_mz=kphi)
mr = Member(start_node=n1, end_node=n2, member_type="RIGID")
mf = Member(start_node=n2, end_node=n3, section=s, start_hinge=hinge)
c.add_member_set(MemberSet(members=[mr, mf]))
lc = c.create_load_case(name="End Load")
NodalLoad(node=n3, load_case=lc, magnitude=1000.0, direction=(0, -1, 0))
emit("hinge_spring", c, [
    X("loadcases/End Load/displacement_nodes/3/dy", -(1000 * 2.5**3 / (3 * E * IZ) + 0.1 * 2.5), 1e-6),
    X("loadcases/End Load/displacement_nodes/3/rz", -(1000 * 2.5**2 / (2 * E * IZ) + 0.1), 1e-6),
    X("loadcases/End Load/reaction_nodes/1/nodal_forces/mz", 5000.0, 1e-3),
])

# 9 tension-only members (reference test_061)
reset()
c = FERS(); s = ipe180(steel())
n1, n2, n3 = Node(0, 0, 0), Node(2.5, 0, 0), Node(5, 0, 0)
n1.nodal_support = NodalSupport(); n3.nodal_support = NodalSupport()
n2.nodal_support = NodalSupport(displacement_conditions={"X": SupportCondition.free(), "Y": SupportCondition.fixed(),
                                                         "Z": SupportCondition.fixed()})
ml = Member(start_node=n1, end_node=n2, section=s, member_type="TENSION")
mr = Member(start_node=n2, end_node=n3, section=s, member_type="TENSION")
c.add_member_set(MemberSet(members=[ml, mr]))
lc = c.create_load_case(name="Mid Load")
NodalLoad(node=n2, load_case=lc, magnitude=10000.0, direction=(1, 0, 0))
emit("tension_only", c, [
    X("loadcases/Mid Load/displacement_nodes/2/dx", 10000.0 * 2.5 / (A * E), 1e-6),
    X("loadcases/Mid Load/reaction_nodes/1/nodal_forces/fx", -10000.0, 1e-3, 1e-4),
    X("loadcases/Mid Load/reaction_nodes/3/nodal_forces/fx", 0.0, 1e-3),
])

# 10 3D frame: skew members, roll angle, several load cases + combinations, local-axis member loads
reset()
c = FERS(); linear(c); mat = steel(); s = ipe180(mat, shear=True); s2 = ipe180(mat)
nodes = [Node(0, 0, 0), Node(0, 3.2, 0), Node(4.0, 3.5, 0.8), Node(4.0, 0, 0.8), Node(2.0, 5.0, 3.0), Node(0, 3.2, 3.0),
         Node(0, 0, 3.0)]
for i in (0, 3, 6):
    nodes[i].nodal_support = NodalSupport()
mem = [Member(start_node=nodes[0], end_node=nodes[1], section=s),
       Member(start_node=nodes[1], end_node=nodes[2], section=s2, rotation_angle=30.0),
       Member(start_node=nodes[3], end_node=nodes[2], section=s),
       Member(start_node=nodes[2], end_node=nodes[4], section=s2, rotation_angle=-75.0),
       Member(start_node=nodes[1], end_node=nodes[4], section=s),
       Member(start_node=nodes[5], end_node=nodes[4], section=s2),
       Member(start_node=nodes[6], end_node=nodes[5], section=s),
       Member(start_node=nodes[1], end_node=nodes[5], section=s2, rotation_angle=12.0)]
c.add_member_set(MemberSet(members=mem))
lc1 = c.create_load_case(name="G"); lc2 = c.create_load_case(name="W"); lc3 = c.create_load_case(name="P")
for mm in (mem[1], mem[3], mem[4], mem[5], mem[7]):
    DistributedLoad(member=mm, load_case=lc1, magnitude=3000.0, direction=(0, -1, 0))
DistributedLoad(member=mem[0], load_case=lc2, magnitude=800.0, end_magnitude=1500.0, direction=(1, 0, 0.3),
                start_frac=0.1, end_frac=0.9)
NodalLoad(node=nodes[4], load_case=lc2, magnitude=2500.0, direction=(0.6, 0, 0.8))
NodalMoment(node=nodes[2], load_case=lc3, magnitude=700.0, direction=(0, 1, 0))
NodalLoad(node=nodes[5], load_case=lc3, magnitude=-4000.0, direction=(0, 1, 0))
c.create_load_combination(name="ULS", load_cases_factors={lc1: 1.35, lc2: 1.5, lc3: 1.05}, situation="ULS", check="ALL")
c.create_load_combination(name="SLS", load_cases_factors={lc1: 1.0, lc2: 0.7}, situation="SLS", check="ALL")


def add_member_point_loads(d):
    d["analysis"]["load_cases"][2]["member_point_loads"] = [
        {"id": 1, "member": 2, "magnitude": 1200.0, "direction": [0, -1, 0], "position": 0.35, "axes": "global"},
        {"id": 2, "member": 4, "magnitude": 900.0, "direction": [0, 0, 1], "position": 0.6, "axes": "local"}]
    d["analysis"]["load_cases"][2]["member_point_moments"] = [
        {"id": 1, "member": 5, "magnitude": 300.0, "direction": [0, 1, 0], "position": 0.5, "axes": "local"}]


emit("space_frame_3d", c, [], add_member_point_loads)

# 11 2D portal frame, load cases + combinations (BASELINE.json configs[1])
reset()
c = FERS(); linear(c); mat = steel(); s = ipe180(mat)
from fers_core.settings.enums import Dimensionality  # noqa: E402
c.settings.analysis_options.dimensionality = Dimensionality.TWO_DIMENSIONAL
h, b = 4.0, 8.0
p = [Node(0, 0, 0), Node(0, h, 0), Node(b / 2, h + 1.0, 0), Node(b, h, 0), Node(b, 0, 0)]
pin = dict(rotation_conditions={"X": SupportCondition.fixed(), "Y": SupportCondition.fixed(), "Z": SupportCondition.free()})
p[0].nodal_support = NodalSupport(**pin); p[4].nodal_support = NodalSupport(**pin)
pm = [Member(start_node=p[0], end_node=p[1], section=s), Member(start_node=p[1], end_node=p[2], section=s),
      Member(start_node=p[2], end_node=p[3], section=s), Member(start_node=p[4], end_node=p[3], section=s)]
c.add_member_set(MemberSet(members=pm))
g = c.create_load_case(name="G"); sn = c.create_load_case(name="S"); wd = c.create_load_case(name="W")
for mm in pm[1:3]:
    DistributedLoad(member=mm, load_case=g, magnitude=1500.0, direction=(0, -1, 0))
    DistributedLoad(member=mm, load_case=sn, magnitude=2200.0, direction=(0, -1, 0))
DistributedLoad(member=pm[0], load_case=wd, magnitude=900.0, direction=(1, 0, 0))
NodalLoad(node=p[1], load_case=wd, magnitude=1800.0, direction=(1, 0, 0))
c.create_load_combination(name="ULS1", load_cases_factors={g: 1.35, sn: 1.5}, situation="ULS", check="ALL")
c.create_load_combination(name="ULS2", load_cases_factors={g: 1.35, sn: 0.75, wd: 1.5}, situation="ULS", check="ALL")
c.create_load_combination(name="SLS", load_cases_factors={g: 1.0, sn: 1.0}, situation="SLS", check="ALL")
tot = 1500.0 * 2 * math.hypot(b / 2, 1.0)
emit("portal_2d_combinations", c, [
    X("loadcases/G/reaction_nodes/1/nodal_forces/fy", tot / 2, 1e-6),
    X("loadcases/G/reaction_nodes/5/nodal_forces/fy", tot / 2, 1e-6),
])

# 12 3D portal, sway imperfection, P-Delta (consistent and simplified / suppressed axis)
for variant in ("consistent", "simplified_suppress_z"):
    reset()
    from fers_core import ImperfectionCase, SwayImperfection  # noqa: E402
    c = FERS(); mat = steel(); s = ipe180(mat)
    q = [Node(0, 0, 0), Node(0, 4, 0), Node(5, 4, 0), Node(5, 0, 0), Node(0, 0, 4), Node(0, 4, 4), Node(5, 4, 4), Node(5, 0, 4)]
    for i in (0, 3, 4, 7):
        q[i].nodal_support = NodalSupport()
    qm = [Member(start_node=q[0], end_node=q[1], section=s), Member(start_node=q[1], end_node=q[2], section=s),
          Member(start_node=q[3], end_node=q[2], section=s), Member(start_node=q[4], end_node=q[5], section=s),
          Member(start_node=q[5], end_node=q[6], section=s), Member(start_node=q[7], end_node=q[6], section=s),
          Member(start_node=q[1], end_node=q[5], section=s), Member(start_node=q[2], end_node=q[6], section=s)]
    c.add_member_set(MemberSet(members=qm))
    gv = c.create_load_case(name="V"); hh = c.create_load_case(name="H")
    for nd_ in (q[1], q[2], q[5], q[6]):
        NodalLoad(node=nd_, load_case=gv, magnitude=60000.0, direction=(0, -1, 0))
    NodalLoad(node=q[1], load_case=hh, magnitude=3000.0, direction=(1, 0, 0))
    NodalLoad(node=q[5], load_case=hh, magnitude=1000.0, direction=(0, 0, 1))
    combo = c.create_load_combination(name="VH", load_cases_factors={gv: 1.0, hh: 1.0}, situation="ULS", check="ALL")
    combo2 = c.create_load_combination(name="V2H", load_cases_factors={gv: 1.35, hh: 0.5}, situation="ULS", check="ALL")
    imp = ImperfectionCase(loadcombinations=[combo])
    imp.add_sway_imperfection(SwayImperfection(magnitude=1.0 / 200.0, axis=(0, 0, 1), height_direction=(0, 1, 0),
                                               reference_point=(0, 0, 0)))
    c.add_imperfection_case(imp)

    def mut(d, variant=variant):
        o = d["analysis"]["options"]
        o["order"] = "NONLINEAR"; o["nonlinear_method"] = "P_DELTA"; o["tolerance"] = 1e-10; o["max_iterations"] = 50
        if variant != "consistent":
            o["pdelta_formulation"] = "SIMPLIFIED"; o["pdelta_suppress_axes"] = ["Z"]
    emit("portal_pdelta_" + variant, c, [], mut)

# 12b P-Delta pinned on closed forms ("unpinned at any tolerance: geometric stiffness / P-Delta iteration,
# imperfection application").  The expected numbers come from the self-contained plane beam-column solver below
# (textbook element matrices, its own fixed-point loop, perturbed coordinates formed HERE) and from the classical
# beam-column formulas -- nothing of fers_core_b200 / oracle is imported by this generator.
import numpy as _np  # noqa: E402


def plane_pdelta(xy, EI, EA, loads, formulation="CONSISTENT", tol=1e-15, max_it=200):
    """Cantilever chain of plane frame elements through the points xy [n+1,2] (node 0 clamped), nodal loads
    {node: (Fx, Fy, Mz)}.  P-Delta fixed point u <- (K_e + K_g(N(u)))^-1 f with N from the axial strain of the
    previous iterate.  Returns (u [n+1,3], N [n], iterations); the reaction K u - f of the clamped node is left in
    ``plane_pdelta.base_reaction`` (Fx, Fy, Mz)."""
    xy = _np.asarray(xy, float)
    n = len(xy) - 1
    nd = 3 * (n + 1)
    f = _np.zeros(nd)
    for node, (fx, fy, mz) in loads.items():
        f[3 * node:3 * node + 3] += (fx, fy, mz)
    d = xy[1:] - xy[:-1]
    L = _np.hypot(d[:, 0], d[:, 1])
    c, s_ = d[:, 0] / L, d[:, 1] / L

    def element(e, N):
        l = L[e]
        k = _np.zeros((6, 6))
        a = EA / l
        k[0, 0] = k[3, 3] = a; k[0, 3] = k[3, 0] = -a
        b = EI / l ** 3
        kb = b * _np.array([[12, 6 * l, -12, 6 * l], [6 * l, 4 * l * l, -6 * l, 2 * l * l],
                            [-12, -6 * l, 12, -6 * l], [6 * l, 2 * l * l, -6 * l, 4 * l * l]])
        if formulation == "CONSISTENT":
            kg = N / (30 * l) * _np.array([[36, 3 * l, -36, 3 * l], [3 * l, 4 * l * l, -3 * l, -l * l],
                                           [-36, -3 * l, 36, -3 * l], [3 * l, -l * l, -3 * l, 4 * l * l]])
        else:
            kg = N / l * _np.array([[1, 0, -1, 0], [0, 0, 0, 0], [-1, 0, 1, 0], [0, 0, 0, 0.0]])
        idx = [1, 2, 4, 5]
        k[_np.ix_(idx, idx)] += kb + kg
        T = _np.zeros((6, 6))
        for o in (0, 3):
            T[o, o] = c[e]; T[o, o + 1] = s_[e]; T[o + 1, o] = -s_[e]; T[o + 1, o + 1] = c[e]; T[o + 2, o + 2] = 1.0
        return T.T @ k @ T, T

    N = _np.zeros(n)
    u = _np.zeros(nd)
    for it in range(max_it):
        K = _np.zeros((nd, nd))
        for e in range(n):
            ke, _ = element(e, N[e])
            sl = slice(3 * e, 3 * e + 6)
            K[sl, sl] += ke
        un = _np.zeros(nd)
        un[3:] = _np.linalg.solve(K[3:, 3:], f[3:])
        Nn = _np.zeros(n)
        for e in range(n):
            _, T = element(e, 0.0)
            ul = T @ un[3 * e:3 * e + 6]
            Nn[e] = EA / L[e] * (ul[3] - ul[0])
        den = _np.linalg.norm(un)
        done = it > 0 and _np.linalg.norm(un - u) <= tol * den
        u, N = un, Nn
        plane_pdelta.base_reaction = (K @ un - f)[:3]
        if done:
            break
    return u.reshape(n + 1, 3), N, it + 1


PD_L, PD_N = 4.0, 8
PD_I = IY                         # doubly symmetric section (i_y = i_z): the local-axis rule cannot matter
PD_EI, PD_EA = E * PD_I, E * A
PD_PCR = math.pi ** 2 * PD_EI / (4 * PD_L ** 2)        # Euler load of the cantilever
PD_H = 100.0
PD_PHI = 1.0 / 200.0


def pd_column(n_el, formulation, P, H, lean=0.0, sway=None, order_linear=False, name_lc="PH"):
    """Vertical cantilever column along +Y (n_el members), tip loads P (down) and H (+X); optional sway imperfection."""
    reset()
    from fers_core import ImperfectionCase, SwayImperfection
    c = FERS(); mat = steel()
    s = Section(name="SYM", material=mat, i_y=PD_I, i_z=PD_I, j=J, area=A)
    nodes = [Node(0.0, PD_L * i / n_el, 0.0) for i in range(n_el + 1)]
    nodes[0].nodal_support = NodalSupport()
    mem = [Member(start_node=nodes[i], end_node=nodes[i + 1], section=s) for i in range(n_el)]
    c.add_member_set(MemberSet(members=mem))
    lc = c.create_load_case(name=name_lc)
    if P:
        NodalLoad(node=nodes[-1], load_case=lc, magnitude=P, direction=(0, -1, 0))
    if H:
        NodalLoad(node=nodes[-1], load_case=lc, magnitude=H, direction=(1, 0, 0))
    combo = c.create_load_combination(name="C", load_cases_factors={lc: 1.0}, situation="ULS", check="ALL")
    if sway is not None:
        imp = ImperfectionCase(loadcombinations=[combo])
        imp.add_sway_imperfection(SwayImperfection(magnitude=sway, axis=(0, 0, 1), height_direction=(0, 1, 0),
                                                   reference_point=(0, 0, 0)))
        c.add_imperfection_case(imp)

    def mut(d):
        o = d["analysis"]["options"]
        o["order"] = "NONLINEAR"; o["nonlinear_method"] = "P_DELTA"; o["tolerance"] = 1e-12; o["max_iterations"] = 50
        o["pdelta_formulation"] = formulation
        o["include_shear_deformation"] = False
    return c, mut, len(nodes)


def amplification(P):
    """Tip deflection of the cantilever beam-column over H L^3 / 3 EI: 3 (tan kL - kL) / (kL)^3, k^2 = P / EI."""
    kl = math.sqrt(P / PD_EI) * PD_L
    return 3.0 * (math.tan(kl) - kl) / kl ** 3


# (a) consistent K_g, 8 elements, P = 0.5 P_cr: the plane solver to 1e-9, the beam-column formula to 1e-5
P = 0.5 * PD_PCR
xy = [(0.0, PD_L * i / PD_N) for i in range(PD_N + 1)]
u, Nax, _ = plane_pdelta(xy, PD_EI, PD_EA, {PD_N: (PD_H, -P, 0.0)})
c, mut, ntip = pd_column(PD_N, "CONSISTENT", P, PD_H)
d_exact = PD_H * PD_L ** 3 / (3 * PD_EI) * amplification(P)
assert abs(u[-1, 0] / d_exact - 1.0) < 2e-6, (u[-1, 0], d_exact)
emit("pdelta_beam_column_consistent", c, [
    X(f"loadcombinations/C/displacement_nodes/{ntip}/dx", float(u[-1, 0]), 0.0, 1e-9),
    X(f"loadcombinations/C/displacement_nodes/{ntip}/dx", d_exact, 0.0, 1e-5),
    X(f"loadcombinations/C/displacement_nodes/{ntip}/dy", float(u[-1, 1]), 0.0, 1e-9),
    X(f"loadcombinations/C/displacement_nodes/{ntip}/rz", float(u[-1, 2]), 0.0, 1e-9),
    # equilibrium in the displaced position: M_base = H L + P delta
    X("loadcombinations/C/reaction_nodes/1/nodal_forces/mz", PD_H * PD_L + P * float(u[-1, 0]), 0.0, 1e-9),
    X("loadcombinations/C/reaction_nodes/1/nodal_forces/fx", -PD_H, 0.0, 1e-9),
    X("loadcombinations/C/reaction_nodes/1/nodal_forces/fy", P, 0.0, 1e-9),
    X("loadcombinations/C/member_results/1/local_start_forces/fx", P, 0.0, 1e-9),      # compression: +P at the start face (stiffness sign)
], mut)

# (b) simplified K_g ("P/L-only diagonal terms"), ONE element: delta = H / (3 EI / L^3 - P / L) exactly
P = 0.3 * PD_PCR
c, mut, ntip = pd_column(1, "SIMPLIFIED", P, PD_H)
d_simpl = PD_H / (3 * PD_EI / PD_L ** 3 - P / PD_L)
u1, _, _ = plane_pdelta([(0.0, 0.0), (0.0, PD_L)], PD_EI, PD_EA, {1: (PD_H, -P, 0.0)}, "SIMPLIFIED")
assert abs(u1[-1, 0] / d_simpl - 1.0) < 1e-12
emit("pdelta_simplified_one_element", c, [
    X(f"loadcombinations/C/displacement_nodes/{ntip}/dx", d_simpl, 0.0, 1e-9),
    X("loadcombinations/C/reaction_nodes/1/nodal_forces/mz", PD_H * PD_L + P * d_simpl, 0.0, 1e-9),
], mut)
# (b') the same column with the consistent matrix: one element, closed form of the 2x2 system
c, mut, ntip = pd_column(1, "CONSISTENT", P, PD_H)
k11 = 12 * PD_EI / PD_L ** 3 - 6 * P / (5 * PD_L); k12 = -6 * PD_EI / PD_L ** 2 + P / 10
k22 = 4 * PD_EI / PD_L - 2 * P * PD_L / 15
d_cons = PD_H / (k11 - k12 * k12 / k22)
assert d_simpl < d_cons < PD_H * PD_L ** 3 / (3 * PD_EI) * amplification(P)      # simplified < consistent < exact
emit("pdelta_consistent_one_element", c, [
    X(f"loadcombinations/C/displacement_nodes/{ntip}/dx", d_cons, 0.0, 1e-9),
], mut)

# (c) sway imperfection phi about Z (lean towards -X: axis x height_direction = Z x Y = -X), vertical load only.
# Expected from the plane solver on coordinates perturbed HERE (x = -phi y); and the notional-load twin H = phi P.
P = 0.5 * PD_PCR
xy_lean = [(-PD_PHI * PD_L * i / PD_N, PD_L * i / PD_N) for i in range(PD_N + 1)]
ul, _, _ = plane_pdelta(xy_lean, PD_EI, PD_EA, {PD_N: (0.0, -P, 0.0)})
r_lean = plane_pdelta.base_reaction.copy()
c, mut, ntip = pd_column(PD_N, "CONSISTENT", P, 0.0, sway=PD_PHI)
d_lean_exact = -PD_PHI * P * PD_L ** 3 / (3 * PD_EI) * amplification(P)          # notional load phi P on the straight column
assert abs(ul[-1, 0] / d_lean_exact - 1.0) < 2e-4, (ul[-1, 0], d_lean_exact)
emit("pdelta_sway_imperfection", c, [
    X(f"loadcombinations/C/displacement_nodes/{ntip}/dx", float(ul[-1, 0]), 0.0, 1e-9),
    X(f"loadcombinations/C/displacement_nodes/{ntip}/dx", d_lean_exact, 0.0, 2e-4),
    X(f"loadcombinations/C/displacement_nodes/{ntip}/dy", float(ul[-1, 1]), 0.0, 1e-9),
    # base moment: the plane solver's own reaction; = -P x (lean + deflection) up to the second-order terms the
    # linearised theory drops (the load sits phi L - dx to the left of the base)
    X("loadcombinations/C/reaction_nodes/1/nodal_forces/mz", float(r_lean[2]), 0.0, 1e-9),
    X("loadcombinations/C/reaction_nodes/1/nodal_forces/mz", -P * (PD_PHI * PD_L - float(ul[-1, 0])), 0.0, 1e-4),
    # the load case alone is solved on the PERFECT geometry (imperfections attach to combinations): no sway
    X(f"loadcases/PH/displacement_nodes/{ntip}/dx", 0.0, 1e-15, 0.0),
], mut)
un, _, _ = plane_pdelta(xy, PD_EI, PD_EA, {PD_N: (-PD_PHI * P, -P, 0.0)})
c, mut, ntip = pd_column(PD_N, "CONSISTENT", P, -PD_PHI * P)
emit("pdelta_notional_load", c, [
    X(f"loadcombinations/C/displacement_nodes/{ntip}/dx", float(un[-1, 0]), 0.0, 1e-9),
    X(f"loadcombinations/C/displacement_nodes/{ntip}/dx", float(ul[-1, 0]), 0.0, 2e-4),     # = the sway imperfection
], mut)

# (d) above the Euler load the tangent matrix is indefinite: the engine must refuse, not return a number
c, mut, ntip = pd_column(PD_N, "CONSISTENT", 1.05 * PD_PCR, PD_H)
CASES.append({"name": "pdelta_above_euler_load", "doc": (lambda d: (mut(d), d.__setitem__("results", None), d)[-1])(c.to_dict()),
              "expect": [], "expect_error": "positive definite"})
# just below it the deflection follows the formula (amplification ~ 1 / (1 - P / P_cr) = 20)
P = 0.95 * PD_PCR
u95, _, _ = plane_pdelta(xy, PD_EI, PD_EA, {PD_N: (PD_H, -P, 0.0)})
c, mut, ntip = pd_column(PD_N, "CONSISTENT", P, PD_H)
emit("pdelta_near_euler_load", c, [
    X(f"loadcombinations/C/displacement_nodes/{ntip}/dx", float(u95[-1, 0]), 0.0, 1e-8),
    X(f"loadcombinations/C/displacement_nodes/{ntip}/dx", PD_H * PD_L ** 3 / (3 * PD_EI) * amplification(P), 0.0, 1e-4),
], mut)

# 12c EN 1993-1-1 member check: the reference's own check_beam builder (fers_core/builders.py:152-250) on the model of its
# acceptance test (tests/functionality/test_check_beam.py:25-45): IPE300, 6 m, S235, simply supported, UDL 10 kN/m x 1.35.
# Named sections come from the catalogue file the reference ships (tests/refshim reads it where `sectionproperties` is absent).
from fers_core import check_beam  # noqa: E402

reset()
cb = check_beam(6.0, "IPE300", material="S235", support="simply_supported", udl=10000.0)
emit("ec3_check_beam_ipe300", cb, [
    X("loadcombinations/ULS/member_results/1/section_forces/1/forces/mz", -60750.0, 1e-6, 1e-9),      # M_Ed = 1.35 q L^2 / 8 (sagging)
    X("unity_check_results/0/governing/trace/label=Bending z (6.2.5)/value", 60.75 / 147.6, 0.0, 0.03),                   # "Bending z (6.2.5)" hand calculation
])
reset()
cb = check_beam(4.0, "HEA200", material="S355", support="fixed", udl=20000.0)
emit("ec3_check_beam_hea200_fixed", cb, [])

# 13 spring support + truss members (auto-restrained rotations)
reset()
c = FERS(); linear(c); mat = steel(); s = ipe180(mat)
t = [Node(0, 0, 0), Node(4, 0, 0), Node(2, 3, 0)]
t[0].nodal_support = NodalSupport(rotation_conditions={k: SupportCondition.free() for k in "XYZ"})
t[1].nodal_support = NodalSupport(displacement_conditions={"X": SupportCondition.spring(2.0e6), "Y": SupportCondition.fixed(),
                                                           "Z": SupportCondition.fixed()},
                                  rotation_conditions={k: SupportCondition.free() for k in "XYZ"})
t[2].nodal_support = NodalSupport(displacement_conditions={"X": SupportCondition.free(), "Y": SupportCondition.free(),
                                                           "Z": SupportCondition.fixed()},
                                  rotation_conditions={k: SupportCondition.free() for k in "XYZ"})
tm = [Member(start_node=t[0], end_node=t[1], section=s, member_type="TRUSS"),
      Member(start_node=t[1], end_node=t[2], section=s, member_type="TRUSS"),
      Member(start_node=t[0], end_node=t[2], section=s, member_type="TRUSS")]
c.add_member_set(MemberSet(members=tm))
lc = c.create_load_case(name="Apex")
NodalLoad(node=t[2], load_case=lc, magnitude=50000.0, direction=(0.2, -1, 0))
emit("truss_spring_support", c, [X("loadcases/Apex/reaction_nodes/2/nodal_forces/fx", None, 0.0)])

# 14 plate strip under a surface load (reference tests/functionality/test_plate_results.py, nx = 10)
reset()
from fers_core import LoadCase, Plate, SurfaceLoad  # noqa: E402
c = FERS(); linear(c); mat = steel(); sup = NodalSupport()
length, width, thick, nx = 2.0, 0.20, 0.03, 10
cols = []
for ix in range(nx + 1):
    x = length * ix / nx
    cols.append((Node(X=x, Y=0.0, Z=0.0, nodal_support=sup if ix == 0 else None),
                 Node(X=x, Y=width, Z=0.0, nodal_support=sup if ix == 0 else None)))
for ix in range(nx):
    a0, a1 = cols[ix]; b0, b1 = cols[ix + 1]
    c.add_plate(Plate(nodes=[a0, b0, b1], material=mat, thickness=thick, classification="Strip", local_x_direction=(1, 0, 0)),
                Plate(nodes=[a0, b1, a1], material=mat, thickness=thick, classification="Strip", local_x_direction=(1, 0, 0)))
lc = LoadCase(name="Pressure")
SurfaceLoad(load_case=lc, polygon=[(0, 0, 0), (length, 0, 0), (length, width, 0), (0, width, 0)], magnitude=1000.0,
            direction=(0, 0, -1))
c.add_load_case(lc)
emit("plate_strip_surface_load", c, [X("loadcases/Pressure/__sum_reaction_fz", 1000.0 * length * width, 1e-6)])

# 15 NAFEMS LE1 (CST membrane) and LE6 (DSG3 Mindlin): the reference's own builders + known answers
import LE1_build  # noqa: E402
import LE6_build  # noqa: E402


def nafems_emit(name, calc, kind, point, thickness, value, abs_tol):
    d = calc.to_dict(include_results=False)
    d["results"] = None
    CASES.append({"name": name, "doc": d,
                  "expect": [{"nafems": kind, "point": list(point), "thickness": thickness, "value": value,
                              "abs": abs_tol, "rel": 0.0}]})


for (nth, ns, val, tol) in ((24, 4, 79129730.5, 0.06), (72, 8, 90160401.4, 0.06)):
    reset()
    calc, nd = LE1_build.build(nth, ns)
    nafems_emit(f"nafems_LE1_{nth}x{ns}", calc, "LE1", LE1_build.D_PT, LE1_build.T, val, tol)
for (n, val, tol) in ((8, 396102.7, 0.06), (32, 741737.1, 0.06)):
    reset()
    calc = LE6_build.build(n)
    nafems_emit(f"nafems_LE6_{n}x{n}", calc, "LE6", LE6_build.E_CENTRE, LE6_build.T, val, tol)

# 16 NAFEMS LE5 (Z-section cantilever, restrained-warping torsion through the 7th nodal DOF): the reference's own
# builder and its known answer (tests/nafems/benchmarks/LE5_result.json: 107 643 284.3 Pa, printed to 0.1 Pa)
import LE5_build  # noqa: E402

reset()
calc, _nodes = LE5_build.build(4)
d = calc.to_dict(include_results=False)
d["results"] = None
CASES.append({"name": "nafems_LE5_4el", "doc": d,
              "expect": [{"nafems": "LE5", "omega": LE5_build.OMEGA_TIP, "cw": LE5_build.CW, "value": 107643284.3,
                          "abs": 0.06, "rel": 0.0}]})

out = os.path.join(HERE, "wire_cases.json.gz")
with gzip.open(out, "wt", encoding="utf-8") as fh:
    json.dump(CASES, fh)
print(f"wrote {len(CASES)} cases -> {out} ({os.path.getsize(out)} bytes)")
