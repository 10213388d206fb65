"""Run with the reference front end on PYTHONPATH (tests/test_from_fers.py does): every golden document is rebuilt as a
FERS object graph by the reference's own ``FERS.from_dict`` and must flatten to the SAME ModelArrays through
``wire.from_fers(calc)`` as through ``wire.parse_document(calc.to_dict())``; a larger grid frame built from reference objects
is timed through both routes."""
import dataclasses
import json
import sys
import time

import numpy as np

from fers_core import FERS, DistributedLoad, Material, Member, MemberSet, NodalLoad, NodalSupport, Node, Section
import gzip
import os

from fers_core_b200 import wire

with gzip.open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wire_cases.json.gz"), "rt", encoding="utf-8") as fh:
    GOLDEN = json.load(fh)          # (tests.helpers is shadowed by the reference's own `tests` package on this path)


def same(a, b, path=""):
    if isinstance(a, np.ndarray):
        assert isinstance(b, np.ndarray) and a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True), path
    elif dataclasses.is_dataclass(a):
        for f in dataclasses.fields(a):
            same(getattr(a, f.name), getattr(b, f.name), f"{path}.{f.name}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            same(x, y, f"{path}[{i}]")
    elif isinstance(a, dict):
        assert a.keys() == b.keys(), path
        for k in a:
            same(a[k], b[k], f"{path}.{k}")
    else:
        assert a == b or (a != a and b != b), (path, a, b)


checked = 0
for case in GOLDEN:
    try:
        calc = FERS.from_dict(json.loads(json.dumps(case["doc"])))
    except Exception as e:                     # a document the reference's own from_dict does not rebuild
        print("skip", case["name"], type(e).__name__, str(e)[:80])
        continue
    want = wire.parse_document(calc.to_dict(include_results=False))
    got = wire.from_fers(calc)
    same(got, want, case["name"])
    checked += 1
assert checked >= 20, checked

# timing on a grid frame built from reference objects (12 x 12 x 8 bays: 1 521 nodes, 4 080 members)
steel = Material(name="S235", e_mod=210e9, g_mod=80.769e9, density=7850.0, yield_stress=235e6)
sec = Section(name="IPE", material=steel, i_y=1.01e-6, i_z=13.21e-6, j=2.7e-8, area=1.96e-3)
nx = nz = 12; ny = 8
grid = {}
for j in range(ny + 1):
    for k in range(nz + 1):
        for i in range(nx + 1):
            n = Node(6.0 * i, 3.5 * j, 6.0 * k)
            if j == 0:
                n.nodal_support = NodalSupport()
            grid[i, j, k] = n
mem = []
for (i, j, k), n in grid.items():
    if j < ny: mem.append(Member(start_node=n, end_node=grid[i, j + 1, k], section=sec))
    if j > 0 and i < nx: mem.append(Member(start_node=n, end_node=grid[i + 1, j, k], section=sec))
    if j > 0 and k < nz: mem.append(Member(start_node=n, end_node=grid[i, j, k + 1], section=sec))
calc = FERS()
calc.add_member_set(MemberSet(members=mem))
lc = calc.create_load_case(name="G")
for m in mem:
    if m.start_node.Y == m.end_node.Y:
        DistributedLoad(member=m, load_case=lc, magnitude=5000.0, direction=(0, -1, 0))
for (i, j, k), n in grid.items():
    if i == 0 and j > 0:
        NodalLoad(node=n, load_case=lc, magnitude=1000.0, direction=(1, 0, 0))
calc.create_load_combination(name="ULS", load_cases_factors={lc: 1.35}, situation="ULS", check="ALL")
t0 = time.perf_counter(); got = wire.from_fers(calc); t_fast = time.perf_counter() - t0
t0 = time.perf_counter(); doc = calc.to_dict(include_results=False); t_dict = time.perf_counter() - t0
t0 = time.perf_counter(); text = json.dumps(doc); back = json.loads(text); t_json = time.perf_counter() - t0
t0 = time.perf_counter(); want = wire.parse_document(back); t_parse = time.perf_counter() - t0
same(got, want, "grid")
t_gate = None
try:
    from fers_core.types.pydantic_models import FERS as Gate
    t0 = time.perf_counter(); Gate(**doc); t_gate = time.perf_counter() - t0
except Exception:
    pass
print(json.dumps({"checked": checked, "members": len(mem), "from_fers_s": t_fast, "to_dict_s": t_dict, "json_round_trip_s": t_json,
                  "parse_document_s": t_parse, "pydantic_gate_s": t_gate}))

# a small guyed mast from reference objects: cables with pretension / unstretched length, a hinge with capacities and a warping
# release, buckling restraints on the member set -- the fields the later parts of round 2 read -- through both input routes
from fers_core import BucklingRestraint, MemberHinge, MemberType  # noqa: E402

base, mid, top = Node(0, 0, 0), Node(0, 6, 0), Node(0, 12, 0)
base.nodal_support = NodalSupport()
anchors = [Node(-8, 0, 0), Node(8, 0, 0), Node(0, 0, 9)]
for nd_ in anchors:
    nd_.nodal_support = NodalSupport()
hinge = MemberHinge(hinge_type="cap", rotational_release_mz=5.0e5, max_moment_mz=2.0e4, max_tension_vx=1.0e5, rotational_release_warp=0.0)
mast = [Member(start_node=base, end_node=mid, section=sec), Member(start_node=mid, end_node=top, section=sec, start_hinge=hinge)]
guys = [Member(start_node=anchors[0], end_node=top, section=sec, member_type=MemberType.CABLE, pretension=3000.0),
        Member(start_node=anchors[1], end_node=top, section=sec, member_type=MemberType.CABLE, unstretched_length=14.44),
        Member(start_node=anchors[2], end_node=top, section=sec, member_type=MemberType.CABLE)]
small = FERS()
small.add_member_set(MemberSet(members=mast, classification="Mast",
                               buckling_restraints=[BucklingRestraint(node_id=mid.id, restrains_local_y=True, restrains_torsion=True)]))
small.add_member_set(MemberSet(members=guys, classification="Guys"))
lc2 = small.create_load_case(name="W")
NodalLoad(node=top, load_case=lc2, magnitude=800.0, direction=(1, 0, 0))
DistributedLoad(member=guys[1], load_case=lc2, magnitude=20.0, direction=(0, -1, 0))
small.create_load_combination(name="ULS", load_cases_factors={lc2: 1.5}, situation="ULS", check="ALL")
got = wire.from_fers(small)
want = wire.parse_document(json.loads(json.dumps(small.to_dict(include_results=False))))
same(got, want, "guyed mast")
chord0 = (8.0 ** 2 + 12.0 ** 2) ** 0.5
assert got.c_member.tolist() == [2, 3, 4] and got.c_l0[1] == 14.44 and got.c_l0[2] == 15.0
assert abs(got.c_l0[0] - chord0 / (1.0 + 3000.0 / got.c_ea[0])) < 1e-12 and got.c_l0[0] < chord0
assert len(got.hinge_caps) == 2 and got.member_sets_raw[0]["buckling_restraints"][0]["restrains_torsion"] is True
assert got.load_cases[0].cab_q[1].tolist() == [0.0, -20.0, 0.0] and len(got.load_cases[0].dl_member) == 0
checked += 1

# the server-side mesher of unmeshed plate surfaces against the reference's own client-side mesher (plates/mesh.py)
from fers_core.plates.mesh import triangulate_surface_polygon  # noqa: E402

for poly in ([(0, 0, 0), (4, 0, 0), (4, 0, 3), (0, 0, 3)],                                  # rectangle in the XZ plane
             [(0, 0, 0), (4, 0, 0), (4, 2, 0), (2, 2, 0), (2, 4, 0), (0, 4, 0)],            # L shape
             [(0, 0, 0), (0, 4, 0), (2, 4, 0), (2, 2, 0), (4, 2, 0), (4, 0, 0)],            # the same, clockwise
             [(0, 0, 0), (3, 0, 1), (4, 2, 2), (2, 5, 3), (-1, 3, 1)]):                     # convex pentagon in a tilted plane
    pts = np.array(poly, float)
    u, v = pts[1] - pts[0], pts[2] - pts[0]
    n = np.cross(u, v); n /= np.linalg.norm(n)
    pts = pts - np.outer((pts - pts[0]) @ n, n)                                             # exactly planar
    for size in (None, 0.9, 1.5, 2.0, 10.0):                                                # mesh.target_size: refined boundary
        want = triangulate_surface_polygon([tuple(p) for p in pts], mesh_size=size)
        dense = wire.densify_polygon(pts, size)
        got = wire.ear_clip_polygon(dense)
        assert len(got) == len(want) == len(dense) - 2
        for (i, j, k), tri in zip(got, want):
            assert np.allclose(dense[[i, j, k]], np.array(tri), atol=1e-12), (poly, size, got, want)
print(json.dumps({"checked": checked, "members": len(mem), "from_fers_s": t_fast, "to_dict_s": t_dict, "json_round_trip_s": t_json,
                  "parse_document_s": t_parse, "pydantic_gate_s": t_gate, "mesher_polygons": 4}))
