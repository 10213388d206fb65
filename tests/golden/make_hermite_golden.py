"""Generate tests/golden/hermite_cases.json with the reference's OWN client-side Hermite reconstruction.

Run in the build container only (needs /root/reference):

    PYTHONPATH=tests/refshim:/root/reference python tests/golden/make_hermite_golden.py

``MemberResult.member_displacements`` (fers_core/types/pydantic_models.py:1622-1625) is documented as "matching
client-side Hermite reconstructions"; that reconstruction is in the repository:
``centerline_path_points`` / ``interpolate_beam_local`` (fers_core/fers/deformation_utils.py:19-114) over
``Member.local_coordinate_system`` (fers_core/members/member.py:309-354).  For seeded member geometries, roll angles
and end displacements / rotations this script records the wire document of the members (``FERS.to_dict``) together
with ``deformed - original`` of the reference's centerline at scale 1 -- the global displacement polyline our
``hostlogic.member_deflected_shape`` has to reproduce.
"""
import json
import os

import numpy as np

from fers_core import FERS, Member, MemberSet, Material, NodalSupport, Node, Section
from fers_core.fers.deformation_utils import centerline_path_points

HERE = os.path.dirname(os.path.abspath(__file__))
STATIONS = 11

rng = np.random.default_rng(104)
mat = Material(name="Steel", e_mod=210e9, g_mod=80.769e9, density=7850.0, yield_stress=235e6)
sec = Section(name="IPE 180", material=mat, i_y=1.01e-6, i_z=13.21e-6, j=0.027e-6, area=0.00196)

calc = FERS()
ends = [((0, 0, 0), (5, 0, 0), 0.0),            # along +X
        ((0, 0, 0), (0, 3.5, 0), 0.0),          # vertical: reference axis switches to +Z (member.py:320-330)
        ((0, 3.5, 0), (0, 0, 0), 30.0),         # vertical, downwards, rolled
        ((1, 2, 3), (4, 6, 3), 0.0),
        ((1, 2, 3), (-2, 2.5, 7), 17.0),
        ((0, 0, 0), (0, 0, 4), -45.0)]
for _ in range(6):
    a = rng.uniform(-5, 5, 3); b = a + rng.uniform(-4, 4, 3)
    ends.append((tuple(a), tuple(b), float(rng.uniform(-180, 180))))
members = []
for a, b, roll in ends:
    n1, n2 = Node(*[float(v) for v in a]), Node(*[float(v) for v in b])
    members.append(Member(start_node=n1, end_node=n2, section=sec, rotation_angle=roll))
members[0].start_node.nodal_support = NodalSupport()
calc.add_member_set(MemberSet(members=members))
calc.create_load_case(name="none")

doc = calc.to_dict()
doc["results"] = None
node_ids = [n.id for m in members for n in (m.start_node, m.end_node)]
u = {int(i): [float(v) for v in np.concatenate([rng.normal(0, 0.02, 3), rng.normal(0, 0.01, 3)])] for i in sorted(set(node_ids))}

shapes = {}
for m in members:
    d0, d1 = np.array(u[m.start_node.id]), np.array(u[m.end_node.id])
    orig, deformed = centerline_path_points(m, d0[:3], d0[3:], d1[:3], d1[3:], STATIONS, 1.0)
    shapes[str(m.id)] = (deformed - orig).tolist()

with open(os.path.join(HERE, "hermite_cases.json"), "w") as fh:
    json.dump({"doc": doc, "node_displacements": {str(k): v for k, v in u.items()}, "stations": STATIONS,
               "member_displacements": shapes}, fh)
print("wrote", len(shapes), "members")
