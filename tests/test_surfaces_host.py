"""Plate surfaces that arrive without elements are meshed on load (ear clipping, the reference's client-side rule), and
surface loads on models without plates travel one way to the members along ``distribution_direction``."""
import copy

import numpy as np
import pytest

from fers_core_b200 import wire
from oracle import oracle_engine
from tests import helpers


def _tri_area(p):
    return 0.5 * np.linalg.norm(np.cross(p[1] - p[0], p[2] - p[0]))


def test_ear_clipping_covers_a_concave_polygon_once():
    pts = np.array([(0, 0, 0), (4, 0, 0), (4, 2, 0), (2, 2, 0), (2, 4, 0), (0, 4, 0)], float)
    for poly in (pts, pts[::-1]):
        tris = wire.ear_clip_polygon(poly)
        assert len(tris) == 4
        assert sum(_tri_area(poly[list(t)]) for t in tris) == pytest.approx(12.0, rel=1e-12)
        # every triangle turns the same way (counter-clockwise about the normal the first corner defines, as in the reference)
        signs = {np.sign(np.cross(poly[t[1]] - poly[t[0]], poly[t[2]] - poly[t[0]])[2]) for t in tris}
        assert len(signs) == 1
    with pytest.raises(wire.WireError):
        wire.ear_clip_polygon(np.array([(0, 0, 0), (1, 0, 0), (2, 0, 0)], float))
    with pytest.raises(wire.WireError):
        wire.ear_clip_polygon(np.array([(0, 0, 0), (1, 0, 0), (1, 1, 0.1), (0, 1, 0)], float))


def test_unmeshed_plate_surface_is_meshed_on_load():
    doc = copy.deepcopy(helpers.golden("plate_strip_surface_load")["doc"])
    m0 = wire.parse_document(doc)
    mat = doc["model"]["plate_elements"][0]["material"]
    th = doc["model"]["plate_elements"][0]["thickness"]
    # replace the elements by ONE surface over the outline of the strip (nodes of the two long edges, counter-clockwise)
    nodes = doc["model"]["nodes"]
    lower = sorted([n for n in nodes if n["Y"] == 0.0], key=lambda n: n["X"])
    upper = sorted([n for n in nodes if n["Y"] != 0.0], key=lambda n: -n["X"])
    doc["model"]["plate_elements"] = []
    doc["model"]["plate_surfaces"] = [{"id": 9, "boundary_node_ids": [n["id"] for n in lower + upper], "material": mat,
                                       "thickness": th, "behavior": "Shell", "mesh": {"target_size": 0.05}}]
    m = wire.parse_document(doc)
    assert m.n_plates == len(lower) + len(upper) - 2 and set(m.p_ids) == set(range(1, m.n_plates + 1))
    area = lambda mm: sum(_tri_area(mm.xyz[t]) for t in mm.p_nodes)      # noqa: E731
    assert area(m) == pytest.approx(area(m0), rel=1e-12)
    assert any(c == "W_SURFACE_MESH_BOUNDARY_ONLY" for c, _ in m.warnings)
    # the surface load of the case finds the generated elements: same total reaction as with the shipped mesh
    out = oracle_engine.solve_document(doc)["results"]
    helpers.check_expectations(out, helpers.golden("plate_strip_surface_load")["expect"])


def _surface_doc(polygon, dd=(1.0, 0.0, 0.0), q=5.0):
    """The model of the reference's tests/functionality/test_surface_loads.py:8-37: two members along X at z = 0 and z = 1."""
    doc = copy.deepcopy(helpers.golden("cantilever_end_load")["doc"])
    m = doc["model"]
    m["nodes"] = [{"id": 1, "X": 0, "Y": 0, "Z": 0, "nodal_support": 1}, {"id": 2, "X": 4.0, "Y": 0, "Z": 0, "nodal_support": None},
                  {"id": 3, "X": 0, "Y": 0, "Z": 1.0, "nodal_support": 1}, {"id": 4, "X": 4.0, "Y": 0, "Z": 1.0, "nodal_support": None}]
    mem = m["members"][0]
    m["members"] = [dict(mem, id=1, start_node_id=1, end_node_id=2), dict(mem, id=2, start_node_id=3, end_node_id=4)]
    m["member_sets"][0]["member_ids"] = [1, 2]
    doc["analysis"]["options"]["order"] = "LINEAR"
    lc = doc["analysis"]["load_cases"][0]
    lc["nodal_loads"] = []
    lc["surface_loads"] = [{"id": 1, "polygon": [{"X": x, "Y": y, "Z": z} for x, y, z in polygon], "magnitude": q,
                            "direction": [0.0, -1.0, 0.0], "distribution_direction": list(dd) if dd else None}]
    return doc


def test_surface_load_goes_to_the_members_by_tributary_strips():
    doc = _surface_doc([(0, 0, 0), (4, 0, 0), (4, 0, 1), (0, 0, 1)])
    m = wire.parse_document(doc)
    lc = m.load_cases[0]
    assert not m.warnings and sorted(lc.dl_member.tolist()) == [0, 1]
    # each member takes half of the 1 m wide slab: 5 N/m2 x 0.5 m = 2.5 N/m, downwards, over its whole length
    for row in lc.dl_data:
        assert row.tolist() == [0.0, -2.5, 0.0, 0.0, -2.5, 0.0, 0.0, 1.0]
    out = oracle_engine.solve_document(doc)["results"]["loadcases"][doc["analysis"]["load_cases"][0]["name"]]
    assert sum(r["nodal_forces"]["fy"] for r in out["reaction_nodes"].values()) == pytest.approx(5.0 * 4.0, rel=1e-12)


def test_triangular_polygon_becomes_trapezoid_member_loads():
    # right triangle (0,0)-(4,0)-(0,1) in the XZ plane: area 2; the strip boundary z = 0.5 cuts the hypotenuse at x = 2
    doc = _surface_doc([(0, 0, 0), (4, 0, 0), (0, 0, 1)], q=3.0)
    lc = wire.parse_document(doc).load_cases[0]

    def resultant(member):
        tot = 0.0
        for mem, row in zip(lc.dl_member, lc.dl_data):
            if mem == member:
                tot += 0.5 * (row[1] + row[4]) * (row[7] - row[6]) * 4.0
        return tot
    # member 1 (z = 0) takes the part below z = 0.5: area 2 - 0.5 = 1.5; member 2 (z = 1) the corner above: 0.5 * 2 * 0.5 = 0.5
    assert resultant(0) == pytest.approx(-3.0 * 1.5, rel=1e-9) and resultant(1) == pytest.approx(-3.0 * 0.5, rel=1e-9)
    # without a distribution direction the common direction of the in-plane members is used; a polygon away from every member warns
    same = wire.parse_document(_surface_doc([(0, 0, 0), (4, 0, 0), (0, 0, 1)], dd=None, q=3.0)).load_cases[0]
    assert np.array_equal(same.dl_data, lc.dl_data)
    far = wire.parse_document(_surface_doc([(0, 9, 0), (4, 9, 0), (0, 9, 1)]))
    assert any(c == "W_SURFACE_LOAD_NO_PLATES" for c, _ in far.warnings) and len(far.load_cases[0].dl_member) == 0
