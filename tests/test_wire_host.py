"""Host logic without a GPU: wire-document parsing (fers_core/types/pydantic_models.py contract), right-hand-side
planning and imperfections -- the edge cases the reference's schema and builders imply."""
import copy

import numpy as np
import pytest

from fers_core_b200 import hostlogic, wire
from tests import helpers


def _doc(name="cantilever_end_load"):
    return copy.deepcopy(helpers.golden(name)["doc"])


def test_minimal_document_arrays():
    m = wire.parse_document(_doc())
    assert (m.n_nodes, m.n_members, m.n_plates, m.n_dof) == (2, 1, 0, 12)
    assert m.node_ids.tolist() == [1, 2] and m.m_n1.tolist() == [0] and m.m_n2.tolist() == [1]
    assert m.bc_code.tolist() == [[wire.BC_FIXED] * 6, [wire.BC_FREE] * 6]
    assert m.node_support_id.tolist() == [1, -1]
    E, G = 210e9, 80.769e9
    np.testing.assert_allclose(m.prop[0, :4], [E * 0.00196, E * 1.01e-6, E * 1.321e-5, G * 2.7e-8], rtol=1e-15)
    assert m.prop[0, 4:].tolist() == [0.0, 0.0, 0.0, 0.0]            # no shear areas, no warping constant
    lc = m.load_cases[0]
    assert lc.nod_node.tolist() == [1] and lc.nod_vec.tolist() == [[0.0, -1000.0, 0.0, 0.0, 0.0, 0.0]]
    assert m.n_real == 2 and len(m.w_member) == 0 and m.unity_checks == []


@pytest.mark.parametrize("mutate,msg", [
    (lambda d: d.pop("model"), "model"),
    (lambda d: d["model"]["members"][0].update(start_node_id=99), "unknown node"),
    (lambda d: d["model"]["members"][0].update(section=7), "unknown section"),
    (lambda d: d["model"]["members"][0].update(section=None), "without section"),
    (lambda d: d["model"]["members"][0].update(member_type="Spaghetti"), "member_type"),
    (lambda d: d["model"]["nodes"][0].update(nodal_support=5), "nodal_support"),
    (lambda d: d["model"]["sections"][0].update(material=9), "material"),
    (lambda d: d["model"]["members"][0].update(start_hinge=3), "hinge"),
    (lambda d: d["model"]["nodal_supports"][0]["X"].update(condition_type="Spring", stiffness=None), "stiffness"),
    (lambda d: d["model"]["nodes"][1].update(X=0.0, Y=0.0, Z=0.0), "zero length"),
    (lambda d: d["model"]["nodes"][1].update(X=float("nan")), "non-finite coordinate"),
    (lambda d: d["model"]["materials"][0].update(e_mod=-1.0), "e_mod must be a positive"),
    (lambda d: d["model"]["sections"][0].update(i_z=0.0), "i_z must be a positive"),
    (lambda d: d["analysis"]["load_cases"][0]["nodal_loads"][0].update(magnitude=float("inf")), "non-finite load"),
])
def test_malformed_documents_raise(mutate, msg):
    d = _doc()
    mutate(d)
    with pytest.raises(wire.WireError) as e:
        wire.parse_document(d)
    assert msg.lower() in str(e.value).lower()


def test_unknown_keys_and_enum_spellings_are_tolerated():
    d = _doc()
    d["model"]["nodes"][0]["classification"] = "anything"          # extra keys on the wire (node.py:32)
    d["analysis"]["options"]["solver"] = "newton_raphson"
    d["settings"]["unit_settings"]["weightUnit"] = "kg"
    d["model"]["nodal_supports"][0]["RZ"] = {"condition_type": "free", "stiffness": None}
    d["model"]["nodal_supports"][0]["RY"] = {"condition_type": "SPRING", "stiffness": 12.5}
    m = wire.parse_document(d)
    assert m.bc_code[0].tolist() == [1, 1, 1, 1, wire.BC_SPRING, wire.BC_FREE] and m.bc_k[0, 4] == 12.5


def test_two_dimensional_models_restrain_out_of_plane_dofs():
    d = _doc()
    d["analysis"]["options"]["dimensionality"] = "2D"
    m = wire.parse_document(d)
    assert m.bc_code[1].tolist() == [0, 0, 1, 1, 1, 0]                # Z, RX, RY of the free node (builders.py:108-110)


def test_shear_areas_follow_the_option():
    d = _doc("fixed_fixed_timoshenko")
    on = wire.parse_document(d)
    assert on.prop[0, wire.P_GASY] > 0 and on.prop[0, wire.P_GASZ] > 0
    d["analysis"]["options"]["include_shear_deformation"] = False
    off = wire.parse_document(d)
    assert off.prop[0, wire.P_GASY] == 0 and off.prop[0, wire.P_GASZ] == 0


def test_hinges_rigid_links_and_member_types():
    m = wire.parse_document(_doc("hinge_spring"))
    assert (m.m_rel >= 0).any() and m.rel.shape[1] == wire.REL_STRIDE
    r = m.rel[m.m_rel[m.m_rel >= 0][0]]
    assert (r >= 0).any() and (r == -1).any()                         # released / spring DOFs and rigid ones
    rigid = wire.parse_document(_doc("rigid_link"))
    assert (rigid.m_type == wire.MT_RIGID).any()
    row = hostlogic.rigid_property_row(rigid.prop)
    assert row[wire.P_EA] > rigid.prop[:, wire.P_EA].max() and row[7] == 0.0
    tens = wire.parse_document(_doc("tension_only"))
    assert (tens.m_type == wire.MT_TENSION).any()


def test_combinations_and_rhs_plan():
    m = wire.parse_document(_doc("portal_2d_combinations"))
    nlc, nco = len(m.load_cases), len(m.combo_ids)
    assert nlc >= 2 and nco >= 2 and m.combo_factors.shape == (nco, nlc)
    plan = hostlogic.rhs_plan(m)
    assert plan.kinds == ["Loadcase"] * nlc + ["Loadcombination"] * nco
    assert np.array_equal(plan.factors[:nlc], np.eye(nlc)) and np.array_equal(plan.factors[nlc:], m.combo_factors)
    assert sorted(c for g in plan.groups for c in g.columns) == list(range(plan.n_rhs))
    d = _doc("portal_2d_combinations")
    d["analysis"]["options"]["solve_loadcases"] = False
    assert hostlogic.rhs_plan(wire.parse_document(d)).kinds == ["Loadcombination"] * nco
    d["analysis"]["load_combinations"][0]["load_case_factors"].append({"load_case_id": 999, "factor": 1.0})
    with pytest.raises(wire.WireError):
        wire.parse_document(d)


def test_imperfection_groups_and_sway_coordinates():
    m = wire.parse_document(_doc("portal_pdelta_consistent"))
    plan = hostlogic.rhs_plan(m)
    imp = [g for g in plan.groups if g.imperfection is not None]
    assert imp, "the P-Delta golden carries a sway imperfection"
    xyz = hostlogic.imperfect_coordinates(m, imp[0].imperfection)
    sw = imp[0].imperfection["sway_imperfections"][0]
    moved = np.abs(xyz - m.xyz).sum(axis=1) > 0
    height = m.xyz @ np.asarray(sw.get("height_direction") or (0, 1, 0), float)
    assert moved.any() and not moved[np.isclose(height, 0.0)].any()   # lean grows with height above the reference point
    lean = np.linalg.norm(xyz - m.xyz, axis=1)[moved] / np.abs(height[moved])
    np.testing.assert_allclose(lean, abs(float(sw["magnitude"])), rtol=1e-12)


def test_plate_documents_split_quads_and_resolve_pressures():
    m = wire.parse_document(_doc("plate_strip_surface_load"))
    assert m.n_plates > 0 and m.p_nodes.shape == (m.n_plates, 3)
    lc = m.load_cases[0]
    assert len(lc.pp_plate) > 0 and lc.pp_vec.shape == (len(lc.pp_plate), 3)
    le6 = wire.parse_document(_doc("nafems_LE6_8x8"))
    assert le6.n_plates == 2 * 8 * 8 and (le6.p_flags != 0).all()
    assert le6.p_data.shape == (le6.n_plates, wire.NPDATA) and (le6.p_data[:, 3:6] == 1.0).all()


def test_plate_options_modifiers_plane_strain_projected_pressure():
    d = _doc("plate_strip_surface_load")
    pl = d["model"]["plate_elements"][0]
    pl["stiffness_modifiers"] = {"membrane": 0.5, "bending": 0.25, "shear": None}
    pl["plane_state"] = "PlaneStrain"
    m = wire.parse_document(d)
    rows = np.nonzero(m.p_ids == int(pl["id"]))[0]
    assert all(m.p_data[r, 3:6].tolist() == [0.5, 0.25, 1.0] and m.p_flags[r] & 4 for r in rows)
    others = np.nonzero(m.p_ids != int(pl["id"]))[0]
    assert (m.p_data[others, 3:6] == 1.0).all() and not (m.p_flags[others] & 4).any()
    # projected pressure: force scales with |n . d|; the strip lies in a coordinate plane, so an oblique direction at
    # 60 degrees from the normal halves the load of the same magnitude
    base = wire.parse_document(_doc("plate_strip_surface_load")).load_cases[0]
    d2 = _doc("plate_strip_surface_load")
    lc = d2["analysis"]["load_cases"][0]
    if lc.get("plate_pressures"):
        x3 = m.xyz[m.p_nodes[0]]
        n = np.cross(x3[1] - x3[0], x3[2] - x3[0]); n /= np.linalg.norm(n)
        tang = (x3[1] - x3[0]) / np.linalg.norm(x3[1] - x3[0])
        direction = 0.5 * n + np.sqrt(0.75) * tang
        for pp in lc["plate_pressures"]:
            pp["direction"] = direction.tolist(); pp["projected"] = True
        proj = wire.parse_document(d2).load_cases[0]
        for pp in lc["plate_pressures"]:
            pp["projected"] = False
        full = wire.parse_document(d2).load_cases[0]
        np.testing.assert_allclose(proj.pp_vec, 0.5 * full.pp_vec, rtol=1e-12)
        assert len(base.pp_plate) == len(full.pp_plate)


def test_degenerate_plates_raise():
    d = _doc("plate_strip_surface_load")
    pl = d["model"]["plate_elements"][0]
    pl["node_ids"] = [pl["node_ids"][0], pl["node_ids"][1], pl["node_ids"][0]]
    with pytest.raises(wire.WireError, match="degenerate"):
        wire.parse_document(d)
    d = _doc("plate_strip_surface_load")
    d["model"]["plate_elements"][0]["thickness"] = 0.0
    with pytest.raises(wire.WireError, match="thickness"):
        wire.parse_document(d)


def test_warping_sections_get_auxiliary_nodes():
    d = _doc("nafems_LE5_4el")
    m = wire.parse_document(d)
    assert (m.n_real, m.n_nodes) == (5, 10) and m.w_member.tolist() == [0, 1, 2, 3]
    assert m.warp_node.tolist() == [5, 6, 7, 8, 9] and m.w_aux.tolist() == [[5, 6], [6, 7], [7, 8], [8, 9]]
    assert (m.node_ids[5:] < 0).all() and np.array_equal(m.xyz[5:], m.xyz[:5])
    assert m.bc_code[5].tolist() == [1] * 6                           # warping restrained at the clamped end
    assert all(row.tolist() == [0, 1, 1, 1, 1, 1] for row in m.bc_code[6:])
    assert m.prop[0, 7] == pytest.approx(210e9 * 0.0416667)
    d["analysis"]["options"]["include_warping"] = False
    off = wire.parse_document(d)
    assert off.n_nodes == 5 and len(off.w_member) == 0 and off.prop[0, 7] == 0.0
    d["analysis"]["options"]["include_warping"] = True
    d["model"]["members"][1]["member_type"] = "Truss"                 # axial-only members carry no torsion
    mixed = wire.parse_document(d)
    assert mixed.w_member.tolist() == [0, 2, 3]


def test_unity_check_inputs_travel_with_the_model():
    d = _doc("portal_2d_combinations")
    d["analysis"]["unity_checks"] = [{"id": "c", "name": "c", "applies_to": {"type": "AllMembers"},
                                      "spec": {"Generic": {"demand": "1", "capacity": "2", "variables": []}}}]
    d["analysis"]["load_combinations"][0]["limit_state"] = "ULS"
    m = wire.parse_document(d)
    assert len(m.unity_checks) == 1 and m.combo_limit_state[0] == "ULS"
    sec, mat = m.member_raw(0)
    assert sec["area"] > 0 and mat["e_mod"] > 0 and len(m.m_classification) == m.n_members


def test_inputs_that_are_read_but_not_acted_on_are_reported():
    """Nothing on the wire is dropped silently: hinge capacities, contradictory principal-axis data and mesh requests that are not honoured each leave a
    warning in ``errors_and_warnings`` (pydantic_models.py:1877-1880)."""
    import copy
    doc = copy.deepcopy(helpers.golden("hinge_spring")["doc"])
    assert not [w for w in wire.parse_document(doc).warnings if w[0].startswith(("W_HINGE", "W_PRINCIPAL", "W_SURFACE_NOT"))]
    doc["model"]["member_hinges"][0]["max_moment_mz"] = 1.0e4
    doc["model"]["sections"][0]["i_yz"] = 1.0e-7
    # (nodes 1, 2, 3 of this model are collinear: a surface over them is rejected, one with an opening is meshed and noted)
    doc["model"]["plate_surfaces"] = [{"id": 7, "boundary_node_ids": [1, 2, 3], "material": 1, "thickness": 0.01}]
    with pytest.raises(wire.WireError, match="degenerate"):
        wire.parse_document(doc)
    doc["model"]["nodes"].append({"id": 99, "X": 1.0, "Y": 1.0, "Z": 0.5, "nodal_support": None})
    doc["model"]["plate_surfaces"] = [{"id": 7, "boundary_node_ids": [1, 2, 99], "material": 1, "thickness": 0.01,
                                       "openings": [{"id": 1, "boundary_node_ids": [1, 2, 99]}]}]
    codes = [w[0] for w in wire.parse_document(doc).warnings]
    assert codes.count("W_HINGE_CAPACITY_IGNORED") == 1
    assert not [c for c in codes if c.startswith("W_PRINCIPAL")]      # applied (tests/test_product_inertia.py), not ignored
    assert codes.count("W_SURFACE_OPENINGS_IGNORED") == 1 and wire.parse_document(doc).n_plates == 1
    # a slack model other than the default on a model with tension-only members
    t_doc = copy.deepcopy(helpers.golden("tension_only")["doc"])
    assert "W_AXIAL_SLACK_IGNORED" not in [w[0] for w in wire.parse_document(t_doc).warnings]
    t_doc["analysis"]["options"]["axial_slack"] = 0.1
    assert "W_AXIAL_SLACK_IGNORED" in [w[0] for w in wire.parse_document(t_doc).warnings]
    # an angle that contradicts the (i_y, i_z, i_yz) triple is resolved in favour of the angle, with a note
    doc["model"]["sections"][0].update(i_yz=1.0e-7, principal_axis_angle=40.0)
    assert "W_PRINCIPAL_AXES" in [w[0] for w in wire.parse_document(doc).warnings]
