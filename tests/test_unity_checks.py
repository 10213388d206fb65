"""Unity checks: expression compiler and check compilation on the CPU, device
evaluation against the oracle restatement (oracle/unity_oracle.py) and against the one numeric statement
the reference makes (example 100: chord-relative sag == global mid-span deflection on rigid supports)."""
import copy
import json
import math

import numpy as np
import pytest

from fers_core_b200 import unity, wire
from oracle import unity_oracle
from tests import helpers


def _run_postfix(code, consts, vals):
    st = []
    for op, arg in code:
        if op == 0: st.append(consts[arg])
        elif op == 1: st.append(vals[arg])
        elif op == 7: st[-1] = -st[-1]
        elif op == 8: st[-1] = abs(st[-1])
        elif op == 9: st[-1] = math.sqrt(st[-1])
        else:
            b = st.pop(); a = st.pop()
            st.append({2: a + b, 3: a - b, 4: a * b, 5: a / b if b else math.inf, 6: a ** b, 10: min(a, b), 11: max(a, b)}[op])
    assert len(st) == 1
    return st[0]


@pytest.mark.parametrize("text", [
    "M * c / I", "abs(d)", "-a ^ 2 + b", "(a + b) * (c - d) / 2.5e-1", "max(a, b, c) - min(d, a)", "sqrt(a*a + b*b) / (c + 1)",
    "a - b - c", "a / b / c", "2 ^ 3 ^ 2", "-(a - -b)", "pow(a, 2) + .5 * c", "y_ref - y_self", "1e3*a+B2_x",
])
def test_expression_compiler_matches_python(text):
    names = {"M": 3.5, "c": 0.11, "I": 2.0e-3, "d": -4.0, "a": 1.7, "b": -2.25, "y_ref": 4.0, "y_self": 3.1, "B2_x": 9.0}
    idx = {n: i for i, n in enumerate(names)}
    consts = []
    code = unity.compile_expression(text, idx, consts)
    got = _run_postfix(code, consts, list(names.values()))
    want = unity_oracle._eval(text, names)
    assert got == pytest.approx(want, rel=1e-15, abs=0.0)


@pytest.mark.parametrize("text", ["a +", "foo(a)", "a b", "(a", "unknown * 2"])
def test_expression_compiler_rejects(text):
    with pytest.raises(unity.UnityError):
        unity.compile_expression(text, {"a": 0, "b": 1}, [])


def _beam_doc():
    """Example 100 (fers_core/examples/100_Unity_Check_Beam_Deflection_SLS.py:30-88): simply supported 6 m beam
    of two half-members, 5 kN at mid-span, SLS deflection check on the member set."""
    doc = copy.deepcopy(helpers.golden("cantilever_end_load")["doc"])
    pin = {k: {"condition_type": "Fixed", "stiffness": None} for k in ("X", "Y", "Z", "RX")}
    pin.update({k: {"condition_type": "Free", "stiffness": None} for k in ("RY", "RZ")})
    m = doc["model"]
    m["nodes"] = [{"id": 1, "X": 0, "Y": 0, "Z": 0, "nodal_support": 1}, {"id": 2, "X": 3.0, "Y": 0, "Z": 0, "nodal_support": None},
                  {"id": 3, "X": 6.0, "Y": 0, "Z": 0, "nodal_support": 1}]
    m["nodal_supports"] = [dict(id=1, classification=None, warping_condition=None, **pin)]
    base = m["members"][0]
    m["members"] = [dict(base, id=1, start_node_id=1, end_node_id=2, classification="Beam"),
                    dict(base, id=2, start_node_id=2, end_node_id=3, classification="Beam")]
    m["member_sets"] = [{"id": 7, "classification": "Beam", "member_ids": [1, 2], "buckling_restraints": []}]
    m["sections"][0].update(h=0.18, b=0.091, wel_z=1.46e-4)
    a = doc["analysis"]
    a["options"]["order"] = "LINEAR"
    a["load_cases"] = [dict(a["load_cases"][0], id=1, name="SLS", nodal_loads=[
        {"id": 1, "node": 2, "magnitude": -5000.0, "direction": [0, 1, 0], "load_type": "force"}])]
    a["load_combinations"] = [
        {"id": 11, "name": "SLS-combo", "check": "ALL", "situation": "SLS", "limit_state": "SLS",
         "load_case_factors": [{"load_case_id": 1, "factor": 1.0}]},
        {"id": 12, "name": "ULS-combo", "check": "ALL", "situation": "ULS", "limit_state": "ULS",
         "load_case_factors": [{"load_case_id": 1, "factor": 1.5}]}]
    q = lambda kind, **kw: {"Quantity": {kind: kw}}
    a["unity_checks"] = [
        {"id": "sls_beam", "name": "Beam mid-span deflection", "applies_to": {"type": "MemberSets", "ids": [7]},
         "limit_state": "SLS", "thresholds": [0.8, 0.95, 1.0],
         "spec": {"Generic": {"demand": "abs(d)", "capacity": "limit", "variables": [
             {"name": "d", "source": q("MemberDeflection", component="Magnitude", aggregation={"type": "MaxAbs"})},
             {"name": "limit", "source": {"Quantity": {"Constant": 6.0 / 250.0}}}]}}},
        {"id": "bending", "name": "Bending stress", "applies_to": {"type": "AllMembers"}, "limit_state": "ULS",
         "spec": {"Generic": {"demand": "M * c / I", "capacity": "fy / 1.0", "variables": [
             {"name": "c", "source": {"Expression": "h / 2"}},
             {"name": "M", "source": q("MemberForce", component="Mz", aggregation={"type": "MaxAbs"})},
             {"name": "h", "source": q("Section", property="H")},
             {"name": "I", "source": q("Section", property="Iz")},
             {"name": "fy", "source": q("Material", property="Fy")}]}}},
        {"id": "shear_end", "name": "End shear", "applies_to": {"type": "Members", "ids": [2]}, "load_combination_ids": [12],
         "spec": {"Generic": {"demand": "max(abs(V0), abs(V1), abs(Vq))", "capacity": "0.58 * fy * A / L ^ 0", "variables": [
             {"name": "V0", "source": q("MemberForce", component="Vy", aggregation={"type": "Start"})},
             {"name": "V1", "source": q("MemberForce", component="Vy", aggregation={"type": "End"})},
             {"name": "Vq", "source": q("MemberForce", component="Vy", aggregation={"type": "AtFraction", "fraction": 0.25})},
             {"name": "A", "source": q("Section", property="Area")},
             {"name": "L", "source": q("Geometry", property="Length")},
             {"name": "fy", "source": q("Material", property="Fy")}]}}},
        {"id": "clear", "name": "Deflected position", "applies_to": {"type": "Members", "ids": [1]},
         "spec": {"Generic": {"demand": "0.0 - y_end", "capacity": "0.05 + dx0", "variables": [
             {"name": "y_end", "source": q("NodePosition", at="End", axis="Y")},
             {"name": "dx0", "source": q("NodeDisplacement", at="Start", component="Magnitude")}]}}},
        {"id": "ec3", "name": "EC3", "applies_to": {"type": "AllMembers"}, "spec": {"Ec3Steel": {"gamma_m0": 1.0}}},
    ]
    return doc


def test_check_compilation_tables():
    model = wire.parse_document(_beam_doc())
    cs = [unity.compile_check(model, d) for d in model.unity_checks]
    sls, bend, shear, clear, ec3 = cs
    assert sls.entities == [{"type": "MemberSet", "memberset_id": 7}]
    assert sls.ent_members.tolist() == [0, 1] and sls.ent_nodes.tolist() == [0, 1, 2]
    assert sls.combo_ids == {11} and bend.combo_ids == {12} and shear.combo_ids == {12} and clear.combo_ids is None
    assert len(bend.entities) == 2 and bend.var_names[-1] == "c"          # the expression moves after what it names
    assert bend.var_desc[bend.var_names.index("M")].tolist() == [unity.V_FORCE, 5, 0, 0]
    assert shear.var_desc[shear.var_names.index("Vq")].tolist() == [unity.V_FORCE, 1, 5, 1]
    # Ec3Steel: compiled into the same tables from the generated program (fers_core_b200/ec3.py); this section has no wel / wpl
    assert ec3.skipped is None and ec3.labels is not None and len(ec3.entities) == 2
    assert any("no wel / wpl" in w for w in ec3.warnings) and len(ec3.var_names) <= unity.MAX_VARS
    assert unity._status(0.5, None) == "Green" and unity._status(0.9, None) == "Yellow"
    assert unity._status(0.97, None) == "Orange" and unity._status(1.2, None) == "Red"


def test_fold_envelope_rules():
    best = np.array([[-np.inf, 0, 0, -1], [0.5, 1, 2, 3], [0.5, 1, 2, 3], [np.nan, 0, 0, 2]], float)
    got = np.array([[0.1, 1, 10, 4], [0.5, 1, 2, 1], [0.4, 9, 9, 0], [0.2, 1, 5, 7]], float)
    unity.fold_envelope(best, got)
    assert best[:, 0].tolist() == [0.1, 0.5, 0.5, 0.2] and best[:, 3].tolist() == [4, 1, 3, 7]


@pytest.mark.gpu
def test_unity_checks_match_oracle_and_example_100():
    import fers_calculations
    from fers_core_b200 import engine
    from oracle import fers_oracle
    doc = _beam_doc()
    out = json.loads(fers_calculations.calculate_from_json(json.dumps(doc)))["results"]
    ucs = {u["check_id"]: u for u in out["unity_check_results"]}
    assert set(ucs) == {"sls_beam", "bending", "shear_end", "clear", "ec3"}
    # example 100: chord-relative sag == global mid-span |Dy| on rigid supports == P L^3 / (48 E Iz)
    gov = ucs["sls_beam"]["governing"]
    dy_mid = abs(out["loadcombinations"]["SLS-combo"]["displacement_nodes"]["2"]["dy"])
    assert abs(gov["demand"] - dy_mid) < 1e-9 * dy_mid
    assert gov["demand"] == pytest.approx(5000.0 * 6.0 ** 3 / (48 * 210e9 * 1.321e-5), rel=1e-9)
    assert gov["governing_combination_id"] == 11 and gov["entity"] == {"type": "MemberSet", "memberset_id": 7}
    assert ucs["sls_beam"]["status"] == unity._status(gov["demand"] / 0.024, None)
    # bending: M = 1.5 * P L / 4 at mid-span (ULS only), sigma = M (h/2) / Iz
    b = ucs["bending"]["governing"]
    assert b["demand"] == pytest.approx(1.5 * 5000 * 6 / 4 * 0.09 / 1.321e-5, rel=1e-9) and b["capacity"] == 235e6
    assert b["governing_combination_id"] == 12
    assert len(ucs["ec3"]["per_entity"]) == 2 and ucs["ec3"]["governing"]["trace"][-1]["label"] == "Governing"
    # every check against the oracle restatement on the oracle's own results
    model = wire.parse_document(doc)
    want = {w["check_id"]: w for w in unity_oracle.evaluate(
        _generic_only(model), fers_oracle.analyze(model))}
    for cid, w in want.items():
        for pe, pw in zip(ucs[cid]["per_entity"], w["per_entity"]):
            assert pe["entity"] == pw["entity"] and pe["governing_combination_id"] == pw["governing_combination_id"]
            for k in ("utilization", "demand", "capacity"):
                assert pe[k] == pytest.approx(pw[k], rel=1e-9, abs=1e-300), (cid, k)


def _clearance_doc():
    """Example 101 (fers_core/examples/101_Unity_Check_Reference_Member_Clearance.py:44-108): a loaded rail cantilever
    that references the beam above it; one check composes quantities of both members."""
    doc = copy.deepcopy(helpers.golden("cantilever_end_load")["doc"])
    m = doc["model"]
    m["nodes"] = [{"id": 1, "X": 0, "Y": 0.5, "Z": 0, "nodal_support": 1}, {"id": 2, "X": 4.0, "Y": 0.5, "Z": 0, "nodal_support": None},
                  {"id": 3, "X": 0, "Y": 0.0, "Z": 0, "nodal_support": 1}, {"id": 4, "X": 4.0, "Y": 0.0, "Z": 0, "nodal_support": None}]
    base = m["members"][0]
    m["members"] = [dict(base, id=1, start_node_id=1, end_node_id=2, classification="Beam"),
                    dict(base, id=2, start_node_id=3, end_node_id=4, classification="Rail", reference_member=1, reference_node=2)]
    m["member_sets"] = [{"id": 1, "classification": "Beam", "member_ids": [1], "buckling_restraints": []},
                        {"id": 2, "classification": "Rail", "member_ids": [2], "buckling_restraints": []}]
    a = doc["analysis"]
    a["options"]["order"] = "LINEAR"
    a["load_cases"] = [dict(a["load_cases"][0], id=1, name="SLS", nodal_loads=[
        {"id": 1, "node": 4, "magnitude": -3000.0, "direction": [0, 1, 0], "load_type": "force"},
        {"id": 2, "node": 2, "magnitude": -100.0, "direction": [0, 1, 0], "load_type": "force"}])]
    a["load_combinations"] = [{"id": 5, "name": "SLS-combo", "check": "ALL", "situation": "SLS",
                               "load_case_factors": [{"load_case_id": 1, "factor": 1.0}]}]
    q = lambda kind, **kw: {"Quantity": {kind: kw}}
    a["unity_checks"] = [{
        "id": "rail_clearance", "name": "Rail-to-beam vertical clearance", "applies_to": {"type": "MemberClassification", "value": "Rail"},
        "spec": {"Generic": {"demand": "gap_min", "capacity": "y_ref - y_self", "variables": [
            {"name": "y_self", "source": q("NodePosition", at="End", axis="Y")},
            {"name": "y_ref", "source": q("NodePosition", at="ReferenceMemberEnd", axis="Y")},
            {"name": "d_self", "source": q("NodeDisplacement", at="End", component="Dy")},
            {"name": "d_ref", "source": q("NodeDisplacement", at="ReferenceMemberEnd", component="Dy")},
            {"name": "d_node", "source": q("NodeDisplacement", at="ReferenceNode", component="Dy")},
            {"name": "Lb", "source": q("Geometry", property="Length", of="ReferenceMember")},
            {"name": "Vb", "source": q("MemberForce", component="Vy", aggregation={"type": "MaxAbs"}, of="ReferenceMember")},
            {"name": "Ab", "source": q("Section", property="Area", of="ReferenceMember")},
            {"name": "dmax", "source": q("Displacement", component="Magnitude")},
            {"name": "gap_min", "source": {"Quantity": {"Constant": 0.40}}}]}}},
        {"id": "no_ref", "name": "members without a reference", "applies_to": {"type": "Members", "ids": [1]},
         "spec": {"Generic": {"demand": "abs(d)", "capacity": "1", "variables": [
             {"name": "d", "source": q("NodeDisplacement", at="ReferenceMemberEnd", component="Dy")}]}}}]
    return doc


@pytest.mark.gpu
def test_reference_member_quantities_and_trace():
    import fers_calculations
    from oracle import fers_oracle
    doc = _clearance_doc()
    out = json.loads(fers_calculations.calculate_from_json(json.dumps(doc)))["results"]
    ucs = {u["check_id"]: u for u in out["unity_check_results"]}
    gov = ucs["rail_clearance"]["governing"]
    assert gov["entity"] == {"type": "Member", "member_id": 2}
    tr = {s["label"]: s["value"] for s in gov["trace"]}
    EI = 210e9 * 1.321e-5
    assert tr["d_self"] == pytest.approx(-3000.0 * 4 ** 3 / (3 * EI), rel=1e-9)        # rail tip: P L^3 / 3 E I
    assert tr["d_ref"] == pytest.approx(-100.0 * 4 ** 3 / (3 * EI), rel=1e-9) and tr["d_node"] == tr["d_ref"]
    assert tr["y_self"] == pytest.approx(0.0 + tr["d_self"], rel=1e-12) and tr["y_ref"] == pytest.approx(0.5 + tr["d_ref"], rel=1e-12)
    assert tr["Lb"] == 4.0 and tr["Vb"] == pytest.approx(100.0, rel=1e-9) and tr["Ab"] == 0.00196
    assert tr["dmax"] == pytest.approx(abs(tr["d_self"]), rel=1e-9)
    assert gov["utilization"] == pytest.approx(0.40 / (tr["y_ref"] - tr["y_self"]), rel=1e-12)
    assert tr["Utilization"] == gov["utilization"]
    lone = ucs["no_ref"]["per_entity"][0]
    assert lone["status"] == "Red" and not math.isfinite(lone["utilization"])           # no reference member: NaN -> Red
    model = wire.parse_document(doc)
    want = unity_oracle.evaluate(model, fers_oracle.analyze(model))[0]
    pe, pw = ucs["rail_clearance"]["per_entity"][0], want["per_entity"][0]
    for k in ("utilization", "demand", "capacity"):
        assert pe[k] == pytest.approx(pw[k], rel=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["plate_strip_surface_load", "nafems_LE6_8x8"])
def test_plate_stress_checks_match_oracle(name):
    """PlateStress quantities on plate entities (all plates / listed plate elements) against the oracle."""
    from fers_core_b200 import engine
    from oracle import fers_oracle
    doc = copy.deepcopy(helpers.golden(name)["doc"])
    a = doc["analysis"]
    lc_id = a["load_cases"][0]["id"]
    a["load_combinations"] = [{"id": 3, "name": "C1", "check": "ALL", "load_case_factors": [{"load_case_id": lc_id, "factor": 1.0}]},
                              {"id": 4, "name": "C2", "check": "ALL", "load_case_factors": [{"load_case_id": lc_id, "factor": -1.5}]}]
    q = lambda kind, **kw: {"Quantity": {kind: kw}}
    first = int(doc["model"]["plate_elements"][0]["id"])
    a["unity_checks"] = [
        {"id": "vm", "name": "von Mises", "applies_to": {"type": "AllPlates"},
         "spec": {"Generic": {"demand": "s", "capacity": "fy / 1.1", "variables": [
             {"name": "s", "source": q("PlateStress", measure="VonMises")}, {"name": "fy", "source": {"Quantity": {"Constant": 235e6}}}]}}},
        {"id": "principal", "name": "principal range", "applies_to": {"type": "PlateElements", "ids": [first]},
         "spec": {"Generic": {"demand": "max(abs(s1), abs(s2)) + 0 * d", "capacity": "235e6", "variables": [
             {"name": "s1", "source": q("PlateStress", measure="PrincipalMax")},
             {"name": "s2", "source": q("PlateStress", measure="PrincipalMin")},
             {"name": "d", "source": q("Displacement", component="Magnitude")}]}}}]
    # the elements of the first half of the list belong to surface 7, the rest to surface 8: one entity per surface
    half = len(doc["model"]["plate_elements"]) // 2
    for k, p in enumerate(doc["model"]["plate_elements"]):
        p["source_surface_id"] = 7 if k < half else 8
    a["unity_checks"].append(
        {"id": "surf", "name": "von Mises per surface", "applies_to": {"type": "PlateSurfaces", "ids": [7, 8]},
         "spec": {"Generic": {"demand": "s", "capacity": "235e6", "variables": [{"name": "s", "source": q("PlateStress", measure="VonMises")}]}}})
    model = wire.parse_document(doc)
    got = engine.analyze(wire.parse_document(doc)).unity.results()
    want = unity_oracle.evaluate(model, fers_oracle.analyze(model))
    assert len(got[0]["per_entity"]) == len({int(i) for i in model.p_ids}) and len(got[1]["per_entity"]) == 1
    assert [pe["entity"] for pe in got[2]["per_entity"]] == [{"type": "PlateSurface", "plate_surface_id": 7},
                                                             {"type": "PlateSurface", "plate_surface_id": 8}]
    # a surface is as utilised as its worst element
    by_el = {pe["entity"]["plate_element_id"]: pe["utilization"] * (235e6 / 1.1) / 235e6 for pe in got[0]["per_entity"]}
    ids7 = [int(p["id"]) for p in doc["model"]["plate_elements"][:half]]
    assert got[2]["per_entity"][0]["utilization"] == pytest.approx(max(by_el[i] for i in ids7), rel=1e-9)
    for g, w in zip(got, want):
        assert g["max_utilization"] > 0
        for pe, pw in zip(g["per_entity"], w["per_entity"]):
            assert pe["entity"] == pw["entity"] and pe["governing_combination_id"] == pw["governing_combination_id"]
            assert pe["utilization"] == pytest.approx(pw["utilization"], rel=1e-8)


def test_plate_surface_selector_on_the_oracle():
    """EntitySelector PlateSurfaces (pydantic_models.py:395-400): one entity per surface = the elements meshed from it."""
    from oracle import fers_oracle
    doc = copy.deepcopy(helpers.golden("plate_strip_surface_load")["doc"])
    els = doc["model"]["plate_elements"]
    for k, p in enumerate(els):
        p["source_surface_id"] = 7 if k < len(els) // 2 else 8
    q = {"Quantity": {"PlateStress": {"measure": "VonMises"}}}
    doc["analysis"]["unity_checks"] = [
        {"id": "el", "name": "per element", "applies_to": {"type": "AllPlates"},
         "spec": {"Generic": {"demand": "s", "capacity": "235e6", "variables": [{"name": "s", "source": q}]}}},
        {"id": "surf", "name": "per surface", "applies_to": {"type": "PlateSurfaces", "ids": [8, 7]},
         "spec": {"Generic": {"demand": "s", "capacity": "235e6", "variables": [{"name": "s", "source": q}]}}}]
    model = wire.parse_document(doc)
    assert sorted(set(model.p_surface.tolist())) == [7, 8]
    out = unity_oracle.evaluate(model, fers_oracle.analyze(model))
    by_el = {pe["entity"]["plate_element_id"]: pe["utilization"] for pe in out[0]["per_entity"]}
    assert [pe["entity"]["plate_surface_id"] for pe in out[1]["per_entity"]] == [8, 7]
    for pe in out[1]["per_entity"]:
        mine = [int(p["id"]) for p in els if p["source_surface_id"] == pe["entity"]["plate_surface_id"]]
        assert pe["utilization"] == pytest.approx(max(by_el[i] for i in mine), rel=1e-12)
    # the product compiles the same entity table
    c = unity.compile_check(model, doc["analysis"]["unity_checks"][1])
    assert c.plates and [e["plate_surface_id"] for e in c.entities] == [8, 7]
    rows = {sid: sorted(np.nonzero(model.p_surface == sid)[0].tolist()) for sid in (7, 8)}
    assert sorted(c.ent_members[c.ent_ptr[0]:c.ent_ptr[1]].tolist()) == rows[8]
    bad = unity.compile_check(model, dict(doc["analysis"]["unity_checks"][1], applies_to={"type": "PlateSurfaces", "ids": [99]}))
    assert bad.skipped and "plate surface 99" in bad.skipped


def _generic_only(model):
    return model          # the oracle evaluates Ec3Steel checks as well (oracle/ec3_oracle.py)


@pytest.mark.gpu
def test_unity_envelope_over_many_combinations_on_frame():
    """72 combinations (two batches of device columns) on a 3D frame: the governing combination and
    utilisation of every member equal the oracle's."""
    from fers_core_b200 import engine, synth
    from oracle import fers_oracle
    model = synth.grid_frame(4, 3, 3, n_combos=72, solve_loadcases=True)
    q = lambda kind, **kw: {"Quantity": {kind: kw}}
    model.unity_checks = [{
        "id": "axial_bending", "name": "N/A + M/W", "applies_to": {"type": "AllMembers"},
        "spec": {"Generic": {"demand": "abs(N) / A + max(abs(My), abs(Mz)) / W", "capacity": "fy", "variables": [
            {"name": "N", "source": q("MemberForce", component="N", aggregation={"type": "MaxAbs"})},
            {"name": "My", "source": q("MemberForce", component="My", aggregation={"type": "Max"})},
            {"name": "Mz", "source": q("MemberForce", component="Mz", aggregation={"type": "Min"})},
            {"name": "A", "source": {"Quantity": {"Constant": 0.0149}}},
            {"name": "W", "source": {"Quantity": {"Constant": 1.3e-3}}},
            {"name": "fy", "source": {"Quantity": {"Constant": 355e6}}}]}}}]
    got = engine.analyze(model, device="cuda:0")
    res = got.unity.results()[0]
    want = unity_oracle.evaluate(model, fers_oracle.analyze(model))[0]
    assert len(res["per_entity"]) == model.n_members
    for pe, pw in zip(res["per_entity"], want["per_entity"]):
        assert pe["governing_combination_id"] == pw["governing_combination_id"]
        assert pe["utilization"] == pytest.approx(pw["utilization"], rel=1e-9)
    assert res["max_utilization"] == pytest.approx(want["max_utilization"], rel=1e-9)
