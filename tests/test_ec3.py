"""EN 1993-1-1 member check (Ec3Steel unity-check spec).

Pins: the reference's own hand calculation (tests/functionality/test_check_beam.py:25-45: IPE300, 6 m, S235, UDL 10 kN/m x
1.35 -> bending z = 60.75 / 147.6 +- 3 %, LTB governs) on the document its check_beam builder emits (golden fixture
``ec3_check_beam_ipe300``); the reference's test file itself runs against the oracle in tests/test_reference_suite.py.
Parity: the compiled program (fers_core_b200/ec3.py -> postfix code for csrc/unity.cu) against the rules written out directly
in oracle/ec3_oracle.py -- algebra on the host here, the kernel on the GPU below."""
import copy
import json
import math

import numpy as np
import pytest

from fers_core_b200 import ec3, unity, wire
from oracle import ec3_oracle, oracle_engine, unity_oracle
from tests import helpers


def _rows(result):
    return {t["label"]: t["value"] for t in result["governing"]["trace"]}


def test_oracle_reproduces_the_reference_hand_calculation():
    out = oracle_engine.solve_document(helpers.golden("ec3_check_beam_ipe300")["doc"])["results"]
    u = out["unity_check_results"][0]
    rows = _rows(u)
    assert u["check_id"] == "ec3"
    assert rows["Bending z (6.2.5)"] == pytest.approx(60.75 / 147.6, rel=0.03)
    assert rows["Bending z (6.2.5)"] == pytest.approx(60750.0 / (6.285287108417889e-4 * 235e6), rel=1e-9)
    assert rows["LTB (6.3.2)"] > rows["Bending z (6.2.5)"] > 0 and rows["Governing"] == pytest.approx(rows["LTB (6.3.2)"], rel=1e-12)
    # C1 of a uniformly loaded simply supported span by the quarter-point formula: 12.5 / (2.5 + 3*0.75 + 4 + 3*0.75) = 1.136
    assert rows["C1 (6.3.2)"] == pytest.approx(12.5 / 11.0, rel=1e-9)
    # M_cr of the doubly symmetric I section, closed form with the catalogue properties
    E, G, Iy, Iw, J, L = 210e9, 80.769e9, 6.037916374311326e-06, 1.2424924816685492e-07, 1.9920391680626774e-07, 6.0
    mcr = rows["C1 (6.3.2)"] * math.pi ** 2 * E * Iy / L ** 2 * math.sqrt(Iw / Iy + L * L * G * J / (math.pi ** 2 * E * Iy))
    assert rows["M_cr"] == pytest.approx(mcr, rel=1e-3)          # g_mod of the builder's S235 differs in the 4th digit at most


def _random_case(seed):
    rng = np.random.default_rng(seed)
    c = dict(fy=235e6 * rng.uniform(1, 2), E=210e9, G=80.77e9, A=rng.uniform(2e-3, 2e-2), Iy=rng.uniform(1e-6, 1e-5),
             Iz=rng.uniform(2e-5, 2e-4), J=rng.uniform(1e-8, 1e-6), Iw=rng.uniform(0, 1e-6), L=rng.uniform(2, 9),
             Wy=rng.uniform(5e-5, 5e-4), Wz=rng.uniform(2e-4, 2e-3), Avy=rng.uniform(1e-3, 5e-3), Avz=ec3.BIG, ay=0.34, az=0.49,
             alt=0.21, gm0=1.0, gm1=1.1, c1=float(rng.choice([0.0, 1.0, 1.35])), on_b=1.0, on_lt=1.0,
             zz=float(rng.choice([0.0, 0.08, -0.05])))        # C2 z_g - C3 z_j of the general M_cr
    c.update(Lcry=c["L"] * rng.uniform(0.5, 1), Lcrz=c["L"], Llt=c["L"] * rng.uniform(0.5, 1))
    m = rng.uniform(1e3, 2e5)
    f = dict(Nmax=rng.uniform(-1e5, 1e5), Nmin=rng.uniform(-3e5, 0), Vy=rng.uniform(0, 1e5), Vz=rng.uniform(0, 1e4),
             My=rng.uniform(0, 2e4), Mz=m, Ma=m * rng.uniform(-1, 1), Mb=m * rng.uniform(-1, 1), Mc=m * rng.uniform(-1, 1))
    return c, f


@pytest.mark.parametrize("seed", range(6))
def test_program_expressions_equal_the_rules_written_out(seed):
    """Every expression of the generated program, evaluated in binding order by the oracle's expression evaluator (Python's own
    parser), against oracle/ec3_oracle.rules; and through the product's postfix compiler."""
    c, f = _random_case(seed)
    want = ec3_oracle.rules(c, f)
    g = ec3.program(ec3.CONST_NAMES)
    env = {**c, **f}
    names = list(env)
    for v in g["variables"]:
        if "Expression" in v["source"]:
            text = v["source"]["Expression"]
            val = unity_oracle._eval(text, env)
            consts, idx = [], {n: i for i, n in enumerate(names)}
            code = unity.compile_expression(text, idx, consts)
            st = []
            for op, arg in code:                       # the postfix machine of csrc/unity.cu on the host
                if op == 0: st.append(consts[arg])
                elif op == 1: st.append(env[names[arg]])
                elif op == 7: st.append(-st.pop())
                elif op == 8: st.append(abs(st.pop()))
                elif op == 9: st.append(math.sqrt(st.pop()))
                else:
                    b, a = st.pop(), st.pop()
                    st.append({2: lambda: a + b, 3: lambda: a - b, 4: lambda: a * b, 5: lambda: a / b, 6: lambda: a ** b,
                               10: lambda: min(a, b), 11: lambda: max(a, b)}[op]())
            assert st[0] == pytest.approx(val, rel=1e-14, abs=0.0), v["name"]
            env[v["name"]] = val
            names.append(v["name"])
    for k, w in want.items():
        assert env[k] == pytest.approx(w, rel=1e-12, abs=1e-300), k
    assert unity_oracle._eval(g["demand"], env) / unity_oracle._eval(g["capacity"], env) == pytest.approx(want["gov"], rel=1e-12)


def test_defaults_follow_the_catalogue_note():
    """No Ec3SectionParams on the section: class from wpl presence (plastic modulus used), buckling curves b."""
    model = wire.parse_document(helpers.golden("ec3_check_beam_ipe300")["doc"])
    cst, notes = ec3.entity_constants(model, {}, [0])
    sec = model.member_raw(0)[0]
    assert cst["Wz"] == sec["wpl_z"] and cst["Wy"] == sec["wpl_y"] and (cst["ay"], cst["az"], cst["alt"]) == (0.34, 0.34, 0.34)
    assert cst["on_lt"] == 1.0 and cst["L"] == 6.0 and not notes
    want = ec3_oracle.constants(model, {}, [0])
    assert {k: cst[k] for k in want} == want
    doc = copy.deepcopy(helpers.golden("ec3_check_beam_ipe300")["doc"])
    doc["model"]["sections"][0]["ec3"] = {"section_class": 3, "buckling_curve_lt": "A", "buckling_curve_y": "C"}
    doc["model"]["member_sets"][0].update(buckling_length_y=3.0, ltb_length=2.0)
    m2 = wire.parse_document(doc)
    c2, _ = ec3.entity_constants(m2, {"gamma_m1": 1.1, "c1": 1.0}, [0])
    assert c2["Wz"] == sec["wel_z"] and (c2["alt"], c2["ay"], c2["az"]) == (0.21, 0.49, 0.34)
    assert (c2["Lcry"], c2["Lcrz"], c2["Llt"], c2["gm1"], c2["c1"]) == (3.0, 6.0, 2.0, 1.1, 1.0)
    assert {k: c2[k] for k in want} == ec3_oracle.constants(m2, {"gamma_m1": 1.1, "c1": 1.0}, [0])


def test_general_mcr_load_height():
    """NCCI SN003: a load on the top flange (z_g > 0, C2 = 0.454 for a uniform load) lowers M_cr, one below the shear centre
    raises it; zero C2 / C3 is the doubly symmetric form; the mono-symmetry term reads Section.wagner_coeff."""
    doc = copy.deepcopy(helpers.golden("ec3_check_beam_ipe300")["doc"])
    model = wire.parse_document(doc)
    base, _ = ec3.entity_constants(model, {}, [0])
    top, notes = ec3.entity_constants(model, {"c2": 0.454, "z_g": 0.15}, [0])
    assert base["zz"] == 0.0 and top["zz"] == pytest.approx(0.454 * 0.15) and not notes
    assert ec3_oracle.constants(model, {"c2": 0.454, "z_g": 0.15}, [0])["zz"] == pytest.approx(top["zz"])
    f = dict(Nmax=0.0, Nmin=0.0, Vy=0.0, Vz=0.0, My=0.0, Mz=60.75e3, Ma=45.6e3, Mb=60.75e3, Mc=45.6e3)
    m0 = ec3_oracle.rules(ec3_oracle.constants(model, {}, [0]), f)["Mcr"]
    m_top = ec3_oracle.rules(ec3_oracle.constants(model, {"c2": 0.454, "z_g": 0.15}, [0]), f)["Mcr"]
    m_bot = ec3_oracle.rules(ec3_oracle.constants(model, {"c2": 0.454, "z_g": -0.15}, [0]), f)["Mcr"]
    assert m_top < m0 < m_bot
    # hand formula
    c = ec3_oracle.constants(model, {"c2": 0.454, "z_g": 0.15}, [0])
    c1 = 12.5 * 60.75 / (2.5 * 60.75 + 3 * 45.6 + 4 * 60.75 + 3 * 45.6)
    ncr = math.pi ** 2 * c["E"] * c["Iy"] / c["Llt"] ** 2
    zz = 0.454 * 0.15
    assert m_top == pytest.approx(c1 * ncr * (math.sqrt(c["Iw"] / c["Iy"] + c["G"] * c["J"] / ncr + zz * zz) - zz), rel=1e-12)
    doc["model"]["sections"][0]["wagner_coeff"] = 0.02
    m3 = wire.parse_document(doc)
    assert ec3.entity_constants(m3, {"c3": 0.525}, [0])[0]["zz"] == pytest.approx(-0.525 * 0.02)
    assert ec3_oracle.constants(m3, {"c3": 0.525}, [0])["zz"] == pytest.approx(-0.525 * 0.02)


def _restrained_doc():
    """The 6 m IPE 300 beam of the reference's check cut into four members of one member set, with buckling restraints: the
    node at 3.0 m holds local y and twist, the node at 1.5 m holds local z."""
    doc = copy.deepcopy(helpers.golden("ec3_check_beam_ipe300")["doc"])
    m = doc["model"]
    n0, n1 = m["nodes"]
    base = m["members"][0]
    xs = [0.0, 1.5, 3.0, 4.5, 6.0]
    m["nodes"] = [dict(n0, id=1, X=xs[0])] + [dict(n0, id=10 + k, X=xs[k], nodal_support=None) for k in (1, 2, 3)] + [dict(n1, id=2, X=xs[4])]
    ids = [1, 11, 12, 13, 2]
    m["members"] = [dict(base, id=k + 1, start_node_id=ids[k], end_node_id=ids[k + 1]) for k in range(4)]
    m["member_sets"] = [dict(m["member_sets"][0], member_ids=[1, 2, 3, 4], buckling_restraints=[
        {"node_id": 12, "restrains_local_y": True, "restrains_local_z": False, "restrains_torsion": True},
        {"node_id": 11, "restrains_local_y": False, "restrains_local_z": True, "restrains_torsion": False}])]
    for lc in doc["analysis"]["load_cases"]:
        new = []
        for dl in lc.get("distributed_loads") or []:
            for k in range(4):
                new.append(dict(dl, id=100 * dl["id"] + k, member=k + 1))
        lc["distributed_loads"] = new
    return doc


def test_buckling_restraints_set_the_unrestrained_lengths():
    """BucklingRestraint (pydantic_models.py:129-143; bucklingrestraint.py:1-10): spacing of consecutive restraint nodes;
    buckling about y moves the beam along local z (held by restrains_local_z) and the other way round; twist -> LTB."""
    model = wire.parse_document(_restrained_doc())
    want = {(0, 1, 2, 3): (4.5, 3.0, 3.0), (0,): (1.5, 3.0, 3.0), (1,): (4.5, 3.0, 3.0), (2,): (4.5, 3.0, 3.0), (3,): (4.5, 3.0, 3.0)}
    for members, (lcry, lcrz, llt) in want.items():
        c, notes = ec3.entity_constants(model, {}, list(members))
        assert (c["Lcry"], c["Lcrz"], c["Llt"]) == pytest.approx((lcry, lcrz, llt)), members
        assert not [n for n in notes if "restraint" in n]
        o = ec3_oracle.constants(model, {}, list(members))
        assert (o["Lcry"], o["Lcrz"], o["Llt"]) == pytest.approx((lcry, lcrz, llt)), members
    # explicit lengths and factors win over the restraints
    doc = _restrained_doc()
    doc["model"]["member_sets"][0].update(buckling_length_y=2.0, effective_length_factor_z=0.5, ltb_length=1.0)
    m2 = wire.parse_document(doc)
    c, _ = ec3.entity_constants(m2, {}, [0, 1, 2, 3])
    assert (c["Lcry"], c["Lcrz"], c["Llt"]) == pytest.approx((2.0, 3.0, 1.0))
    o = ec3_oracle.constants(m2, {}, [0, 1, 2, 3])
    assert (o["Lcry"], o["Lcrz"], o["Llt"]) == pytest.approx((2.0, 3.0, 1.0))


def test_restrained_beam_has_a_higher_ltb_resistance():
    free = oracle_engine.solve_document(helpers.golden("ec3_check_beam_ipe300")["doc"])["results"]["unity_check_results"][0]
    held = oracle_engine.solve_document(_restrained_doc())["results"]["unity_check_results"][0]
    ltb = "LTB (6.3.2)"
    u_free = max(_rows({"governing": p})[ltb] for p in free["per_entity"])
    u_held = max(_rows({"governing": p})[ltb] for p in held["per_entity"])
    assert 0.0 < u_held < u_free


def _frame_doc():
    """Portal of the golden 2D fixture with an Ec3Steel check on every member: columns in compression, beam in bending."""
    doc = copy.deepcopy(helpers.golden("portal_2d_combinations")["doc"])
    cat = helpers.golden("ec3_check_beam_ipe300")["doc"]["model"]["sections"][0]
    for s in doc["model"]["sections"]:
        for k in ("wel_y", "wel_z", "wpl_y", "wpl_z", "i_w", "a_sy", "a_sz", "area", "i_y", "i_z", "j"):
            s[k] = cat[k]
    for m in doc["model"]["materials"]:
        m["yield_stress"] = 355e6
    doc["analysis"]["unity_checks"] = [
        {"id": "ec3", "name": "EC3", "applies_to": {"type": "AllMembers"}, "limit_state": None,
         "spec": {"Ec3Steel": {"gamma_m0": 1.0, "gamma_m1": 1.0, "include_buckling": True, "include_ltb": True}}},
        {"id": "ec3_c1", "name": "EC3, C1 = 1", "applies_to": {"type": "AllMembers"}, "limit_state": None,
         "spec": {"Ec3Steel": {"gamma_m1": 1.1, "c1": 1.0, "include_buckling": False}}}]
    return doc


def test_oracle_frame_check_has_compression_members():
    out = oracle_engine.solve_document(_frame_doc())["results"]
    u = {r["check_id"]: r for r in out["unity_check_results"]}
    assert len(u["ec3"]["per_entity"]) == len(_frame_doc()["model"]["members"])
    assert any(_rows({"governing": p})["Flexural buckling y (6.3.1)"] > 0 for p in u["ec3"]["per_entity"])
    assert all(_rows({"governing": p})["Flexural buckling y (6.3.1)"] == 0 for p in u["ec3_c1"]["per_entity"])


@pytest.mark.gpu
@pytest.mark.parametrize("make", [lambda: helpers.golden("ec3_check_beam_ipe300")["doc"],
                                  lambda: helpers.golden("ec3_check_beam_hea200_fixed")["doc"], _frame_doc, _restrained_doc])
def test_engine_matches_oracle(make):
    from fers_core_b200 import engine
    doc = make()
    got = json.loads(engine.calculate_from_json(json.dumps(doc)))["results"]["unity_check_results"]
    want = oracle_engine.solve_document(doc)["results"]["unity_check_results"]
    assert [g["check_id"] for g in got] == [w["check_id"] for w in want]
    for g, w in zip(got, want):
        assert g["status"] == w["status"] and g["max_utilization"] == pytest.approx(w["max_utilization"], rel=1e-9)
        for pg, pw in zip(g["per_entity"], w["per_entity"]):
            assert pg["entity"] == pw["entity"] and pg["governing_combination_id"] == pw["governing_combination_id"]
            assert pg["utilization"] == pytest.approx(pw["utilization"], rel=1e-9, abs=1e-300)
            rg, rw = _rows({"governing": pg}), _rows({"governing": pw})
            for label, val in rw.items():
                assert rg[label] == pytest.approx(val, rel=1e-9, abs=1e-12), label
