"""CPU ORACLE -- a second, minimal reading of the wire document.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The oracle solves on the arrays `fers_core_b200.wire.parse_document` produces ("What's weak" 2: a bug in that shared
parser is invisible to every GPU <-> oracle comparison).  This module reads the core of a document a second time, straight from
the schema text (fers_core/types/pydantic_models.py: Node, Member 1507-1549, Section 1052-1114, Material 1497-1504,
NodalSupport 2104-2115 / SupportCondition 1956-1959, NodalLoad 1684-1689, NodalMoment, DistributedLoad 1434-1446,
LoadCombination 491-505), with dictionaries and plain loops, for tests/test_wire_check.py to hold against the parser's arrays:
coordinates, connectivity, member types and roll angles, the first six stiffnesses of every member's section, support codes
and spring stiffnesses, hinge release springs, plate triangles with thickness / E / nu, nodal load vectors, distributed member
loads, combination factors.  (End offsets, shear-centre offsets, products of inertia, stiffness curves and surface loads are
left to their own closed-form tests.)"""
from __future__ import annotations

import math

DOFS = ("X", "Y", "Z", "RX", "RY", "RZ")
CODES = {"free": 0, "fixed": 1, "spring": 2, "positiveonly": 3, "negativeonly": 4}
TYPES = {"normal": 0, "truss": 1, "tension": 2, "compression": 3, "rigid": 4, "cable": 5}


def _key(s) -> str:
    return "".join(ch for ch in str(s).lower() if ch.isalnum())


def _id(v):
    return int(v["id"]) if isinstance(v, dict) else int(v)


def read(doc: dict) -> dict:
    model, analysis = doc["model"], doc["analysis"]
    opts = analysis.get("options") or {}
    shear = opts.get("include_shear_deformation", True) is not False
    two_d = _key(opts.get("dimensionality") or "3D") == "2d"
    nodes = model.get("nodes") or []
    supports = {int(s["id"]): s for s in model.get("nodal_supports") or []}
    materials = {int(m["id"]): m for m in model.get("materials") or []}
    sections = {int(s["id"]): s for s in model.get("sections") or []}
    out = {"node_ids": [int(n["id"]) for n in nodes],
           "xyz": [[float(n["X"]), float(n["Y"]), float(n["Z"])] for n in nodes],
           "bc_code": [], "bc_k": []}
    for n in nodes:
        code, k = [0] * 6, [0.0] * 6
        sup = n.get("nodal_support")
        if sup is not None:
            sup = sup if isinstance(sup, dict) else supports[int(sup)]
            for d, name in enumerate(DOFS):
                cond = sup.get(name)
                if cond is None:
                    continue
                code[d] = CODES[_key(cond.get("condition_type", "Fixed"))]
                if code[d] == 2:
                    curve = cond.get("stiffness_curve")
                    k[d] = float(cond["stiffness"]) if not (curve and curve.get("points")) else math.nan   # curves: own tests
                elif code[d] in (3, 4) and cond.get("stiffness"):
                    k[d] = float(cond["stiffness"])
        if two_d:
            for d in (2, 3, 4):
                if code[d] == 0:
                    code[d] = 1
        out["bc_code"].append(code); out["bc_k"].append(k)
    index = {nid: i for i, nid in enumerate(out["node_ids"])}
    members = model.get("members") or []
    out["m_ids"] = [int(m["id"]) for m in members]
    out["m_n1"] = [index[int(m["start_node_id"])] for m in members]
    out["m_n2"] = [index[int(m["end_node_id"])] for m in members]
    out["m_type"] = [TYPES[_key(m.get("member_type", "Normal"))] for m in members]
    out["m_roll"] = [math.radians(float(m.get("rotation_angle") or 0.0)) for m in members]
    stiff = []
    for m in members:
        sid = m.get("section")
        if sid is None:
            stiff.append(None)
            continue
        s = sections[int(sid)]
        mat = s["material"] if isinstance(s["material"], dict) else materials[int(s["material"])]
        E, G = float(mat["e_mod"]), float(mat["g_mod"])
        asy = float(s["a_sy"]) if (shear and s.get("a_sy")) else 0.0
        asz = float(s["a_sz"]) if (shear and s.get("a_sz")) else 0.0
        plain = not s.get("i_yz") and not s.get("principal_axis_angle")
        stiff.append([E * float(s["area"]), E * float(s["i_y"]) if plain else math.nan, E * float(s["i_z"]) if plain else math.nan,
                      G * float(s["j"]), G * asy, G * asz])
    out["m_stiffness"] = stiff
    # hinges: twelve release springs per member (start u v w rx ry rz, end ...), -1 = rigidly connected (MemberHinge 2062-2087:
    # "release" fields are spring stiffnesses; a stiffness curve overrides its constant: left to the curve tests -> nan)
    hinges = {int(h["id"]): h for h in model.get("member_hinges") or []}
    fields = ("translational_release_vx", "translational_release_vy", "translational_release_vz",
              "rotational_release_mx", "rotational_release_my", "rotational_release_mz")
    curves = ("stiffness_curve_vx", "stiffness_curve_vy", "stiffness_curve_vz", "stiffness_curve_mx", "stiffness_curve_my",
              "stiffness_curve_mz")
    releases = []
    for m in members:
        if m.get("start_hinge") is None and m.get("end_hinge") is None:
            releases.append(None)
            continue
        row = []
        for hid in (m.get("start_hinge"), m.get("end_hinge")):
            h = hinges[_id(hid)] if hid is not None else {}
            for f, c in zip(fields, curves):
                if h.get(c) and h[c].get("points"):
                    row.append(math.nan)
                else:
                    row.append(-1.0 if h.get(f) is None else max(float(h[f]), 0.0))
        releases.append(row)
    out["m_release"] = releases
    # plates: triangles (a quad is cut into (0, 1, 2) and (0, 2, 3)), thickness, E, nu = E / 2G - 1 (PlateElement 1700-1726)
    plates = []
    for p in model.get("plate_elements") or []:
        ids = [_id(v) for v in p["node_ids"]]
        mat = p["material"] if isinstance(p["material"], dict) else materials[int(p["material"])]
        E, G = float(mat["e_mod"]), float(mat["g_mod"])
        for tri in ([ids] if len(ids) == 3 else [[ids[0], ids[1], ids[2]], [ids[0], ids[2], ids[3]]]):
            plates.append((int(p["id"]), [index[v] for v in tri], float(p["thickness"]), E, E / (2.0 * G) - 1.0))
    out["plates"] = plates
    member_index = {mid: i for i, mid in enumerate(out["m_ids"])}
    cases = []
    for lc in analysis.get("load_cases") or []:
        nodal = {}
        for key, lo in (("nodal_loads", 0), ("nodal_moments", 3)):
            for rec in lc.get(key) or []:
                i = index[_id(rec["node"])]
                v = nodal.setdefault(i, [0.0] * 6)
                for c in range(3):
                    v[lo + c] += float(rec["magnitude"]) * float(rec["direction"][c])
        dist = []
        for rec in lc.get("distributed_loads") or []:
            q1 = float(rec["magnitude"])
            q2 = q1 if rec.get("end_magnitude") is None else float(rec["end_magnitude"])
            a = float(rec.get("start_frac") or 0.0)
            b = 1.0 if rec.get("end_frac") is None else float(rec["end_frac"])
            d = [float(x) for x in rec["direction"]]
            dist.append((member_index[_id(rec["member"])], [q1 * x for x in d] + [q2 * x for x in d] + [a, b]))
        cases.append({"id": int(lc["id"]), "nodal": nodal, "distributed": dist})
    out["load_cases"] = cases
    case_col = {c["id"]: i for i, c in enumerate(cases)}
    combos = []
    for c in analysis.get("load_combinations") or []:
        row = [0.0] * len(cases)
        fac = c.get("load_cases_factors", c.get("load_case_factors")) or {}
        items = fac.items() if isinstance(fac, dict) else [(e["load_case_id"], e["factor"]) for e in fac]
        for k, v in items:
            if int(k) in case_col:
                row[case_col[int(k)]] += float(v)
        combos.append((int(c["id"]), row))
    out["combinations"] = combos
    return out
