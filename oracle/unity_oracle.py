"""CPU oracle of the Generic unity checks -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Restates the evaluation the wire schema documents (fers_core/types/pydantic_models.py: GenericSpec
2136-2151 "utilization = demand / capacity"; quantity sources 829-969; UnityCheckDefinition 2202-2220
"enveloped over the applicable load combinations"; thresholds [0.8, 0.95, 1.0]) on the host result
arrays of ``oracle.fers_oracle.analyze``.  Expressions are evaluated with Python's own parser (``ast``),
independently of the postfix compiler in fers_core_b200/unity.py.

Parity unpinned: the engine that defines these numbers is the closed wheel fers_calculations==0.2.43;
the only in-repo numeric statement is example 100 (chord-relative sag == global mid-span deflection on
rigid supports, fers_core/examples/100_Unity_Check_Beam_Deflection_SLS.py:118-128), which
tests/test_unity_checks.py reproduces.
"""
from __future__ import annotations

import ast
import math

import numpy as np

FORCE = {"N": 0, "Vy": 1, "Vz": 2, "Mx": 3, "My": 4, "Mz": 5}
DISP = {"Dx": 0, "Dy": 1, "Dz": 2, "Rx": 3, "Ry": 4, "Rz": 5}
SECTION = {"Area": "area", "Iy": "i_y", "Iz": "i_z", "J": "j", "Iw": "i_w", "WelY": "wel_y", "WelZ": "wel_z",
           "WplY": "wpl_y", "WplZ": "wpl_z", "H": "h", "B": "b", "Asy": "a_sy", "Asz": "a_sz"}
MATERIAL = {"E": "e_mod", "G": "g_mod", "Fy": "yield_stress", "Density": "density"}
_FUNCS = {"abs": abs, "sqrt": math.sqrt, "min": min, "max": max, "pow": math.pow}


def _eval(text: str, env: dict) -> float:
    tree = ast.parse(text.replace("^", "**"), mode="eval")

    def go(n):
        if isinstance(n, ast.Expression):
            return go(n.body)
        if isinstance(n, ast.Constant):
            return float(n.value)
        if isinstance(n, ast.Name):
            return env[n.id]
        if isinstance(n, ast.UnaryOp):
            v = go(n.operand)
            return -v if isinstance(n.op, ast.USub) else v
        if isinstance(n, ast.BinOp):
            a, b = go(n.left), go(n.right)
            try:
                if isinstance(n.op, ast.Add): return a + b
                if isinstance(n.op, ast.Sub): return a - b
                if isinstance(n.op, ast.Mult): return a * b
                if isinstance(n.op, ast.Div): return a / b
                if isinstance(n.op, ast.Pow): return a ** b
            except ZeroDivisionError:
                return math.nan if a == 0 or math.isnan(a) else math.copysign(math.inf, a)
        if isinstance(n, ast.Call):
            return _FUNCS[n.func.id.lower()](*[go(a) for a in n.args])
        raise ValueError(f"unsupported expression node {ast.dump(n)}")
    return go(tree)


def _entities(model, sel):
    kind = sel.get("type", "AllMembers")
    idx = {int(i): k for k, i in enumerate(model.m_ids)}
    if kind == "AllMembers":
        return [({"type": "Member", "member_id": int(model.m_ids[k])}, [k]) for k in range(model.n_members)]
    if kind == "Members":
        return [({"type": "Member", "member_id": int(i)}, [idx[int(i)]]) for i in sel["ids"]]
    if kind == "MemberSets":
        return [({"type": "MemberSet", "memberset_id": int(i)}, list(model.member_sets[int(i)])) for i in sel["ids"]]
    if kind == "MemberClassification":
        return [({"type": "Member", "member_id": int(model.m_ids[k])}, [k]) for k in range(model.n_members)
                if model.m_classification[k] == sel["value"]]
    if kind in ("AllPlates", "PlateElements"):
        ids = list(dict.fromkeys(int(i) for i in model.p_ids)) if kind == "AllPlates" else [int(i) for i in sel["ids"]]
        return [({"type": "PlateElement", "plate_element_id": i}, [r for r, pid in enumerate(model.p_ids) if int(pid) == i])
                for i in ids]
    if kind == "PlateSurfaces":         # one entity per surface: every element meshed from it
        return [({"type": "PlateSurface", "plate_surface_id": int(i)},
                 [r for r in range(model.n_plates) if int(model.p_surface[r]) == int(i)]) for i in sel["ids"]]
    raise ValueError(kind)


def _nodes_along(model, members):
    nodes = [int(model.m_n1[members[0]]), int(model.m_n2[members[0]])]
    for m in members[1:]:                       # members listed tip to tail
        a, b = int(model.m_n1[m]), int(model.m_n2[m])
        nodes.append(b if a == nodes[-1] else a)
    return nodes


def evaluate(model, res) -> list:
    """Unity-check results (same dict layout as fers_core_b200.unity.check_result) from oracle results."""
    plan = res.plan
    out = []
    for d in model.unity_checks:
        is_ec3 = "Ec3Steel" in d["spec"]
        g = None if is_ec3 else d["spec"]["Generic"]
        cols = [c for c in range(plan.n_rhs) if plan.kinds[c] == "Loadcombination"]
        ids = [int(i) for i in (d.get("load_combination_ids") or [])]
        if ids:
            cols = [c for c in cols if int(plan.ids[c]) in ids]
        elif d.get("limit_state") is not None and any(x is not None for x in model.combo_limit_state):
            ls = {int(model.combo_ids[i]): x for i, x in enumerate(model.combo_limit_state)}
            cols = [c for c in cols if ls[int(plan.ids[c])] in (None, d["limit_state"])]
        per = []
        for ref, members in _entities(model, d.get("applies_to") or {"type": "AllMembers"}):
            if ref["type"] in ("PlateElement", "PlateSurface"):
                nodes = list(dict.fromkeys(int(n) for r in members for n in model.p_nodes[r]))
            else:
                nodes = _nodes_along(model, members)
            best = None
            if is_ec3:
                from oracle import ec3_oracle
                if not model.member_raw(members[0])[0]:
                    continue                                          # rigid links carry no section
                cst = ec3_oracle.constants(model, d["spec"]["Ec3Steel"] or {}, members)
                mf = lambda comp, agg, c, **kw: _quantity(model, res, c, members, nodes, {"MemberForce": {   # noqa: E731
                    "component": comp, "aggregation": dict(type=agg, **kw)}})
                rows = None
                for c in cols:
                    f = dict(Nmax=mf("N", "Max", c), Nmin=mf("N", "Min", c), Vy=mf("Vy", "MaxAbs", c), Vz=mf("Vz", "MaxAbs", c),
                             My=mf("My", "MaxAbs", c), Mz=mf("Mz", "MaxAbs", c), Ma=mf("Mz", "AtFraction", c, fraction=0.25),
                             Mb=mf("Mz", "AtFraction", c, fraction=0.5), Mc=mf("Mz", "AtFraction", c, fraction=0.75))
                    r = ec3_oracle.rules(cst, f)
                    if best is None or r["gov"] > best[0]:
                        best, rows = (r["gov"], r["gov"], 1.0, c), r
                per.append({"entity": ref, "utilization": 0.0, "demand": 0.0, "capacity": None, "governing_combination_id": None}
                           if best is None else
                           {"entity": ref, "utilization": best[0], "demand": best[1], "capacity": best[2],
                            "governing_combination_id": int(plan.ids[best[3]]), "rules": rows})
                continue
            for c in cols:
                env = {}
                pending = list(g.get("variables") or [])
                while pending:
                    progressed = False
                    for v in list(pending):
                        src = v["source"]
                        if "Expression" in src:
                            try:
                                env[v["name"]] = _eval(src["Expression"], env)
                            except KeyError:
                                continue
                        else:
                            env[v["name"]] = _quantity(model, res, c, members, nodes, src["Quantity"])
                        pending.remove(v); progressed = True
                    if not progressed:
                        raise ValueError("cyclic unity-check variables")
                dem, cap = _eval(g["demand"], env), _eval(g["capacity"], env)
                try:
                    u = dem / cap
                except ZeroDivisionError:
                    u = math.nan if dem == 0 else math.copysign(math.inf, dem)
                if best is None or (math.isnan(best[0]) and not math.isnan(u)) or (not math.isnan(u) and u > best[0]):
                    best = (u, dem, cap, c)
            if best is None:
                per.append({"entity": ref, "utilization": 0.0, "demand": 0.0, "capacity": None, "governing_combination_id": None})
            else:
                per.append({"entity": ref, "utilization": best[0], "demand": best[1], "capacity": best[2],
                            "governing_combination_id": int(plan.ids[best[3]])})
        out.append({"check_id": str(d["id"]), "per_entity": per,
                    "max_utilization": max((p["utilization"] for p in per), default=0.0)})
    return out


def _quantity(model, res, c, members, nodes, q) -> float:
    (kind, v), = q.items()
    if isinstance(v, dict) and v.get("of") == "ReferenceMember":
        # the member the entity's first member points at (Member.reference_member, example 101)
        rm = int(model.m_ref_member[members[0]])
        if rm < 0:
            return math.nan
        members, nodes = [rm], [int(model.m_n1[rm]), int(model.m_n2[rm])]
    if kind == "Constant":
        return float(v)
    if kind == "DimensionedConstant":
        return float(v["value"])
    if kind == "Section":
        x = model.member_raw(members[0])[0].get(SECTION[v["property"]])
        return math.nan if x is None else float(x)
    if kind == "Material":
        x = model.member_raw(members[0])[1].get(MATERIAL[v["property"]])
        return math.nan if x is None else float(x)
    if kind == "Geometry":
        from oracle.fers_oracle import member_offsets
        off = member_offsets(model)
        return float(sum(np.linalg.norm(model.xyz[model.m_n2[m]] - model.xyz[model.m_n1[m]] + off[m, 1] - off[m, 0]) for m in members))
    if kind == "MemberForce":
        k = FORCE[v["component"]]
        agg = (v.get("aggregation") or {"type": "MaxAbs"})
        t = agg["type"]
        if t == "Start":
            return float(res.f_local[c, members[0], k])
        if t == "End":
            return float(res.f_local[c, members[-1], 6 + k])
        if t == "AtFraction":
            st = int(round(float(agg.get("fraction", 0.5)) * 4))
            m = members[0]
            return float(res.f_local[c, m, k] if st == 0 else -res.f_local[c, m, 6 + k] if st == 4 else res.sec[c, m, st - 1, k])
        hi = max(float(res.lmax[c, m, k]) for m in members)
        lo = min(float(res.lmin[c, m, k]) for m in members)
        return hi if t == "Max" else lo if t == "Min" else max(abs(hi), abs(lo))
    if kind == "PlateStress":
        # surface stresses n / t +- 6 m / t^2 at the centroid of every TRI3 row of the plate, both surfaces
        vals = []
        for row in members:
            nx, ny, nxy, mx, my, mxy = (float(x) for x in res.p_res[c, row, :6])
            t = float(model.p_data[row, 0])
            for side in (-1.0, 1.0):
                sx, sy, txy = nx / t + side * 6 * mx / t**2, ny / t + side * 6 * my / t**2, nxy / t + side * 6 * mxy / t**2
                mean, rad = 0.5 * (sx + sy), math.sqrt(0.25 * (sx - sy) ** 2 + txy ** 2)
                vals.append({"VonMises": math.sqrt(sx * sx - sx * sy + sy * sy + 3 * txy * txy), "PrincipalMax": mean + rad,
                             "PrincipalMin": mean - rad}[v.get("measure", "VonMises")])
        return min(vals) if v.get("measure") == "PrincipalMin" else max(vals)
    if kind == "Displacement":
        comp = v.get("component", "Magnitude")
        if comp == "Magnitude":
            return max(float(np.linalg.norm(res.U[c, n, :3])) for n in nodes)
        return max(abs(float(res.U[c, n, DISP[comp]])) for n in nodes)
    if kind in ("NodeDisplacement", "NodePosition"):
        at = v["at"]
        if at in ("Start", "End"):
            node = nodes[0] if at == "Start" else nodes[-1]
        elif at == "ReferenceNode":
            node = int(model.m_ref_node[members[0]])
        else:
            rm = int(model.m_ref_member[members[0]])
            node = -1 if rm < 0 else int(model.m_n1[rm] if at == "ReferenceMemberStart" else model.m_n2[rm])
        if node < 0:
            return math.nan
        if kind == "NodePosition":
            a = "XYZ".index(v["axis"])
            return float(model.xyz[node, a] + res.U[c, node, a])
        if v["component"] == "Magnitude":
            return float(np.linalg.norm(res.U[c, node, :3]))
        return float(res.U[c, node, DISP[v["component"]]])
    if kind == "MemberDeflection":
        p0, p1 = model.xyz[nodes[0]], model.xyz[nodes[-1]]
        u0, u1 = res.U[c, nodes[0], :3], res.U[c, nodes[-1], :3]
        ch = p1 - p0
        rel = []
        for n in nodes:
            s = float((model.xyz[n] - p0) @ ch / (ch @ ch))
            rel.append(res.U[c, n, :3] - (u0 + s * (u1 - u0)))
        comp = v.get("component", "Magnitude")
        if comp == "Magnitude":
            vals = [float(np.linalg.norm(r)) for r in rel]
        else:
            from fers_core_b200.unity import _member_frame   # geometry helper only (axis rule member.py:309-354)
            fr = _member_frame(model, members[0]).reshape(3, 3)
            vals = [float(r @ fr[1 if comp == "LocalY" else 2]) for r in rel]
        t = (v.get("aggregation") or {"type": "MaxAbs"})["type"]
        return {"Max": max(max(vals), 0.0), "Min": min(min(vals), 0.0), "Start": vals[0], "End": vals[-1]}.get(t, max(abs(x) for x in vals))
    raise ValueError(f"quantity {kind}")
