"""EN 1993-1-1 steel member check (``Ec3Steel`` unity-check spec) as a program for the unity-check kernel.

Wire contract: ``UnityCheckDefinition.spec = {"Ec3Steel": Ec3SteelSpec}`` (fers_core/types/pydantic_models.py:1448-1469;
builder fers_core/unity_checks.py:171-211; acceptance tests/functionality/test_check_beam.py:25-61).  Section data:
``Section.wel_* / wpl_* / i_w / a_s*`` (1052-1114), ``Section.ec3`` (``Ec3SectionParams`` 197-223: section class, buckling
curves), material ``yield_stress`` (1504); member-set overrides ``buckling_length_*``, ``effective_length_factor_*``,
``ltb_length`` (548-571).  The catalogue's own note says what the closed engine assumes where a section carries no EC3
parameters: "the solver defaults the class from wpl presence and buckling curves to B"
(fers_core/sections/steel_sections.generated.json ``_note``) -- followed here.

The check is COMPILED, not interpreted on the host: ``program(...)`` returns a Generic specification (variables bound to
member forces, per-entity section constants and expressions) that ``unity.compile_check`` turns into the same postfix code
``csrc/unity.cu`` evaluates for user-written Generic checks -- one warp per entity, lanes over load combinations, envelope
over the applicable combinations.  Rules (FERS axes: local z is the major axis of an I section, ``i_z > i_y``):

  6.2.3 / 6.2.4   N_Ed / (A f_y / gM0)                                        tension / compression
  6.2.5           M_y,Ed / (W_y f_y / gM0),  M_z,Ed / (W_z f_y / gM0)           W = W_pl (class 1, 2) or W_el (class 3)
  6.2.6           V_Ed / (A_v f_y / (sqrt 3 gM0))                               when the section carries shear areas
  6.2.1 (7)       n + m_y + m_z                                                 conservative linear N + M
  6.3.1           N_c,Ed / (chi A f_y / gM1), chi from the buckling curve (default b), L_cr = member(-set) length or override
  6.3.2           M_z,Ed / (chi_LT W_z f_y / gM1); M_cr = C1 pi^2 E I_y / L^2 [sqrt(I_w / I_y + L^2 G J / (pi^2 E I_y) + zz^2) - zz],
                  zz = C2 z_g - C3 z_j (load height above the shear centre, mono-symmetry z_j = Section.wagner_coeff; NCCI SN003),
                  C1 given or from the quarter-point moments 12.5 M_max / (2.5 M_max + 3 M_A + 4 M_B + 3 M_C); general case (6.56)
  6.3.3           N_c / N_b,Rd,min + M_z / M_b,Rd + M_y / M_c,Rd,y  (interaction factors k = 1: conservative; Annex A / B not evaluated)

Unpinned beyond the reference's own hand calculation (test_check_beam.py:25-45: bending z = 60.75 / 147.6, LTB governs):
the engine that defines the remaining numbers is closed."""
from __future__ import annotations

import numpy as np

from . import wire

ALPHA = {"A0": 0.13, "A": 0.21, "B": 0.34, "C": 0.49, "D": 0.76}
BIG = 1.0e30          # capacity stand-in of a rule that does not apply to an entity (utilisation -> 0)
TINY = 1.0e-30

# (variable name, trace label) of every rule, in report order
RULES = [("u_n", "Axial (6.2.3 / 6.2.4)"), ("u_my", "Bending y (6.2.5)"), ("u_mz", "Bending z (6.2.5)"),
         ("u_vy", "Shear y (6.2.6)"), ("u_vz", "Shear z (6.2.6)"), ("u_nm", "N + M (6.2.1(7))"),
         ("u_by", "Flexural buckling y (6.3.1)"), ("u_bz", "Flexural buckling z (6.3.1)"), ("u_ltb", "LTB (6.3.2)"),
         ("u_int", "N + M buckling (6.3.3, k = 1)")]


def _num(x, default=np.nan):
    return default if x is None else float(x)


def entity_constants(model: wire.ModelArrays, spec: dict, members: list) -> tuple:
    """Per-entity constants of the program for an entity made of ``members`` (a member or the chain of a member set).
    Returns (dict name -> value, list of notes)."""
    sec, mat = model.member_raw(members[0])
    notes = []
    ec3 = sec.get("ec3") or {}
    fy, E, G = _num(mat.get("yield_stress")), _num(mat.get("e_mod")), _num(mat.get("g_mod"))
    A, Iy, Iz, J = _num(sec.get("area")), _num(sec.get("i_y")), _num(sec.get("i_z")), _num(sec.get("j"))
    Iw = _num(sec.get("i_w"), 0.0)
    cls = ec3.get("section_class")
    plastic = (cls in (1, 2)) if cls else (sec.get("wpl_y") is not None or sec.get("wpl_z") is not None)
    if cls == 4:
        A = _num(ec3.get("a_eff"), A)
        notes.append("class 4: effective properties as given, no plate-buckling analysis")

    def modulus(ax):
        if cls == 4 and ec3.get(f"wel_{ax}_eff") is not None:
            return float(ec3[f"wel_{ax}_eff"])
        w = sec.get(f"wpl_{ax}") if plastic else None
        if w is None:
            w = sec.get(f"wel_{ax}")
        return np.nan if w is None else float(w)
    Wy, Wz = modulus("y"), modulus("z")
    if not np.isfinite(Wy) or not np.isfinite(Wz):
        notes.append("section carries no wel / wpl: bending rules not evaluated")
    # shear areas: V_y (along local y) is carried by the area paired with bending about z (a_sz <-> i_z)
    Avy, Avz = _num(sec.get("a_sz")), _num(sec.get("a_sy"))
    L = float(sum(np.linalg.norm(wire.member_axis(model, m)) for m in members))
    ms = _member_set_of(model, members)
    Lcr = {}
    for ax in ("y", "z"):
        if ms.get(f"buckling_length_{ax}") is not None:
            Lcr[ax] = float(ms[f"buckling_length_{ax}"])
        elif ms.get(f"effective_length_factor_{ax}") is not None:
            Lcr[ax] = float(ms[f"effective_length_factor_{ax}"]) * L
        else:
            Lcr[ax] = L
    Llt = float(ms["ltb_length"]) if ms.get("ltb_length") is not None else L
    if ms.get("buckling_restraints"):
        # BucklingRestraint (pydantic_models.py:129-143; fers_core/members/bucklingrestraint.py:1-10: "the spacing between
        # consecutive restraint nodes yields the unrestrained length used for the corresponding buckling check"): explicit
        # lengths and factors win (548-566); buckling about y moves the beam along local z and the other way round; a twist
        # restraint is a lateral-torsional restraint
        spans = restraint_spans(model, ms, members)
        if spans is None:
            notes.append("buckling_restraints not evaluated: the member set is not a chain of members")
        else:
            for ax, held in (("y", "z"), ("z", "y")):
                if ms.get(f"buckling_length_{ax}") is None and ms.get(f"effective_length_factor_{ax}") is None:
                    Lcr[ax] = spans[held]
            if ms.get("ltb_length") is None:
                Llt = spans["t"]
    alpha = {k: ALPHA.get(str(ec3.get(f"buckling_curve_{k}") or "B").upper(), ALPHA["B"]) for k in ("y", "z", "lt")}
    # general M_cr (NCCI SN003, k = k_w = 1): load height z_g above the shear centre with C2, mono-symmetry z_j (the
    # section's `wagner_coeff`, pydantic_models.py:1092-1094) with C3; both zero -> the doubly symmetric form
    zz = _num(spec.get("c2"), 0.0) * _num(spec.get("z_g"), 0.0) - _num(spec.get("c3"), 0.0) * _num(sec.get("wagner_coeff"), 0.0)
    ltb_applies = bool(spec.get("include_ltb", True)) and np.isfinite(Wz) and Iz > Iy and J > 0.0
    buckling = bool(spec.get("include_buckling", True))
    c = dict(fy=fy, E=E, G=G, A=A, Iy=Iy, Iz=Iz, J=J, Iw=Iw, L=L,
             Wy=Wy if np.isfinite(Wy) else BIG, Wz=Wz if np.isfinite(Wz) else BIG,
             Avy=Avy if np.isfinite(Avy) and Avy > 0 else BIG, Avz=Avz if np.isfinite(Avz) and Avz > 0 else BIG,
             Lcry=Lcr["y"], Lcrz=Lcr["z"], Llt=Llt, ay=alpha["y"], az=alpha["z"], alt=alpha["lt"],
             gm0=float(spec.get("gamma_m0") or 1.0), gm1=float(spec.get("gamma_m1") or 1.0),
             c1=_num(spec.get("c1"), 0.0), zz=zz, on_b=1.0 if buckling else 0.0, on_lt=1.0 if ltb_applies else 0.0)
    return c, notes


def restraint_spans(model, ms: dict, members: list):
    """Unrestrained lengths of an entity (member rows ``members``) inside the member set ``ms`` (one physical beam): per
    restrained quantity ("y", "z": displacement along that local axis; "t": twist) the longest interval between
    consecutive restraint points -- the two ends of the beam and the nodes of ``buckling_restraints`` that hold the
    quantity -- that overlaps the entity.  None when the set's members do not form a chain."""
    index = {int(i): r for r, i in enumerate(model.m_ids)}
    try:
        rows = [index[int(i["id"] if isinstance(i, dict) else i)] for i in (ms.get("member_ids") or ms.get("members") or [])]
    except KeyError:
        return None
    if not rows:
        return None
    # walk the chain: station of every node along the beam
    first = rows[0]
    a, b = int(model.m_n1[first]), int(model.m_n2[first])
    if len(rows) > 1:
        nxt = {int(model.m_n1[rows[1]]), int(model.m_n2[rows[1]])}
        if a in nxt and b not in nxt:
            a, b = b, a
    station = {a: 0.0}
    span_of = {}
    at, s_at = a, 0.0
    for r in rows:
        n1, n2 = int(model.m_n1[r]), int(model.m_n2[r])
        if at not in (n1, n2):
            return None
        far = n2 if at == n1 else n1
        length = float(np.linalg.norm(wire.member_axis(model, r)))
        span_of[r] = (s_at, s_at + length)
        s_at += length
        station[far] = s_at
        at = far
    if any(m not in span_of for m in members):
        return None
    lo = min(span_of[m][0] for m in members)
    hi = max(span_of[m][1] for m in members)
    node_index = {int(i): k for k, i in enumerate(model.node_ids)}
    out = {}
    for key, flag in (("y", "restrains_local_y"), ("z", "restrains_local_z"), ("t", "restrains_torsion")):
        pts = {0.0, s_at}
        for br in ms.get("buckling_restraints") or []:
            node = node_index.get(int(br["node_id"]["id"] if isinstance(br["node_id"], dict) else br["node_id"]))
            if br.get(flag) and node in station:
                pts.add(station[node])
        pts = sorted(pts)
        eps = 1e-9 * max(s_at, 1.0)
        out[key] = max((q - p for p, q in zip(pts[:-1], pts[1:]) if min(q, hi) - max(p, lo) > eps), default=hi - lo)
    return out


def _member_set_of(model, members: list) -> dict:
    """Raw member-set record that lists exactly these members (set entities) or contains the single member."""
    best = {}
    for raw in getattr(model, "member_sets_raw", []) or []:
        ids = {int(i["id"] if isinstance(i, dict) else i) for i in (raw.get("member_ids") or raw.get("members") or [])}
        mine = {int(model.m_ids[m]) for m in members}
        if mine == ids:
            return raw
        if mine <= ids and not best:
            best = raw
    return best


def _force(comp, agg, fraction=None):
    a = {"type": agg} if fraction is None else {"type": "AtFraction", "fraction": fraction}
    return {"Quantity": {"MemberForce": {"component": comp, "aggregation": a}}}


def program(const_names: list) -> dict:
    """The Generic specification of the check: per-entity constants (bound by name), member-force variables, expressions."""
    v = [{"name": n, "source": {"Quantity": {"_EntityConstant": n}}} for n in const_names]
    v += [{"name": "Nmax", "source": _force("N", "Max")}, {"name": "Nmin", "source": _force("N", "Min")},
          {"name": "Vy", "source": _force("Vy", "MaxAbs")}, {"name": "Vz", "source": _force("Vz", "MaxAbs")},
          {"name": "My", "source": _force("My", "MaxAbs")}, {"name": "Mz", "source": _force("Mz", "MaxAbs")},
          {"name": "Ma", "source": _force("Mz", None, 0.25)}, {"name": "Mb", "source": _force("Mz", None, 0.5)},
          {"name": "Mc", "source": _force("Mz", None, 0.75)}]

    def e(name, text):
        v.append({"name": name, "source": {"Expression": text}})
    e("Nt", "max(Nmax, 0)")
    e("Nc", "max(0 - Nmin, 0)")
    e("NRd", "A * fy / gm0")
    e("u_n", "max(Nt, Nc) / NRd")
    e("u_my", "My / (Wy * fy / gm0)")
    e("u_mz", "Mz / (Wz * fy / gm0)")
    e("u_vy", "Vy / (Avy * fy / (sqrt(3) * gm0))")
    e("u_vz", "Vz / (Avz * fy / (sqrt(3) * gm0))")
    e("u_nm", "u_n + u_my + u_mz")
    pi2 = repr(float(np.pi ** 2))
    for ax, inertia in (("y", "Iy"), ("z", "Iz")):
        e(f"Ncr{ax}", f"{pi2} * E * {inertia} / (Lcr{ax} * Lcr{ax})")
        e(f"lam{ax}", f"sqrt(A * fy / Ncr{ax})")
        e(f"phi{ax}", f"0.5 * (1 + a{ax} * (lam{ax} - 0.2) + lam{ax} * lam{ax})")
        e(f"chi{ax}", f"min(1, 1 / (phi{ax} + sqrt(max(phi{ax} * phi{ax} - lam{ax} * lam{ax}, 0))))")
        e(f"u_b{ax}", f"on_b * Nc / (chi{ax} * A * fy / gm1)")
    # lateral-torsional buckling about the major (z) axis; C1 from the quarter-point moments unless given
    e("C1q", f"12.5 * (Mz + {TINY!r}) / (2.5 * (Mz + {TINY!r}) + 3 * abs(Ma) + 4 * abs(Mb) + 3 * abs(Mc))")
    e("C1", "max(c1, 0) + (1 - min(c1 * 1e30, 1)) * C1q")          # c1 > 0: the given value; c1 = 0: the computed one
    e("Mcr", f"C1 * {pi2} * E * Iy / (Llt * Llt) * (sqrt(Iw / Iy + Llt * Llt * G * J / ({pi2} * E * Iy) + zz * zz) - zz)")
    e("lamlt", "sqrt(Wz * fy / Mcr)")
    e("philt", "0.5 * (1 + alt * (lamlt - 0.2) + lamlt * lamlt)")
    e("chilt", "min(1, 1 / (philt + sqrt(max(philt * philt - lamlt * lamlt, 0))))")
    e("u_ltb", "on_lt * Mz / (chilt * Wz * fy / gm1)")
    e("u_int", "on_b * (max(u_by, u_bz) + max(u_ltb, u_mz * gm1 / gm0) + u_my * gm1 / gm0)")
    e("gov", "max(" + ", ".join(n for n, _ in RULES) + ")")
    return {"variables": v, "demand": "gov", "capacity": "1"}


CONST_NAMES = ["fy", "E", "G", "A", "Iy", "Iz", "J", "Iw", "L", "Wy", "Wz", "Avy", "Avz", "Lcry", "Lcrz", "Llt", "ay", "az", "alt",
               "gm0", "gm1", "c1", "zz", "on_b", "on_lt"]
LABELS = dict(RULES, gov="Governing", C1="C1 (6.3.2)", Mcr="M_cr", chilt="chi_LT", chiy="chi_y", chiz="chi_z", lamlt="lambda_LT",
              lamy="lambda_y", lamz="lambda_z")
