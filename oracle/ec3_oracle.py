"""CPU oracle of the EN 1993-1-1 member check (``Ec3Steel`` unity-check spec) -- TEST INFRASTRUCTURE ONLY.

The rules written out directly in Python floats for one entity and one load combination, independently of the postfix
program the product compiles (fers_core_b200/ec3.py + csrc/unity.cu).  Contract: Ec3SteelSpec
(fers_core/types/pydantic_models.py:1448-1469), Ec3SectionParams (197-223), member-set buckling lengths (548-571), the
catalogue note on defaults (fers_core/sections/steel_sections.generated.json ``_note``: class from wpl presence, curves b).
Pinned on the reference's own hand calculation (tests/functionality/test_check_beam.py:25-45: IPE300, 6 m, S235, UDL 10 kN/m
x 1.35: M_Ed = 60.75 kNm, M_c,Rd = 147.6 kNm -> 0.412 +- 3 %; LTB governs) -- tests/test_ec3.py.  The closed engine defines
everything beyond that: unpinned."""
from __future__ import annotations

import math

import numpy as np

ALPHA = {"A0": 0.13, "A": 0.21, "B": 0.34, "C": 0.49, "D": 0.76}
BIG, TINY = 1.0e30, 1.0e-30
RULES = ["u_n", "u_my", "u_mz", "u_vy", "u_vz", "u_nm", "u_by", "u_bz", "u_ltb", "u_int"]


def constants(model, spec: dict, members: list) -> dict:
    """Section / material / length constants of one entity (members = member rows of the entity, chain order)."""
    sec, mat = model.member_raw(members[0])
    ec3 = sec.get("ec3") or {}
    g = lambda d, k, default=math.nan: default if d.get(k) is None else float(d[k])   # noqa: E731
    cls = ec3.get("section_class")
    plastic = (cls in (1, 2)) if cls else (sec.get("wpl_y") is not None or sec.get("wpl_z") is not None)

    def W(ax):
        if cls == 4 and ec3.get(f"wel_{ax}_eff") is not None:
            return float(ec3[f"wel_{ax}_eff"])
        w = sec.get(f"wpl_{ax}") if plastic else None
        w = sec.get(f"wel_{ax}") if w is None else w
        return BIG if w is None else float(w)
    off = np.zeros((model.n_members, 2, 3))
    rows = np.nonzero(model.m_rel >= 0)[0]
    if len(rows):
        off[rows] = model.rel[model.m_rel[rows], 14:20].reshape(-1, 2, 3)
    L = float(sum(np.linalg.norm(model.xyz[model.m_n2[m]] - model.xyz[model.m_n1[m]] + off[m, 1] - off[m, 0]) for m in members))
    mine = {int(model.m_ids[m]) for m in members}
    ms = {}
    for raw in model.member_sets_raw:
        ids = {int(i["id"] if isinstance(i, dict) else i) for i in (raw.get("member_ids") or raw.get("members") or [])}
        if mine == ids:
            ms = raw
            break
        if mine <= ids and not ms:
            ms = raw

    def unrestrained(flag):
        """Longest stretch between restraint points (beam ends + nodes of `buckling_restraints` with ``flag``) that the
        entity touches (BucklingRestraint, pydantic_models.py:129-143; bucklingrestraint.py:1-10); None without restraints."""
        if not ms.get("buckling_restraints"):
            return None
        row_of = {int(i): r for r, i in enumerate(model.m_ids)}
        chain = [row_of[int(i["id"] if isinstance(i, dict) else i)] for i in (ms.get("member_ids") or ms.get("members") or [])]
        # node sequence of the beam, tip to tail
        seq = [int(model.m_n1[chain[0]]), int(model.m_n2[chain[0]])]
        if len(chain) > 1 and seq[1] not in (int(model.m_n1[chain[1]]), int(model.m_n2[chain[1]])):
            seq.reverse()
        for r in chain[1:]:
            a, b = int(model.m_n1[r]), int(model.m_n2[r])
            if seq[-1] not in (a, b):
                return None
            seq.append(b if a == seq[-1] else a)
        lengths = [float(np.linalg.norm(model.xyz[model.m_n2[r]] - model.xyz[model.m_n1[r]] + off[r, 1] - off[r, 0])) for r in chain]
        x = np.concatenate([[0.0], np.cumsum(lengths)])
        held_ids = {int(br["node_id"]) for br in ms["buckling_restraints"] if br.get(flag)}
        cuts = [0] + [k for k in range(1, len(seq) - 1) if int(model.node_ids[seq[k]]) in held_ids] + [len(seq) - 1]
        mine_pos = [k for k, r in enumerate(chain) if r in members]
        best = 0.0
        for c0, c1 in zip(cuts[:-1], cuts[1:]):
            if any(c0 <= k < c1 for k in mine_pos):                 # member k spans nodes k .. k+1
                best = max(best, x[c1] - x[c0])
        return best

    def lcr(ax):
        if ms.get(f"buckling_length_{ax}") is not None:
            return float(ms[f"buckling_length_{ax}"])
        if ms.get(f"effective_length_factor_{ax}") is not None:
            return float(ms[f"effective_length_factor_{ax}"]) * L
        free = unrestrained("restrains_local_z" if ax == "y" else "restrains_local_y")      # buckling about y moves along z
        return L if free is None else free
    Iy, Iz, J = g(sec, "i_y"), g(sec, "i_z"), g(sec, "j")
    av = lambda k: (float(sec[k]) if sec.get(k) else BIG)                               # noqa: E731
    return dict(fy=g(mat, "yield_stress"), E=g(mat, "e_mod"), G=g(mat, "g_mod"),
                A=g(ec3, "a_eff", g(sec, "area")) if cls == 4 else g(sec, "area"), Iy=Iy, Iz=Iz, J=J, Iw=g(sec, "i_w", 0.0), L=L,
                Wy=W("y"), Wz=W("z"), Avy=av("a_sz"), Avz=av("a_sy"), Lcry=lcr("y"), Lcrz=lcr("z"),
                Llt=float(ms["ltb_length"]) if ms.get("ltb_length") is not None else (
                    L if unrestrained("restrains_torsion") is None else unrestrained("restrains_torsion")),
                ay=ALPHA.get(str(ec3.get("buckling_curve_y") or "B").upper(), 0.34),
                az=ALPHA.get(str(ec3.get("buckling_curve_z") or "B").upper(), 0.34),
                alt=ALPHA.get(str(ec3.get("buckling_curve_lt") or "B").upper(), 0.34),
                gm0=float(spec.get("gamma_m0") or 1.0), gm1=float(spec.get("gamma_m1") or 1.0), c1=float(spec.get("c1") or 0.0),
                zz=float(spec.get("c2") or 0.0) * float(spec.get("z_g") or 0.0) - float(spec.get("c3") or 0.0) * float(sec.get("wagner_coeff") or 0.0),
                on_b=1.0 if spec.get("include_buckling", True) else 0.0,
                on_lt=1.0 if (spec.get("include_ltb", True) and W("z") < BIG and Iz > Iy and J > 0.0) else 0.0)


def rules(c: dict, f: dict) -> dict:
    """Utilisation of every rule for the constants ``c`` and the forces ``f`` (Nmax, Nmin, Vy, Vz, My, Mz: envelope values of
    the local section forces along the entity, N tension-positive; Ma, Mb, Mc: M_z at the quarter points)."""
    nt, nc = max(f["Nmax"], 0.0), max(-f["Nmin"], 0.0)
    out = {"u_n": max(nt, nc) / (c["A"] * c["fy"] / c["gm0"]),
           "u_my": f["My"] / (c["Wy"] * c["fy"] / c["gm0"]), "u_mz": f["Mz"] / (c["Wz"] * c["fy"] / c["gm0"]),
           "u_vy": f["Vy"] / (c["Avy"] * c["fy"] / (math.sqrt(3.0) * c["gm0"])),
           "u_vz": f["Vz"] / (c["Avz"] * c["fy"] / (math.sqrt(3.0) * c["gm0"]))}
    out["u_nm"] = out["u_n"] + out["u_my"] + out["u_mz"]

    def chi(lam, alpha):                                       # EN 1993-1-1 (6.49) / (6.56)
        phi = 0.5 * (1.0 + alpha * (lam - 0.2) + lam * lam)
        return min(1.0, 1.0 / (phi + math.sqrt(max(phi * phi - lam * lam, 0.0))))
    for ax, inertia in (("y", c["Iy"]), ("z", c["Iz"])):
        ncr = math.pi ** 2 * c["E"] * inertia / c[f"Lcr{ax}"] ** 2
        out[f"chi{ax}"] = chi(math.sqrt(c["A"] * c["fy"] / ncr), c[f"a{ax}"])
        out[f"u_b{ax}"] = c["on_b"] * nc / (out[f"chi{ax}"] * c["A"] * c["fy"] / c["gm1"])
    c1 = c["c1"] if c["c1"] > 0.0 else \
        12.5 * (f["Mz"] + TINY) / (2.5 * (f["Mz"] + TINY) + 3 * abs(f["Ma"]) + 4 * abs(f["Mb"]) + 3 * abs(f["Mc"]))
    L = c["Llt"]
    # general form (NCCI SN003, k = k_w = 1): zz = C2 z_g - C3 z_j, z_g load height above the shear centre, z_j = wagner_coeff
    zz = c["zz"]
    ncr_minor = math.pi ** 2 * c["E"] * c["Iy"] / L ** 2
    mcr = c1 * ncr_minor * (math.sqrt(c["Iw"] / c["Iy"] + c["G"] * c["J"] / ncr_minor + zz * zz) - zz)
    out.update(C1=c1, Mcr=mcr, chilt=chi(math.sqrt(c["Wz"] * c["fy"] / mcr), c["alt"]))
    out["u_ltb"] = c["on_lt"] * f["Mz"] / (out["chilt"] * c["Wz"] * c["fy"] / c["gm1"])
    out["u_int"] = c["on_b"] * (max(out["u_by"], out["u_bz"]) + max(out["u_ltb"], out["u_mz"] * c["gm1"] / c["gm0"])
                               + out["u_my"] * c["gm1"] / c["gm0"])
    out["gov"] = max(out[n] for n in RULES)
    return out
