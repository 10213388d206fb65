"""CPU ORACLE -- host-side rules of the analysis, restated for the oracle alone.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The product keeps these rules in fers_core_b200/hostlogic.py; until round 2 the oracle imported them from there, so a bug in
them was invisible to every GPU <-> oracle comparison ("What's weak" 2).  This file states the same rules a second
time, written from the reference's schema text and examples (file:line of /root/reference below), with plain loops instead of
the product's vectorised forms; tests/test_host_oracle.py holds the two against each other on every golden model.  Only the
wire parser (the data container ``wire.ModelArrays``) and the result container stay shared.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# wire codes of the support conditions (fers_core/supports/supportcondition.py:11-32)
BC_FREE, BC_FIXED, BC_SPRING, BC_POS_ONLY, BC_NEG_ONLY = 0, 1, 2, 3, 4
RIGID_STIFFNESS_RATIO = 1.0e6          # a Rigid member is a member this many times stiffer than the stiffest section


@dataclass
class Group:
    imperfection: dict | None
    columns: list


@dataclass
class Plan:
    kinds: list
    ids: list
    names: list
    factors: np.ndarray
    groups: list = field(default_factory=list)

    @property
    def n_rhs(self) -> int:
        return len(self.kinds)


def rhs_plan(model) -> Plan:
    """Result columns: every load case on its own when ``solve_loadcases`` (pydantic_models.py:1389-1391: ResultsBundle
    carries ``loadcases`` and ``loadcombinations``, 2171-2185), then every combination with its load-case factors
    (LoadCombination.load_case_factors, 491-505).  A combination listed by an imperfection case (ImperfectionCase
    .load_combinations, 1993-2003) is solved on that case's geometry; the first case that lists it wins."""
    n_lc = len(model.load_cases)
    kinds, ids, names, rows = [], [], [], []
    if model.options.solve_loadcases:
        for i in range(n_lc):
            row = [0.0] * n_lc
            row[i] = 1.0
            kinds.append("Loadcase"); ids.append(model.load_cases[i].id); names.append(model.load_cases[i].name); rows.append(row)
    for c in range(len(model.combo_ids)):
        kinds.append("Loadcombination"); ids.append(int(model.combo_ids[c])); names.append(model.combo_names[c])
        rows.append([float(v) for v in model.combo_factors[c]])
    plan = Plan(kinds, ids, names, np.asarray(rows, float).reshape(len(rows), n_lc))
    case_of_combo = {}
    for case in model.imperfections:
        for entry in case.get("load_combinations") or []:
            cid = int(entry["id"]) if isinstance(entry, dict) else int(entry)
            if cid not in case_of_combo:
                case_of_combo[cid] = case
    groups: list = []
    for col in range(plan.n_rhs):
        case = case_of_combo.get(ids[col]) if kinds[col] == "Loadcombination" else None
        for g in groups:
            if g.imperfection is case:
                g.columns.append(col)
                break
        else:
            groups.append(Group(case, [col]))
    plan.groups = groups
    return plan


def _normalised(v, fallback):
    v = list(fallback) if v is None else [float(x) for x in v]
    n = math.sqrt(sum(x * x for x in v))
    return [x / n for x in v] if n > 0.0 else list(fallback)


def imperfect_coordinates(model, case: dict) -> np.ndarray:
    """Geometry of one imperfection case.  Sway (SwayImperfection, pydantic_models.py:1970-1985): "lean angle in radians"
    about ``axis``; a node at height h above ``reference_point`` (measured along ``height_direction``) moves by
    magnitude * h in the direction axis x height_direction.  Translation (TranslationImperfection, 1988-1991): the nodes of
    the listed member sets move by magnitude * axis (once per node, however many members share it)."""
    xyz = np.array(model.xyz, float, copy=True)
    n_nodes = xyz.shape[0]
    for sw in case.get("sway_imperfections") or []:
        up = _normalised(sw.get("height_direction"), (0.0, 1.0, 0.0))
        ax = _normalised(sw.get("axis"), (0.0, 0.0, 1.0))
        ref = sw.get("reference_point") or (0.0, 0.0, 0.0)
        d = [ax[1] * up[2] - ax[2] * up[1], ax[2] * up[0] - ax[0] * up[2], ax[0] * up[1] - ax[1] * up[0]]
        dn = math.sqrt(d[0] ** 2 + d[1] ** 2 + d[2] ** 2)
        if dn == 0.0:
            continue
        phi = float(sw["magnitude"])
        for i in range(n_nodes):
            h = sum((float(model.xyz[i, c]) - float(ref[c])) * up[c] for c in range(3))
            for c in range(3):
                xyz[i, c] += phi * h * d[c] / dn
    for tr in case.get("translation_imperfections") or []:
        ax = tr.get("axis") or (0.0, 0.0, 0.0)
        moved = set()
        for entry in tr.get("memberset_ids") or tr.get("memberset") or []:
            sid = int(entry["id"]) if isinstance(entry, dict) else int(entry)
            for m in model.member_sets.get(sid, []):
                for node in (int(model.m_n1[m]), int(model.m_n2[m])):
                    if node not in moved:
                        moved.add(node)
                        for c in range(3):
                            xyz[node, c] += float(tr["magnitude"]) * float(ax[c])
    return xyz


def rigid_property_row(prop: np.ndarray) -> np.ndarray:
    """[EA EIy EIz GJ GAsy GAsz . .] of a MemberType.Rigid member (pydantic_models.py:592-598; RigidStrategy.RigidMember,
    1047-1049): 1e6 x the largest axial and the largest bending stiffness of the model, no shear flexibility."""
    row = np.zeros(8)
    ea = max((float(p[0]) for p in prop), default=2.1e11 * 1e-2)
    ei = max((max(float(p[1]), float(p[2])) for p in prop), default=2.1e11 * 1e-4)
    row[0] = RIGID_STIFFNESS_RATIO * ea
    row[1] = row[2] = row[3] = RIGID_STIFFNESS_RATIO * ei
    return row


def pdelta_suppressed_axes(model) -> tuple:
    """Global axes without geometric stiffness (AnalysisOptions.pdelta_mode / pdelta_suppress_axes,
    pydantic_models.py:1374-1384): IN_PLANE_ONLY drops the axis of least extent of the model."""
    o = model.options
    if o.pdelta_mode == "IN_PLANE_ONLY" and model.n_nodes:
        ext = [float(model.xyz[:, c].max() - model.xyz[:, c].min()) for c in range(3)]
        return (ext.index(min(ext)),)
    return tuple("XYZ".index(a) for a in o.pdelta_suppress_axes if a in ("X", "Y", "Z"))


# ---- unilateral supports (supportcondition.py:25-32) ---------------------------------------------------------------------
def unilateral_dofs(model):
    code = model.bc_code.reshape(-1)
    k = model.bc_k.reshape(-1)
    idx = [d for d in range(code.shape[0]) if code[d] in (BC_POS_ONLY, BC_NEG_ONLY)]
    sign = [1.0 if code[d] == BC_POS_ONLY else -1.0 for d in idx]
    return np.asarray(idx, np.int64), np.asarray(sign, float), np.asarray([float(k[d]) for d in idx], float)


def effective_bc(model, active, curve_k=None):
    """Support codes / spring stiffness of one outer iteration: an active unilateral DOF is Fixed (ideal) or Spring (with its
    stiffness), an open one Free; load-dependent springs carry the stiffness of the current iteration."""
    code = model.bc_code.reshape(-1).copy()
    k = model.bc_k.reshape(-1).copy()
    if curve_k is not None:
        for (dof, _drv, _pts, _signed), kj in zip(model.support_curves, curve_k):
            k[dof] = kj
    idx, _sign, ku = unilateral_dofs(model)
    for j, d in enumerate(idx):
        if not active[j]:
            code[d], k[d] = BC_FREE, 0.0
        elif ku[j] > 0.0:
            code[d], k[d] = BC_SPRING, ku[j]
        else:
            code[d], k[d] = BC_FIXED, 0.0
    return code.reshape(model.bc_code.shape), k.reshape(model.bc_k.shape)


def unilateral_update(model, active, u, r):
    """Active set of the next outer iteration.  A PositiveOnly support pushes the structure in + only: an ideal one stays
    closed while its reaction (force of the support on the structure, K u - F) is >= 0, a sprung one while the node sits on
    it (u <= 0 in the support's direction); an open one closes when the node has moved against it."""
    idx, sign, k = unilateral_dofs(model)
    rmax = float(np.abs(r).max()) if np.size(r) else 0.0
    tol = 1e-9 * max(rmax, 1e-300)
    out = np.zeros(len(idx), bool)
    for j, d in enumerate(idx):
        uu, rr = float(u.reshape(-1)[d]) * sign[j], float(r.reshape(-1)[d]) * sign[j]
        if active[j]:
            out[j] = (uu <= 0.0) if k[j] > 0.0 else (rr >= -tol)
        else:
            out[j] = uu < 0.0
    return out


# ---- load-dependent springs (SupportStiffnessCurve 1249-1265, MemberStiffnessCurve 1638-1652; example 131:126-139) -------
def _interp(x: float, pts) -> float:
    """Piecewise linear through the points, constant beyond the ends (example 131 ``interpolate_k``)."""
    if x <= pts[0][0]:
        return float(pts[0][1])
    for (x0, y0), (x1, y1) in zip(pts[:-1], pts[1:]):
        if x <= x1:
            return float(y0 + (y1 - y0) * (x - x0) / (x1 - x0)) if x1 > x0 else float(y1)
    return float(pts[-1][1])


def curve_stiffness(model, r) -> np.ndarray:
    """Support springs driven by the reaction component ``depends_on`` at their own node (|force| unless signed)."""
    out = np.zeros(len(model.support_curves))
    for j, (_dof, drv, pts, signed) in enumerate(model.support_curves):
        f = 0.0 if r is None else float(np.asarray(r).reshape(-1)[drv])
        out[j] = _interp(f if signed else abs(f), [tuple(p) for p in pts])
    return out


def hinge_curve_stiffness(model, f_local) -> np.ndarray:
    """Hinge springs driven by the member's own internal force at the hinge end, tension-positive face convention: minus
    the stiffness-sign end force at the start, plus at the end."""
    out = np.zeros(len(model.hinge_curves))
    for j, (m, _row, slot, comp, pts, signed) in enumerate(model.hinge_curves):
        f = 0.0
        if f_local is not None:
            f = -float(f_local[m, comp]) if slot < 6 else float(f_local[m, 6 + comp])
        out[j] = _interp(f if signed else abs(f), [tuple(p) for p in pts])
    return out


def rel_with_hinge_curves(model, k) -> np.ndarray:
    rel = np.array(model.rel, float, copy=True)
    for (_m, row, slot, _comp, _pts, _signed), kj in zip(model.hinge_curves, k):
        rel[row, slot] = kj
    return rel


def curves_converged(k_new, k_old, tol: float) -> bool:
    return all(abs(a - b) <= max(tol, 1e-12) * abs(b) for a, b in zip(k_new, k_old))


# ---- hinge capacities (MemberHinge.max_tension_v* / max_moment_m*, pydantic_models.py:2069-2074) ----------------------------
def hinge_capacity_notes(model, f_local, names) -> list:
    """Capacities are not enforced (no behaviour stated in the schema): one note per capacity whose demand -- tension for the
    axial one, magnitude for shears and moments, largest over the result columns -- exceeds it."""
    label = ("N", "Vy", "Vz", "Mx", "My", "Mz")
    out = []
    for m, end, comp, cap, hid in model.hinge_caps:
        worst, at = -math.inf, 0
        for col in range(f_local.shape[0]):
            v = float(f_local[col, m, 6 * end + comp])
            d = ((-v) if end == 0 else v) if comp == 0 else abs(v)
            if d > worst:
                worst, at = d, col
        if worst > cap * (1.0 + 1e-12):
            out.append(("W_HINGE_CAPACITY_EXCEEDED",
                        f"member {int(model.m_ids[m])} {'start' if end == 0 else 'end'} (hinge {hid}): {label[comp]} = {worst:.6g} "
                        f"exceeds the capacity {cap:.6g} (utilisation {worst / cap:.3f}, {names[at]}); capacities are not enforced "
                        "in the analysis"))
    return out
