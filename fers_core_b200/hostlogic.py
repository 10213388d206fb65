"""Host-side analysis bookkeeping shared by the CUDA engine and the test oracle.

Nothing here is arithmetic of the hot path: it decides WHICH right-hand sides exist
(load cases + combinations, pydantic_models.py:2171-2174), which geometry each one is solved
on (imperfection cases, pydantic_models.py:2036-2040) and a few scalar conventions.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field

import numpy as np

from . import wire

# stiffness multiplier of a tension-/compression-only member that is switched off by the
# active-set loop (keeps nodes that hang only on inactive members non-singular)
INACTIVE_SCALE = 1.0e-9
# rigid members (RigidStrategy "RigidMember", pydantic_models.py:1047-1049) are modelled as
# members RIGID_FACTOR times stiffer than the stiffest section of the model
RIGID_FACTOR = 1.0e6


@dataclass
class RhsGroup:
    imperfection: dict | None        # raw imperfection case or None (perfect geometry)
    columns: list                    # rhs column indices solved on this geometry


@dataclass
class RhsPlan:
    kinds: list                      # "Loadcase" | "Loadcombination" per column
    ids: list                        # wire id per column
    names: list                      # result key per column
    factors: np.ndarray              # [n_rhs, Nlc] load-case factors per column
    groups: list = field(default_factory=list)

    @property
    def n_rhs(self) -> int:
        return len(self.kinds)


def rhs_plan(model: wire.ModelArrays) -> RhsPlan:
    """Columns = load cases (when ``solve_loadcases``) followed by combinations, wire order."""
    nlc = len(model.load_cases)
    kinds, ids, names, rows = [], [], [], []
    if model.options.solve_loadcases:
        for i, lc in enumerate(model.load_cases):
            kinds.append("Loadcase"); ids.append(lc.id); names.append(lc.name)
            r = np.zeros(nlc); r[i] = 1.0
            rows.append(r)
    for ci in range(model.combo_ids.shape[0]):
        kinds.append("Loadcombination"); ids.append(int(model.combo_ids[ci])); names.append(model.combo_names[ci])
        rows.append(model.combo_factors[ci].copy())
    factors = np.array(rows, np.float64).reshape(len(rows), nlc)
    plan = RhsPlan(kinds, ids, names, factors)
    # group columns by imperfection case (first case listing a combination wins)
    imp_of = {}
    for case in model.imperfections:
        for cid in case.get("load_combinations") or []:
            cid = cid["id"] if isinstance(cid, dict) else cid
            imp_of.setdefault(int(cid), case)
    by_case: dict = {}
    order: list = []
    for col in range(plan.n_rhs):
        case = imp_of.get(ids[col]) if kinds[col] == "Loadcombination" else None
        key = id(case) if case is not None else None
        if key not in by_case:
            by_case[key] = RhsGroup(case, [])
            order.append(key)
        by_case[key].columns.append(col)
    plan.groups = [by_case[k] for k in order]
    return plan


def _unit(v, default):
    a = np.asarray(default if v is None else v, np.float64)
    n = np.linalg.norm(a)
    return a / n if n > 0 else np.asarray(default, np.float64)


def imperfect_coordinates(model: wire.ModelArrays, case: dict) -> np.ndarray:
    """Node coordinates perturbed by one imperfection case.

    Sway (pydantic_models.py:1970-1985): lean angle ``magnitude`` [rad] about ``axis``; a node's
    offset is magnitude * h * (axis x height_direction) with h its height above
    ``reference_point`` along ``height_direction``.  Translation (1988-1991): nodes of the listed
    member sets are shifted by magnitude * axis.
    """
    xyz = model.xyz.copy()
    for sw in case.get("sway_imperfections") or []:
        hd = _unit(sw.get("height_direction"), (0.0, 1.0, 0.0))
        ax = _unit(sw.get("axis"), (0.0, 0.0, 1.0))
        ref = np.asarray(sw.get("reference_point") or (0.0, 0.0, 0.0), np.float64)
        sway_dir = np.cross(ax, hd)
        n = np.linalg.norm(sway_dir)
        if n == 0.0:
            continue
        sway_dir /= n
        h = (model.xyz - ref) @ hd
        xyz += float(sw["magnitude"]) * h[:, None] * sway_dir[None, :]
    for tr in case.get("translation_imperfections") or []:
        sets = tr.get("memberset_ids") or tr.get("memberset") or []
        ax = np.asarray(tr.get("axis") or (0.0, 0.0, 0.0), np.float64)
        nodes = set()
        for sid in sets:
            sid = sid["id"] if isinstance(sid, dict) else sid
            for mi in model.member_sets.get(int(sid), []):
                nodes.add(int(model.m_n1[mi])); nodes.add(int(model.m_n2[mi]))
        if nodes:
            idx = np.fromiter(nodes, np.int64)
            xyz[idx] += float(tr["magnitude"]) * ax[None, :]
    return xyz


def rigid_property_row(prop: np.ndarray) -> np.ndarray:
    """Property row [EA EIy EIz GJ GAsy GAsz ..] used for MemberType.Rigid members."""
    row = np.zeros(wire.NPROP)
    if prop.size:
        ea = float(prop[:, wire.P_EA].max())
        ei = float(max(prop[:, wire.P_EIY].max(), prop[:, wire.P_EIZ].max()))
    else:
        ea, ei = 2.1e11 * 1e-2, 2.1e11 * 1e-4
    row[wire.P_EA] = RIGID_FACTOR * ea
    row[wire.P_EIY] = row[wire.P_EIZ] = row[wire.P_GJ] = RIGID_FACTOR * ei
    return row


def member_frames_host(model: wire.ModelArrays, members: np.ndarray | None = None):
    """Local axes R [n,3,3] (rows x, y, z in global coordinates), axis lengths L [n] and end offsets [n,2,3] of the
    given member rows (all by default) on the host: the rule of fers_core/members/member.py:309-354 over the elastic
    axis (end offsets included).  Host-side post-processing only (deflected-shape stations); the kernels form their
    own frames."""
    members = np.arange(model.n_members) if members is None else np.asarray(members)
    xyz, m_rel = model.xyz, model.m_rel[members]
    off = np.zeros((len(members), 2, 3))
    rows = np.nonzero(m_rel >= 0)[0]
    if len(rows):
        off[rows] = model.rel[m_rel[rows], wire.REL_O1:wire.REL_O1 + 6].reshape(-1, 2, 3)
    d = xyz[model.m_n2[members]] - xyz[model.m_n1[members]] + off[:, 1] - off[:, 0]
    L = np.linalg.norm(d, axis=1)
    lx = d / L[:, None]
    ref = np.where((np.abs(lx[:, 1]) > 1.0 - 1e-6)[:, None], np.array([0.0, 0.0, 1.0]), np.array([0.0, 1.0, 0.0]))
    lz = np.cross(lx, ref)
    lz /= np.linalg.norm(lz, axis=1)[:, None]
    ly = np.cross(lz, lx)
    ly /= np.linalg.norm(ly, axis=1)[:, None]
    c, s = np.cos(model.m_roll[members])[:, None], np.sin(model.m_roll[members])[:, None]
    return np.stack([lx, c * ly + s * lz, -s * ly + c * lz], axis=1), L, off


DEFLECTED_STATIONS = 11      # x_frac = 0, 0.1 .. 1 (the count is the closed engine's choice; stated here, unpinned)


def member_deflected_shape(model: wire.ModelArrays, u_nodes: np.ndarray, members: np.ndarray | None = None) -> np.ndarray:
    """MemberResult.member_displacements (pydantic_models.py:1622-1625): global displacement [n, stations, 3] at
    evenly spaced stations of the given member rows (all by default), "the end-DOF homogeneous cubic": Hermite
    interpolation of the end translations and rotations in the member's local axes, the reconstruction of
    fers_core/fers/deformation_utils.py:46-114 (pinned on it: tests/golden/hermite_cases.json); linear for truss
    members, whose ends are pinned.  ``u_nodes`` [Nn,6] are the global node displacements of one result column."""
    members = np.arange(model.n_members) if members is None else np.asarray(members)
    nm = len(members)
    xi = np.linspace(0.0, 1.0, DEFLECTED_STATIONS)
    if nm == 0:
        return np.zeros((0, len(xi), 3))
    R, L, off = member_frames_host(model, members)
    ends = []
    for e, nodes in enumerate((model.m_n1[members], model.m_n2[members])):
        t, r = u_nodes[nodes, 0:3], u_nodes[nodes, 3:6]
        t = t + np.cross(r, off[:, e])                       # end of the elastic axis: rigid arm from the node
        ends.append((np.einsum("mij,mj->mi", R, t), np.einsum("mij,mj->mi", R, r)))
    (t1, r1), (t2, r2) = ends
    x2, x3 = xi * xi, xi * xi * xi
    n1, n2, n3, n4 = 1 - 3 * x2 + 2 * x3, xi - 2 * x2 + x3, 3 * x2 - 2 * x3, -x2 + x3
    Lc = L[:, None]
    loc = np.empty((nm, len(xi), 3))
    loc[:, :, 0] = t1[:, 0:1] * (1 - xi) + t2[:, 0:1] * xi
    loc[:, :, 1] = t1[:, 1:2] * n1 + Lc * r1[:, 2:3] * n2 + t2[:, 1:2] * n3 + Lc * r2[:, 2:3] * n4
    loc[:, :, 2] = t1[:, 2:3] * n1 - Lc * r1[:, 1:2] * n2 + t2[:, 2:3] * n3 - Lc * r2[:, 1:2] * n4
    truss = model.m_type[members] == wire.MT_TRUSS
    if truss.any():
        loc[truss] = t1[truss][:, None, :] * (1 - xi)[None, :, None] + t2[truss][:, None, :] * xi[None, :, None]
    return np.einsum("mji,msj->msi", R, loc)


# ---- RigidStrategy.LinearMpc (pydantic_models.py:1047-1049): rigid members as exact kinematic constraints -------------------
@dataclass
class MpcInfo:
    """What engine.analyze needs after solving the transformed model (mpc_transform)."""
    slaves: np.ndarray          # [ns] node indices carried by a master
    master: np.ndarray          # [ns] root node of each slave's rigid cluster
    arm: np.ndarray             # [ns,3] x_slave - x_master
    links: list                 # rigid members far node first: (member row, near node, far node), deepest first
    att_member: np.ndarray      # [na] elastic member ends that sat on a slave node: member row ...
    att_end: np.ndarray         # [na] ... 0 start / 1 end ...
    att_slave: np.ndarray       # [na] ... index into `slaves`
    ext: np.ndarray             # [ns,6,Nlc] nodal loads that sat on the slave nodes, per load case


def mpc_transform(model: wire.ModelArrays):
    """(transformed model, MpcInfo), or (None, reason) when the rigid members stay stiff members.

    Every node of a cluster of rigid members rides on the cluster's root (its supported node, else its lowest index):
    u_s = u_m + theta_m x (x_s - x_m), theta_s = theta_m.  In the transformed model the elastic members that ended on a slave
    end on the master instead, with the arm added to their end offset (the rigid arms of Member.start_offset / end_offset do
    exactly this kinematics, csrc/fers_elem.cuh beam_arms); nodal loads on slaves move to the master with their moment
    r x F; slave nodes are held (their displacements follow from the master afterwards) and the rigid members get a
    zero-stiffness section.  Nothing is approximated: the matrix that is factorised is T^T K0 T written member by member."""
    opts = model.options
    rigid = np.nonzero(model.m_type == wire.MT_RIGID)[0]
    if opts.rigid_strategy != "LinearMpc" or not len(rigid):
        return None, "not requested"
    if opts.order != "LINEAR":
        return None, "second-order analysis"
    if model.imperfections or len(model.support_curves) or len(model.hinge_curves):
        return None, "imperfection cases or load-dependent springs"
    if np.isin(model.bc_code, (wire.BC_POS_ONLY, wire.BC_NEG_ONLY)).any() or \
            np.isin(model.m_type, (wire.MT_TENSION, wire.MT_COMPRESSION, wire.MT_CABLE)).any() or len(model.c_member):
        return None, "unilateral supports or members, cables"
    nn = model.n_nodes
    adj: dict = {}
    for m in rigid:
        a, b = int(model.m_n1[m]), int(model.m_n2[m])
        adj.setdefault(a, []).append((b, int(m))); adj.setdefault(b, []).append((a, int(m)))
    loaded = set()
    for lc in model.load_cases:
        loaded.update(np.asarray(lc.dl_member).tolist()); loaded.update(np.asarray(lc.pl_member).tolist())
    if loaded & set(rigid.tolist()):
        return None, "a rigid member carries a member load"
    master = np.arange(nn)
    depth = np.zeros(nn, np.int64)
    links = []
    seen: set = set()
    for start in sorted(adj):
        if start in seen:
            continue
        comp, stack = [], [start]
        seen.add(start)
        while stack:
            v = stack.pop(); comp.append(v)
            for w, _ in adj[v]:
                if w not in seen:
                    seen.add(w); stack.append(w)
        supported = [v for v in sorted(comp) if model.node_support_id[v] >= 0]
        if len(supported) > 1:
            return None, "a rigid cluster with more than one supported node"
        root = supported[0] if supported else min(comp)
        done, stack, n_links = {root}, [root], 0
        while stack:
            v = stack.pop()
            for w, m in adj[v]:
                if w not in done:
                    done.add(w); depth[w] = depth[v] + 1; links.append((m, v, w)); stack.append(w); n_links += 1
        if n_links != sum(len(adj[v]) for v in comp) // 2:
            return None, "rigid members form a loop"
        for v in comp:
            master[v] = root
    slaves = np.nonzero(master != np.arange(nn))[0]
    if (model.bc_code[slaves] != wire.BC_FREE).any():
        return None, "a support on a slave node"
    is_slave = np.zeros(nn, bool); is_slave[slaves] = True
    if model.n_plates and is_slave[model.p_nodes].any():
        return None, "a plate on a slave node"
    if len(model.w_member) and (is_slave[model.m_n1[model.w_member]].any() or is_slave[model.m_n2[model.w_member]].any()):
        return None, "a warping member on a slave node"
    elastic = model.m_type != wire.MT_RIGID
    if (elastic & (master[model.m_n1] == master[model.m_n2])).any():
        return None, "an elastic member with both ends in one rigid cluster"
    slave_pos = {int(s): i for i, s in enumerate(slaves)}
    arm = model.xyz[slaves] - model.xyz[master[slaves]]
    # ---- the transformed model
    m_n1, m_n2, m_rel = model.m_n1.copy(), model.m_n2.copy(), model.m_rel.copy()
    rel = [r for r in model.rel]
    att_m, att_e, att_s = [], [], []
    for end, nodes in ((0, m_n1), (1, m_n2)):
        for m in np.nonzero(elastic & is_slave[nodes])[0]:
            s_i = slave_pos[int(nodes[m])]
            if m_rel[m] < 0:
                row = np.zeros(wire.REL_STRIDE); row[:12] = wire.RIGID_RELEASE
            else:
                row = rel[m_rel[m]].copy()                      # records may be shared between members: copy on write
            m_rel[m] = len(rel)
            o = wire.REL_O1 if end == 0 else wire.REL_O2
            row[o:o + 3] += arm[s_i]
            rel.append(row)
            nodes[m] = master[nodes[m]]
            att_m.append(int(m)); att_e.append(end); att_s.append(s_i)
    prop = np.vstack([model.prop, np.zeros((1, model.prop.shape[1]))]) if model.prop.size else np.zeros((1, wire.NPROP))
    m_prop, m_type = model.m_prop.copy(), model.m_type.copy()
    m_prop[rigid] = prop.shape[0] - 1
    m_type[rigid] = wire.MT_NORMAL
    bc_code = model.bc_code.copy()
    bc_code[slaves] = wire.BC_FIXED
    nlc = len(model.load_cases)
    ext = np.zeros((len(slaves), 6, max(nlc, 1)))
    cases = []
    for li, lc in enumerate(model.load_cases):
        node, vec = np.asarray(lc.nod_node).copy(), np.asarray(lc.nod_vec, np.float64).copy()
        for k in np.nonzero(is_slave[node])[0] if len(node) else []:
            s_i = slave_pos[int(node[k])]
            ext[s_i, :, li] += vec[k]
            vec[k, 3:] += np.cross(arm[s_i], vec[k, :3])
            node[k] = master[node[k]]
        cases.append(dataclasses.replace(lc, nod_node=node, nod_vec=vec))
    rho = model.prop_rho_a
    if rho is not None and len(rho) == model.prop.shape[0]:
        rho = np.concatenate([rho, [0.0]])
    model_t = dataclasses.replace(model, m_n1=m_n1, m_n2=m_n2, m_rel=m_rel, rel=np.array(rel).reshape(len(rel), wire.REL_STRIDE),
                                  prop=prop, prop_rho_a=rho, m_prop=m_prop, m_type=m_type, bc_code=bc_code, load_cases=cases)
    links.sort(key=lambda t: -depth[t[2]])
    return model_t, MpcInfo(slaves, master[slaves], arm, links, np.array(att_m, np.int64), np.array(att_e, np.int64),
                            np.array(att_s, np.int64), ext)


def pdelta_suppressed_axes(model: wire.ModelArrays) -> tuple:
    """Global translational axes (0,1,2) excluded from P-Delta (pydantic_models.py:1374-1384)."""
    o = model.options
    if o.pdelta_mode == "IN_PLANE_ONLY" and model.n_nodes:
        ext = np.ptp(model.xyz, axis=0)
        return (int(np.argmin(ext)),)
    return tuple("XYZ".index(a) for a in o.pdelta_suppress_axes if a in ("X", "Y", "Z"))


# ---- unilateral supports (SupportConditionType PositiveOnly / NegativeOnly, fers_core/supports/supportcondition.py:25-32:
#      "unilateral (contact-like) ... a spring active only in the allowed direction; if omitted ... an ideal unilateral
#      constraint").  Convention used here (the engine that defines it is closed): a PositiveOnly support can only push
#      the structure in the + direction of its DOF (reaction >= 0), i.e. it resists negative displacements and lets the
#      node lift off in the + direction; NegativeOnly is the mirror image.
def unilateral_dofs(model: wire.ModelArrays):
    """(flat DOF indices, sign +1 / -1, spring stiffness or 0 for an ideal constraint)."""
    code = model.bc_code.reshape(-1)
    idx = np.nonzero(np.isin(code, (wire.BC_POS_ONLY, wire.BC_NEG_ONLY)))[0]
    sign = np.where(code[idx] == wire.BC_POS_ONLY, 1.0, -1.0)
    return idx, sign, model.bc_k.reshape(-1)[idx].copy()


# ---- load-dependent support springs (SupportStiffnessCurve, pydantic_models.py:1249-1265; supportcondition.py:19-23:
#      "the curve overrides the constant value and the solver re-evaluates the stiffness each iteration based on the
#      specified ... component").  The driver is the global reaction component `depends_on` AT THE SAME NODE (force of
#      the support on the structure, K u - F), its absolute value unless `signed`; between the points the stiffness is
#      interpolated linearly, outside it keeps the end values (fers_core/examples/131_...py:126-139 `interpolate_k`).
def curve_stiffness(model: wire.ModelArrays, r: np.ndarray | None) -> np.ndarray:
    """Stiffness of every curve-driven support DOF for the reactions r [Ndof] (None: zero force)."""
    k = np.zeros(len(model.support_curves))
    for j, (_dof, drv, pts, signed) in enumerate(model.support_curves):
        f = 0.0 if r is None else float(r.reshape(-1)[drv])
        k[j] = np.interp(f if signed else abs(f), pts[:, 0], pts[:, 1])
    return k


def curves_converged(k_new: np.ndarray, k_old: np.ndarray, tol: float) -> bool:
    return bool(np.all(np.abs(k_new - k_old) <= max(tol, 1e-12) * np.abs(k_old)))


# ---- load-dependent hinge springs (MemberHinge.stiffness_curve_*; MemberStiffnessCurve pydantic_models.py:1638-1652:
#      "which member-local force component the stiffness depends on", |force| unless `signed`).  The member-side twin
#      of the support curves, same interpolation and the same outer iteration.  The driver is the member's own local
#      internal force at the END THE HINGE SITS ON, positive-face convention: what the member beyond the cut exerts on
#      the +x face, i.e. +local_end_forces at the end and -local_start_forces at the start (stiffness-sign end forces,
#      SURVEY A.6), so N is tension-positive ("Axial force (tension-positive)", pydantic_models.py:644) and every
#      component has one sign along an unloaded member.  Only `signed` curves see the sign.  The engine that defines
#      the convention is closed: stated, unpinned.
def hinge_curve_stiffness(model: wire.ModelArrays, f_local: np.ndarray | None) -> np.ndarray:
    """Stiffness of every curve-driven hinge spring for the local member end forces f_local [Nm,12] (None: zero force)."""
    k = np.zeros(len(model.hinge_curves))
    for j, (m, _row, slot, comp, pts, signed) in enumerate(model.hinge_curves):
        f = 0.0
        if f_local is not None:
            f = -float(f_local[m, comp]) if slot < 6 else float(f_local[m, 6 + comp])
        k[j] = np.interp(f if signed else abs(f), pts[:, 0], pts[:, 1])
    return k


def rel_with_hinge_curves(model: wire.ModelArrays, k: np.ndarray) -> np.ndarray:
    """Member modifier records with the curve-driven release springs at the stiffness ``k`` of the current iteration."""
    rel = model.rel.copy()
    for (_m, row, slot, _comp, _pts, _signed), kj in zip(model.hinge_curves, k):
        rel[row, slot] = kj
    return rel


def effective_bc(model: wire.ModelArrays, active: np.ndarray, curve_k: np.ndarray | None = None):
    """bc_code / bc_k with every unilateral DOF replaced by Fixed (ideal) or Spring (with stiffness) when it is active
    and by Free when it is not; curve-driven springs take the stiffness ``curve_k`` of the current iteration."""
    idx, _, k = unilateral_dofs(model)
    code = model.bc_code.reshape(-1).copy()
    kk = model.bc_k.reshape(-1).copy()
    if curve_k is not None and len(model.support_curves):
        kk[[c[0] for c in model.support_curves]] = curve_k
    code[idx] = np.where(active, np.where(k > 0.0, wire.BC_SPRING, wire.BC_FIXED), wire.BC_FREE)
    kk[idx] = np.where(active & (k > 0.0), k, 0.0)
    return code.reshape(model.bc_code.shape), kk.reshape(model.bc_k.shape)


def unilateral_update(model: wire.ModelArrays, active: np.ndarray, u: np.ndarray, r: np.ndarray) -> np.ndarray:
    """Next active set from the displacements u and reactions r (force of the support on the structure) at all DOFs:
    an active support stays while its reaction has the allowed sign, an inactive one closes when the node has moved
    against it."""
    idx, sign, k = unilateral_dofs(model)
    uu, rr = u.reshape(-1)[idx] * sign, r.reshape(-1)[idx] * sign
    tol_r = 1e-9 * max(float(np.abs(r).max()) if r.size else 0.0, 1e-300)
    keep = np.where(k > 0.0, uu <= 0.0, rr >= -tol_r)
    close = uu < 0.0
    return np.where(active, keep, close)


# ---- hinge capacities (MemberHinge.max_tension_v* / max_moment_m*, pydantic_models.py:2069-2074) -------------------------------
HINGE_FORCE_NAMES = ("N", "Vy", "Vz", "Mx", "My", "Mz")


def hinge_capacity_notes(model: wire.ModelArrays, f_local: np.ndarray, names: list) -> list:
    """The schema gives the hinge capacities no behaviour in the analysis, so they are not enforced; what the analysis can
    say is where they are exceeded.  For every capacity the largest demand over the result columns (f_local [ncols, Nm, 12],
    member end forces in the stiffness sign): the axial one counts TENSION only ("max_tension_vx": minus the end force at the
    start, plus at the end), the others by magnitude.  Returns W_HINGE_CAPACITY_EXCEEDED notes (code, message)."""
    notes = []
    for m, end, comp, cap, hid in model.hinge_caps:
        v = f_local[:, m, 6 * end + comp]
        demand = (-v if end == 0 else v) if comp == 0 else np.abs(v)
        col = int(np.argmax(demand))
        if demand[col] > cap * (1.0 + 1e-12):
            notes.append(("W_HINGE_CAPACITY_EXCEEDED",
                          f"member {int(model.m_ids[m])} {'start' if end == 0 else 'end'} (hinge {hid}): {HINGE_FORCE_NAMES[comp]} = "
                          f"{demand[col]:.6g} exceeds the capacity {cap:.6g} (utilisation {demand[col] / cap:.3f}, {names[col]}); "
                          "capacities are not enforced in the analysis"))
    return notes
