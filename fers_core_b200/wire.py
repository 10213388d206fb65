"""Wire-format parser: reference JSON document -> structure-of-arrays model.

The reference front end (``FERS.to_dict``, /root/reference/fers_core/fers/fers.py:315-361)
serialises its object graph to the document described by the generated schema
``fers_core/types/pydantic_models.py`` (``FERS`` 2241-2258).  This module is the host-side
half of the drop-in boundary: it turns that document into flat numpy arrays (the layout the
CUDA path consumes, see DESIGN.md "Data layout") and nothing else -- no arithmetic of the
analysis itself happens here.

Accepted spelling drifts: legacy enum values ``"Linear"``/
``"Nonlinear"``, ``"Rigid_Member"``/``"Linear_MPC"``; ``"Positive-only"``/``"Negative-only"``;
``memberset`` for ``memberset_ids``; unknown extra keys are ignored.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

import numpy as np

# member_type codes (pydantic_models.py:592-598)
MT_NORMAL, MT_TRUSS, MT_TENSION, MT_COMPRESSION, MT_RIGID, MT_CABLE = range(6)
_MEMBER_TYPES = {
    "normal": MT_NORMAL, "truss": MT_TRUSS, "tension": MT_TENSION,
    "compression": MT_COMPRESSION, "rigid": MT_RIGID, "cable": MT_CABLE,
}
# boundary-condition codes per DOF (pydantic_models.py:1232-1237)
BC_FREE, BC_FIXED, BC_SPRING, BC_POS_ONLY, BC_NEG_ONLY = range(5)
_BC_TYPES = {
    "free": BC_FREE, "fixed": BC_FIXED, "spring": BC_SPRING,
    "positiveonly": BC_POS_ONLY, "positive-only": BC_POS_ONLY,
    "negativeonly": BC_NEG_ONLY, "negative-only": BC_NEG_ONLY,
}
DOF_NAMES = ("X", "Y", "Z", "RX", "RY", "RZ")
_SUPPORT_FORCE = {"Fx": 0, "Fy": 1, "Fz": 2, "Mx": 3, "My": 4, "Mz": 5}      # SupportForceComponent, pydantic_models.py:1240-1246

# columns of the per-section property table (one row per distinct section x option set)
P_EA, P_EIY, P_EIZ, P_GJ, P_GASY, P_GASZ, P_WPL, P_SPARE = range(8)
NPROP = 8
NPDATA = 16         # doubles per plate record (csrc/fers_elem.cuh PD_*)
PD_THETA = 6        # in-plane angle of local x from the default plate frame (local_x_direction), radians
PD_ORTHO = 7        # != 0: orthotropic plate material in slots 8..13 (E_x, E_y, G_xy, nu_xy, G_xz, G_yz)
PD_EX, PD_EY, PD_GXY, PD_NUXY, PD_GXZ, PD_GYZ = 8, 9, 10, 11, 12, 13
PD_OFFSET = 14      # distance of the mid-surface from the nodes along the local plate normal (PlateElement.offset)
# local release order = local DOF order [u v w rx ry rz]
_HINGE_FIELDS = (
    "translational_release_vx", "translational_release_vy", "translational_release_vz",
    "rotational_release_mx", "rotational_release_my", "rotational_release_mz",
)
_HINGE_CURVES = ("stiffness_curve_vx", "stiffness_curve_vy", "stiffness_curve_vz",
                 "stiffness_curve_mx", "stiffness_curve_my", "stiffness_curve_mz")
_HINGE_CAPACITIES = ("max_tension_vx", "max_tension_vy", "max_tension_vz", "max_moment_mx", "max_moment_my",
                     "max_moment_mz", "max_bimoment_warp")
_MEMBER_FORCE = {"N": 0, "Vy": 1, "Vz": 2, "Mx": 3, "My": 4, "Mz": 5}       # MemberStiffnessCurveDriver, pydantic_models.py:583-589
RIGID_RELEASE = -1.0  # sentinel: DOF rigidly connected (no hinge spring)
# one row of ModelArrays.rel (csrc/fers_elem.cuh REL_*): 12 release springs, the shear-centre offset, the two end
# offsets (global X Y Z of start_offset, end_offset)
REL_YS, REL_ZS, REL_O1, REL_O2, REL_EIYZ, REL_STRIDE = 12, 13, 14, 17, 20, 21


class WireError(ValueError):
    """Malformed input document (maps to the reference's RuntimeError at fers.py:301-302)."""


@dataclass
class LoadCaseArrays:
    """One load case flattened to arrays (pydantic_models.py:2048-2059)."""
    id: int
    name: str
    # nodal forces+moments: (node index, 6-vector)
    nod_node: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    nod_vec: np.ndarray = field(default_factory=lambda: np.zeros((0, 6)))
    # distributed (trapezoid) member loads, global direction; row = [q1(3), q2(3), a, b]
    dl_member: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    dl_data: np.ndarray = field(default_factory=lambda: np.zeros((0, 8)))
    cab_q: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))     # [Ncable,3] uniform load per unstretched length on the cables
    # member point loads / moments; row = [F(3), M(3), position, is_local]
    pl_member: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    pl_data: np.ndarray = field(default_factory=lambda: np.zeros((0, 8)))
    # plate pressures resolved to elements: (plate index, global traction vector per area)
    pp_plate: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    pp_vec: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))


@dataclass
class Options:
    """analysis.options (pydantic_models.py:1334-1398; defaults anlysis_options.py:15-28)."""
    order: str = "NONLINEAR"            # LINEAR | NONLINEAR
    nonlinear_method: str = "COROTATIONAL"  # engine default (anlysis_options.py:53-55)
    pdelta_formulation: str = "CONSISTENT"
    pdelta_mode: str = "FULL"
    pdelta_suppress_axes: tuple = ()
    tolerance: float = 0.01
    max_iterations: int = 30
    dimensionality: str = "3D"
    rigid_strategy: str = "RigidMember"
    axial_slack: float = 500.0
    include_shear_deformation: bool = True
    include_warping: bool = True
    include_shear_center_coupling: bool = True
    include_member_deflected_shape: bool = False      # MemberResult.member_displacements (pydantic_models.py:1350-1353)
    solve_loadcases: bool = True
    enable_self_weight: bool = False
    gravity_direction: tuple = (0.0, -1.0, 0.0)
    gravity_factor: float = -9.81
    self_weight_load_case_id: int | None = None


@dataclass
class ModelArrays:
    """Structure-of-arrays view of one wire document."""
    node_ids: np.ndarray          # int64 [Nn]      wire ids, wire order
    xyz: np.ndarray               # f64   [Nn,3]
    node_support_id: np.ndarray   # int64 [Nn]      nodal_support id or -1
    bc_code: np.ndarray           # uint8 [Nn,6]    BC_* per DOF
    bc_k: np.ndarray              # f64   [Nn,6]    spring stiffness where BC_SPRING
    m_ids: np.ndarray             # int64 [Nm]
    m_n1: np.ndarray              # int32 [Nm]      start node INDEX
    m_n2: np.ndarray              # int32 [Nm]
    m_type: np.ndarray            # int32 [Nm]      MT_*
    m_roll: np.ndarray            # f64   [Nm]      roll angle, radians
    m_prop: np.ndarray            # int32 [Nm]      row of `prop` (-1: rigid member without section)
    m_rel: np.ndarray             # int32 [Nm]      row of `rel` or -1
    prop: np.ndarray              # f64   [Np,8]    EA EIy EIz GJ GAsy GAsz Wpl spare
    prop_rho_a: np.ndarray        # f64   [Np]      density*area (self weight)
    rel: np.ndarray               # f64   [Nr,REL_STRIDE] member modifier records: [0:12] hinge spring stiffness per
                                  #                 local end DOF (-1 rigid), [12] y_s, [13] z_s (shear-centre offset),
                                  #                 [14:17] start_offset, [17:20] end_offset (global axes), [20] E*I_yz
    # plates (TRI3): node indices, thickness, E, nu, behaviour code
    p_ids: np.ndarray             # int64 [Npl]
    p_nodes: np.ndarray           # int32 [Npl,3]
    p_data: np.ndarray            # f64   [Npl,8]  thickness, E, nu, modifiers (membrane, bending, shear), local-x angle, 1 spare
    p_flags: np.ndarray           # int32 [Npl]    bit0 membrane, bit1 bending, bit2 plane strain, bit3 Kirchhoff (DKT)
    load_cases: list              # [LoadCaseArrays]
    combo_ids: np.ndarray         # int64 [Nc]
    combo_names: list
    combo_factors: np.ndarray     # f64 [Nc,Nlc]
    imperfections: list           # raw imperfection_cases (resolved by analysis host code)
    member_sets: dict             # id -> [member index]
    options: Options
    warnings: list = field(default_factory=list)
    # load-dependent support springs (SupportStiffnessCurve, pydantic_models.py:1249-1265): one entry per DOF,
    # (flat DOF, flat DOF of the driving reaction component at the same node, points [n,2] sorted by force, signed)
    support_curves: list = field(default_factory=list)
    # load-dependent hinge springs (MemberHinge.stiffness_curve_*, MemberStiffnessCurve pydantic_models.py:1638-1652):
    # one entry per member end DOF, (member row, rel row, slot 0..11 of the release springs, driving local force
    # component 0..5 = N Vy Vz Mx My Mz at the same end, points [n,2] sorted by force, signed)
    hinge_curves: list = field(default_factory=list)
    # MemberHinge.max_tension_v* / max_moment_m* (pydantic_models.py:2069-2074): (member row, end 0 / 1, local force component
    # 0..5 = N Vy Vz Mx My Mz, capacity, hinge id).  Not enforced in the analysis (the schema states no behaviour); the
    # end forces are held against them afterwards (hostlogic.hinge_capacity_notes)
    hinge_caps: list = field(default_factory=list)
    # unity checks (pydantic_models.py:2202-2220): raw definitions + what their quantities read
    unity_checks: list = field(default_factory=list)
    member_sets_raw: list = field(default_factory=list)        # raw member_sets records (buckling lengths of the EC3 check)
    combo_limit_state: list = field(default_factory=list)      # per combination: "SLS" | "ULS" | ... | None
    m_classification: list = field(default_factory=list)       # per member: classification string or None
    m_section_raw: list = field(default_factory=list)          # per member: (section dict, material dict) or None
    m_ref_member: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))   # [Nm] reference member index or -1
    m_ref_node: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))     # [Nm] reference node index or -1
    # warping torsion (7th nodal DOF of the reference, pydantic_models.py:1384-1387): the warping DOF psi of a real
    # node lives in DOF 0 of an auxiliary node appended after the real ones (its other five DOFs are restrained)
    n_real_nodes: int = -1                                       # -1: no auxiliary nodes (n_nodes real nodes)
    warp_node: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))   # [n_real] aux node or -1
    w_member: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))    # [Nw] members with warping
    w_aux: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))  # [Nw,2] aux nodes of start / end
    # MemberType.Cable (pydantic_models.py:592-598): elastic catenary elements (fers_core_b200/cable.py); LoadCaseArrays.cab_q
    # carries their uniform member loads in the same order
    c_member: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))    # [Ncable] member rows
    c_l0: np.ndarray = field(default_factory=lambda: np.zeros(0))                  # [Ncable] unstretched length
    c_ea: np.ndarray = field(default_factory=lambda: np.zeros(0))                  # [Ncable] axial stiffness
    c_w: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))              # [Ncable,3] own weight per unstretched length
    p_surface: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))   # [Npl] PlateElement.source_surface_id or -1

    def member_raw(self, m: int):
        """(section dict, material dict) of member index m as they came over the wire."""
        if m < len(self.m_section_raw) and self.m_section_raw[m] is not None:
            return self.m_section_raw[m]
        return ({}, {})

    @property
    def n_nodes(self) -> int:
        """Nodes of the stiffness matrix: real nodes followed by the auxiliary warping nodes."""
        return int(self.node_ids.shape[0])

    @property
    def n_real(self) -> int:
        return self.n_nodes if self.n_real_nodes < 0 else int(self.n_real_nodes)

    @property
    def n_members(self) -> int:
        return int(self.m_ids.shape[0])

    @property
    def n_plates(self) -> int:
        return int(self.p_ids.shape[0])

    @property
    def n_dof(self) -> int:
        return 6 * self.n_nodes


def _norm_enum(v: Any) -> str:
    return str(v).strip().lower().replace("_", "").replace(" ", "")


def parse_options(raw: dict | None) -> Options:
    raw = raw or {}
    o = Options()
    order = _norm_enum(raw.get("order", "NONLINEAR"))
    o.order = "LINEAR" if order in ("linear", "firstorder") else "NONLINEAR"
    nm = raw.get("nonlinear_method")
    if nm is not None:
        o.nonlinear_method = "P_DELTA" if _norm_enum(nm) in ("pdelta",) else "COROTATIONAL"
    pf = raw.get("pdelta_formulation")
    if pf is not None:
        o.pdelta_formulation = "SIMPLIFIED" if _norm_enum(pf) == "simplified" else "CONSISTENT"
    pm = raw.get("pdelta_mode")
    if pm is not None:
        o.pdelta_mode = "IN_PLANE_ONLY" if _norm_enum(pm) == "inplaneonly" else "FULL"
    o.pdelta_suppress_axes = tuple(str(a).upper() for a in (raw.get("pdelta_suppress_axes") or ()))
    if raw.get("tolerance") is not None:
        o.tolerance = float(raw["tolerance"])
    if raw.get("max_iterations") is not None:
        o.max_iterations = int(raw["max_iterations"])
    dim = _norm_enum(raw.get("dimensionality", "3D"))
    o.dimensionality = "2D" if dim in ("2d", "twodimensional") else "3D"
    rs = _norm_enum(raw.get("rigid_strategy", "RigidMember"))
    o.rigid_strategy = "LinearMpc" if rs == "linearmpc" else "RigidMember"
    if raw.get("axial_slack") is not None:
        o.axial_slack = float(raw["axial_slack"])
    for k in ("include_shear_deformation", "include_warping", "include_shear_center_coupling",
              "include_member_deflected_shape", "solve_loadcases", "enable_self_weight"):
        if raw.get(k) is not None:
            setattr(o, k, bool(raw[k]))
    if raw.get("gravity_direction") is not None:
        o.gravity_direction = tuple(float(x) for x in raw["gravity_direction"])
    if raw.get("gravity_factor") is not None:
        o.gravity_factor = float(raw["gravity_factor"])
    if raw.get("self_weight_load_case_id") is not None:
        o.self_weight_load_case_id = int(raw["self_weight_load_case_id"])
    return o


def _vec3(v, what: str) -> np.ndarray:
    if isinstance(v, dict):
        v = [v.get("X", v.get("x", 0.0)), v.get("Y", v.get("y", 0.0)), v.get("Z", v.get("z", 0.0))]
    a = np.asarray(v, dtype=np.float64).reshape(-1)
    if a.shape[0] != 3:
        raise WireError(f"{what}: expected a 3-vector, got {v!r}")
    return a


def _index_map(ids: np.ndarray, what: str) -> dict:
    m = {int(v): i for i, v in enumerate(ids.tolist())}
    if len(m) != ids.shape[0]:
        raise WireError(f"duplicate {what} ids")
    return m


# ---- column records of the two big entity lists ------------------------------------------------------------------
# The wire document and the reference's object graph are read into the same column records (one list / array per field);
# everything after that is shared, so `from_fers(calc)` equals `parse_document(calc.to_dict())` by construction.
@dataclass
class NodeColumns:
    ids: np.ndarray               # int64 [Nn]
    xyz: np.ndarray               # f64 [Nn,3]
    support: list                 # per node: None | support id | inline support dict


@dataclass
class MemberColumns:
    ids: np.ndarray               # int64 [Nm]
    start: np.ndarray             # int64 [Nm] node ids
    end: np.ndarray
    mtype: list                   # member_type spellings
    rotation: np.ndarray          # f64 [Nm] degrees
    section: list                 # section id or None
    mirror: list
    start_hinge: list
    end_hinge: list
    start_offset: list            # {X, Y, Z} or None
    end_offset: list
    classification: list
    ref_member: list
    ref_node: list
    weight: list
    pretension: list = None        # Cable members (pydantic_models.py:1525-1528)
    unstretched_length: list = None  # Cable members (pydantic_models.py:1542-1545)


def node_columns(nodes: list) -> NodeColumns:
    nn = len(nodes)
    ids = np.fromiter((int(n["id"]) for n in nodes), np.int64, nn)
    xyz = np.array([[float(n["X"]), float(n["Y"]), float(n["Z"])] for n in nodes], np.float64).reshape(nn, 3)
    return NodeColumns(ids, xyz, [n.get("nodal_support") for n in nodes])


def member_columns(members: list) -> MemberColumns:
    nm = len(members)
    g = lambda k: [m.get(k) for m in members]          # noqa: E731
    return MemberColumns(
        ids=np.fromiter((int(m["id"]) for m in members), np.int64, nm),
        start=np.fromiter((int(m["start_node_id"]) for m in members), np.int64, nm),
        end=np.fromiter((int(m["end_node_id"]) for m in members), np.int64, nm),
        mtype=[m.get("member_type", "Normal") for m in members],
        rotation=np.fromiter((float(m.get("rotation_angle") or 0.0) for m in members), np.float64, nm),
        section=g("section"), mirror=g("mirror"), start_hinge=g("start_hinge"), end_hinge=g("end_hinge"),
        start_offset=g("start_offset"), end_offset=g("end_offset"), classification=g("classification"),
        ref_member=g("reference_member"), ref_node=g("reference_node"), weight=g("weight"),
        pretension=g("pretension"), unstretched_length=g("unstretched_length"))


def parse_document(doc: dict) -> ModelArrays:
    """Flatten a wire document (dict) into :class:`ModelArrays`."""
    if not isinstance(doc, dict) or "model" not in doc or "analysis" not in doc:
        raise WireError("document must contain 'model' and 'analysis'")
    model, analysis = doc["model"], doc["analysis"]
    return build_model_arrays(node_columns(model.get("nodes") or []), member_columns(model.get("members") or []), model, analysis)


def from_fers(calc) -> ModelArrays:
    """Input fast path: the reference's FERS object graph straight to :class:`ModelArrays`, without
    the ``to_dict`` -> pydantic gate -> ``json.dumps`` -> ``json.loads`` round trip of ``FERS.run_analysis``
    (fers_core/fers/fers.py:275-313; ~8 s at configs[3]).  Entities are collected the way ``FERS.to_dict`` collects them
    (fers.py:315-361: ``get_all_nodes`` / ``get_all_members`` order, fers.py:899-973, defines the DOF numbering); the
    two big lists (nodes, members) and the two big load lists (nodal, distributed) are read attribute by attribute into
    the column records the wire parser produces, the small tables go through their own ``to_dict``.  Everything after the
    columns is the code ``parse_document`` runs, so the result equals ``parse_document(calc.to_dict())`` field for field
    (tests/test_from_fers.py)."""
    members = calc.get_all_members()
    nodes = calc.get_all_nodes()
    nn, nm = len(nodes), len(members)
    ncols = NodeColumns(ids=np.fromiter((int(n.id) for n in nodes), np.int64, nn),
                        xyz=np.array([[float(n.X), float(n.Y), float(n.Z)] for n in nodes], np.float64).reshape(nn, 3),
                        support=[n.nodal_support.id if n.nodal_support else None for n in nodes])
    oid = lambda o: o.id if o else None          # noqa: E731
    mcols = MemberColumns(
        ids=np.fromiter((int(m.id) for m in members), np.int64, nm),
        start=np.fromiter((int(m.start_node.id) for m in members), np.int64, nm),
        end=np.fromiter((int(m.end_node.id) for m in members), np.int64, nm),
        mtype=[m.member_type.value for m in members],
        rotation=np.fromiter((float(m.rotation_angle or 0.0) for m in members), np.float64, nm),
        section=[m.section.id if m.section is not None else None for m in members],
        mirror=[m.mirror for m in members], start_hinge=[oid(m.start_hinge) for m in members],
        end_hinge=[oid(m.end_hinge) for m in members], start_offset=[m.start_offset for m in members],
        end_offset=[m.end_offset for m in members], classification=[m.classification for m in members],
        ref_member=[oid(m.reference_member) for m in members], ref_node=[oid(m.reference_node) for m in members],
        weight=[0.0] * nm,                         # Member.to_dict emits weight 0.0 (members/member.py:186-193)
        pretension=[getattr(m, "pretension", None) for m in members],
        unstretched_length=[getattr(m, "unstretched_length", None) for m in members])
    settings = calc.settings.to_dict()
    model = {
        "member_sets": [ms.to_dict() for ms in calc.member_sets],
        "materials": [m.to_dict() for m in calc.get_unique_materials_from_all_member_sets()],
        "sections": [s_.to_dict() for s_ in calc.get_unique_sections_from_all_member_sets()],
        "nodal_supports": [ns.to_dict() for ns in calc.get_unique_nodal_support_from_all_member_sets()],
        "member_hinges": [h.to_dict() for h in calc.get_unique_member_hinges_from_all_member_sets()],
        "plate_surfaces": [ps.to_dict() for ps in calc.plate_surfaces],
        "plate_elements": [p.to_dict() for p in calc.plates],
    }

    def vec(v):
        return (v["X"], v["Y"], v["Z"]) if isinstance(v, dict) else tuple(v)
    cases = []
    for lc in calc.load_cases:
        cols = {}
        for key, items in (("nodal_loads", lc.nodal_loads), ("nodal_moments", lc.nodal_moments)):
            n = len(items)
            cols[key] = {"node": np.fromiter((int(x.node.id) for x in items), np.int64, n),
                         "magnitude": np.fromiter((float(x.magnitude) for x in items), np.float64, n),
                         "direction": np.array([vec(x.direction) for x in items], np.float64).reshape(n, 3)}
        dls = lc.distributed_loads
        n = len(dls)
        cols["distributed_loads"] = {
            "member": np.fromiter((int(x.member.id) for x in dls), np.int64, n),
            "magnitude": np.fromiter((float(x.magnitude) for x in dls), np.float64, n),
            "end_magnitude": np.fromiter((float(x.magnitude if x.end_magnitude is None else x.end_magnitude) for x in dls), np.float64, n),
            "direction": np.array([vec(x.direction) for x in dls], np.float64).reshape(n, 3),
            "start_frac": np.fromiter((float(x.start_frac or 0.0) for x in dls), np.float64, n),
            "end_frac": np.fromiter((float(1.0 if x.end_frac is None else x.end_frac) for x in dls), np.float64, n)}
        cases.append({"id": lc.id, "name": lc.name, "_columns": cols,
                      "surface_loads": [x.to_dict() for x in lc.surface_loads],
                      "member_point_loads": [x.to_dict() for x in lc.member_point_loads],
                      "member_point_moments": [x.to_dict() for x in lc.member_point_moments],
                      "plate_pressures": [x.to_dict() for x in lc.plate_pressures]})
    analysis = {"options": settings.get("analysis_options"),
                "load_cases": cases,
                "load_combinations": [c.to_dict() for c in calc.load_combinations],
                "imperfection_cases": [c.to_dict() for c in calc.imperfection_cases],
                "unity_checks": [c.to_dict() if hasattr(c, "to_dict") else c for c in calc.unity_checks]}
    return build_model_arrays(ncols, mcols, model, analysis)


def build_model_arrays(ncols: NodeColumns, mcols: MemberColumns, model: dict, analysis: dict) -> ModelArrays:
    """Everything after the two big lists were read into columns; ``model`` / ``analysis`` carry the small tables (supports,
    materials, sections, hinges, member sets, plates, load cases, combinations, imperfections, unity checks)."""
    opts = parse_options(analysis.get("options"))
    warnings: list = []
    if opts.axial_slack != Options().axial_slack and any(
            _MEMBER_TYPES.get(_norm_enum(t)) in (MT_TENSION, MT_COMPRESSION) for t in set(mcols.mtype)):
        # AnalysisOptions.axial_slack (pydantic_models.py:1335, default 500): the reference steers the residual stiffness of a
        # slack tension- / compression-only member with it ("slack model", tests/functionality/test_functionality.py:202-204);
        # its meaning is stated nowhere.  Here a switched-off member keeps hostlogic.INACTIVE_SCALE of its stiffness
        warnings.append(("W_AXIAL_SLACK_IGNORED", f"analysis option axial_slack = {opts.axial_slack:g} is not applied: inactive tension- / "
                         "compression-only members keep a fixed 1e-9 of their stiffness"))
    # RigidStrategy.LinearMpc (pydantic_models.py:1047-1049): applied by engine.analyze where hostlogic.mpc_transform can (first-
    # order analyses, slaves without supports / plates); elsewhere the result carries W_LINEAR_MPC_AS_RIGID_MEMBER
    nd = _parse_nodes(ncols, model, opts)
    sc = _parse_sections(model, opts, warnings)
    mats = {int(m["id"]): m for m in (model.get("materials") or [])}
    mb = _parse_members(mcols, nd, sc, model, opts, warnings)
    pl = _parse_plates(model, mats, nd, warnings)
    load_cases = _parse_load_cases(analysis, opts, nd, mb, sc, pl, mcols, warnings)
    combos = analysis.get("load_combinations") or []
    combo_ids, combo_names, combo_factors = _parse_combinations(combos, load_cases)
    wp = _warping_nodes(nd, mb, sc, warnings)
    return ModelArrays(
        n_real_nodes=wp["n_real"], warp_node=wp["warp_node"], w_member=wp["w_member"], w_aux=wp["w_aux"],
        node_ids=wp["node_ids"], xyz=wp["xyz"], node_support_id=wp["node_support_id"], bc_code=wp["bc_code"], bc_k=wp["bc_k"],
        m_ids=mcols.ids, m_n1=mb["m_n1"], m_n2=mb["m_n2"], m_type=mb["m_type"], m_roll=mb["m_roll"], m_prop=mb["m_prop"],
        m_rel=mb["m_rel"], prop=sc["prop"], prop_rho_a=np.array(sc["prop_rho_a"], np.float64), rel=mb["rel"],
        p_ids=pl["p_ids"], p_nodes=pl["p_nodes"], p_data=pl["p_data"], p_flags=pl["p_flags"],
        load_cases=load_cases, combo_ids=combo_ids, combo_names=combo_names, combo_factors=combo_factors,
        imperfections=list(analysis.get("imperfection_cases") or []), member_sets=mb["member_sets"],
        options=opts, warnings=warnings, support_curves=nd["support_curves"], hinge_curves=mb["hinge_curves"],
        unity_checks=list(analysis.get("unity_checks") or []), member_sets_raw=list(model.get("member_sets") or []),
        combo_limit_state=[c.get("limit_state") for c in combos],
        m_classification=list(mcols.classification), m_section_raw=mb["m_section_raw"],
        m_ref_member=mb["m_ref_member"], m_ref_node=mb["m_ref_node"],
        c_member=mb["c_member"], c_l0=mb["c_l0"], c_ea=mb["c_ea"], c_w=mb["c_w"], p_surface=pl["p_surface"],
        hinge_caps=mb["hinge_caps"],
    )


def _parse_nodes(ncols: NodeColumns, model: dict, opts: Options) -> dict:
    """Nodes (pydantic_models.py:Node; wire order = DOF numbering order) and supports (2104-2115; default all-Fixed
    nodalsupport.py:29-41)."""
    node_ids, xyz = ncols.ids, ncols.xyz
    nn = len(node_ids)
    node_index = _index_map(node_ids, "node")
    if not np.isfinite(xyz).all():
        raise WireError(f"node {int(node_ids[np.nonzero(~np.isfinite(xyz).all(axis=1))[0][0]])}: non-finite coordinate")
    supports = {int(s["id"]): s for s in (model.get("nodal_supports") or [])}
    bc_code = np.zeros((nn, 6), np.uint8)
    bc_k = np.zeros((nn, 6), np.float64)
    node_support_id = np.full(nn, -1, np.int64)
    warp_fixed = np.zeros(nn, bool)         # support restrains the warping DOF (NodalSupport.warping_condition)
    support_curves: list = []
    for i, sid in enumerate(ncols.support):
        if sid is None:
            continue
        if isinstance(sid, dict):
            sup, sid = sid, int(sid.get("id", 0))
        else:
            sid = int(sid)
            if sid not in supports:
                raise WireError(f"node {int(node_ids[i])}: unknown nodal_support {sid}")
            sup = supports[sid]
        node_support_id[i] = sid
        wc = sup.get("warping_condition")
        if isinstance(wc, dict):
            warp_fixed[i] = _norm_enum(wc.get("condition_type", "Free")) == "fixed"
        for d, name in enumerate(DOF_NAMES):
            cond = sup.get(name)
            if cond is None:
                continue  # absent = free
            ctype = _BC_TYPES.get(_norm_enum(cond.get("condition_type", "Fixed")).replace("-", ""))
            if ctype is None:
                ctype = _BC_TYPES.get(str(cond.get("condition_type")).strip().lower())
            if ctype is None:
                raise WireError(f"support {sid}.{name}: unknown condition {cond.get('condition_type')!r}")
            bc_code[i, d] = ctype
            if ctype == BC_SPRING:
                k = cond.get("stiffness")
                curve = cond.get("stiffness_curve")
                if curve and curve.get("points"):
                    # the curve overrides the constant value (supportcondition.py:19-23); start from the stiffness at zero force
                    pts = np.array([[float(p[0]), float(p[1])] for p in curve["points"]], np.float64)
                    comp = _SUPPORT_FORCE.get(str(curve.get("depends_on")))
                    if comp is None:
                        raise WireError(f"support {sid}.{name}: stiffness_curve.depends_on must be Fx/Fy/Fz/Mx/My/Mz, got {curve.get('depends_on')!r}")
                    if len(pts) < 2 or (np.diff(pts[:, 0]) < 0.0).any() or not (pts[:, 1] > 0.0).all() or not np.isfinite(pts).all():
                        raise WireError(f"support {sid}.{name}: stiffness_curve needs >= 2 points sorted by force with positive stiffness")
                    support_curves.append((6 * i + d, 6 * i + comp, pts, bool(curve.get("signed"))))
                    k = float(np.interp(0.0, pts[:, 0], pts[:, 1]))
                elif k is None:
                    raise WireError(f"support {sid}.{name}: Spring without stiffness")
                bc_k[i, d] = float(k)
            elif ctype in (BC_POS_ONLY, BC_NEG_ONLY) and cond.get("stiffness"):
                bc_k[i, d] = float(cond["stiffness"])        # unilateral spring (supportcondition.py:25-32)
    if opts.dimensionality == "2D":
        # out-of-plane DOFs restrained for every node (reference builders.py:108-110)
        for d in (2, 3, 4):
            free = bc_code[:, d] == BC_FREE
            bc_code[free, d] = BC_FIXED
    return dict(node_ids=node_ids, xyz=xyz, node_index=node_index, bc_code=bc_code, bc_k=bc_k, node_support_id=node_support_id,
                warp_fixed=warp_fixed, support_curves=support_curves, two_d=opts.dimensionality == "2D")


def _parse_sections(model: dict, opts: Options, warnings: list) -> dict:
    """materials / sections -> property table (pydantic_models.py:1052-1114, 1497-1504)."""
    mats = {int(m["id"]): m for m in (model.get("materials") or [])}
    secs = {int(s["id"]): s for s in (model.get("sections") or [])}
    prop_rows: list = []
    prop_rho_a: list = []
    sec_row: dict = {}
    sec_raw: dict = {}
    sec_eiyz: dict = {}
    for sid, s in secs.items():
        mid = s.get("material")
        if isinstance(mid, dict):
            mat = mid
        else:
            if int(mid) not in mats:
                raise WireError(f"section {sid}: unknown material {mid}")
            mat = mats[int(mid)]
        sec_raw[sid] = (s, mat)
        E, G = float(mat["e_mod"]), float(mat["g_mod"])
        A, Iy, Iz, J = float(s["area"]), float(s["i_y"]), float(s["i_z"]), float(s["j"])
        for what, v in (("e_mod", E), ("g_mod", G), ("area", A), ("i_y", Iy), ("i_z", Iz), ("j", J)):
            if not (np.isfinite(v) and v > 0.0):          # PositiveFloat in the schema (pydantic_models.py:1052-1114, 1497-1504)
                raise WireError(f"section {sid}: {what} must be a positive number, got {v!r}")
        asy = s.get("a_sy") if opts.include_shear_deformation else None
        asz = s.get("a_sz") if opts.include_shear_deformation else None
        eiw = E * float(s["i_w"]) if (opts.include_warping and s.get("i_w")) else 0.0
        row = [E * A, E * Iy, E * Iz, G * J,
               G * float(asy) if asy else 0.0, G * float(asz) if asz else 0.0,
               float(s.get("wpl_y") or 0.0), eiw]
        # product of inertia / principal axes (pydantic_models.py:1078-1090)
        iyz_m4, unit_note = i_yz_metres4(Iy, Iz, s.get("i_yz"))
        if unit_note:
            warnings.append(("W_IYZ_UNITS", f"section {sid}: {unit_note}"))
        Iy, Iz, Iyz, note = section_bending_inertias(Iy, Iz, iyz_m4, s.get("principal_axis_angle"))
        if not Iy * Iz > Iyz * Iyz:
            raise WireError(f"section {sid}: i_yz^2 must be smaller than i_y * i_z")
        row[1], row[2] = E * Iy, E * Iz
        sec_eiyz[sid] = E * Iyz
        if Iyz != 0.0 and (asy or asz):
            note = (note + "; " if note else "") + "the y-z bending coupling is Euler-Bernoulli (shear areas act about local y, z only)"
        if note:
            warnings.append(("W_PRINCIPAL_AXES", f"section {sid}: {note}"))
        sec_row[sid] = len(prop_rows)
        prop_rows.append(row)
        prop_rho_a.append(float(mat.get("density") or 0.0) * A)
    prop = np.array(prop_rows, np.float64).reshape(len(prop_rows), NPROP)
    return dict(prop=prop, prop_rho_a=prop_rho_a, sec_row=sec_row, sec_raw=sec_raw, sec_eiyz=sec_eiyz)


def _parse_members(mc: MemberColumns, nd: dict, sc: dict, model: dict, opts: Options, warnings: list) -> dict:
    """Members (pydantic_models.py:1507-1549), hinges (2062-2087; release fields are spring stiffnesses), member sets.
    The common case (no hinge, no end offset, a section without shear-centre offset or product of inertia) is vectorised;
    members that need a modifier record are visited one by one, in wire order."""
    hinges = {int(h["id"]): h for h in (model.get("member_hinges") or [])}
    sec_row, sec_raw, sec_eiyz = sc["sec_row"], sc["sec_raw"], sc["sec_eiyz"]
    node_index, xyz = nd["node_index"], nd["xyz"]
    capacity_warned: set = set()

    def hinge_curve(hid, name, curve):
        pts = np.array([[float(p[0]), float(p[1])] for p in curve["points"]], np.float64)
        comp = _MEMBER_FORCE.get(str(curve.get("depends_on")))
        if comp is None:
            raise WireError(f"member hinge {hid}.{name}: depends_on must be N/Vy/Vz/Mx/My/Mz, got {curve.get('depends_on')!r}")
        if len(pts) < 2 or (np.diff(pts[:, 0]) < 0.0).any() or not (pts[:, 1] > 0.0).all() or not np.isfinite(pts).all():
            raise WireError(f"member hinge {hid}.{name}: stiffness curve needs >= 2 points sorted by force with positive stiffness")
        return comp, pts, bool(curve.get("signed"))

    def hinge_row(hid) -> np.ndarray:
        r = np.full(6, RIGID_RELEASE)
        if hid is None:
            return r
        if int(hid) not in hinges:
            raise WireError(f"unknown member hinge {hid}")
        h = hinges[int(hid)]
        for d, f in enumerate(_HINGE_FIELDS):
            v = h.get(f)
            if v is not None:
                r[d] = max(float(v), 0.0)
        if hid not in capacity_warned and any(h.get(f) is not None for f in _HINGE_CAPACITIES):
            capacity_warned.add(hid)
            warnings.append(("W_HINGE_CAPACITY_IGNORED", f"member hinge {hid}: max_tension_* / max_moment_* capacities "
                             "are not enforced (the hinge keeps its elastic spring); the end forces are checked against them "
                             "after the analysis (W_HINGE_CAPACITY_EXCEEDED)"))
        for d, f in enumerate(_HINGE_CURVES):
            curve = h.get(f)
            if curve and curve.get("points"):
                # like the support curves, the curve overrides the constant; start from the stiffness at zero force
                r[d] = float(np.interp(0.0, *hinge_curve(int(hid), f, curve)[1].T))
        return r

    nm = len(mc.ids)
    m_ids = mc.ids
    member_index = _index_map(m_ids, "member")
    try:
        m_n1 = np.fromiter((node_index[int(v)] for v in mc.start), np.int32, nm)
        m_n2 = np.fromiter((node_index[int(v)] for v in mc.end), np.int32, nm)
    except KeyError:
        for i in range(nm):
            for v in (mc.start[i], mc.end[i]):
                if int(v) not in node_index:
                    raise WireError(f"member {int(m_ids[i])}: unknown node {int(v)!r}") from None
        raise
    type_code = {}
    for t in set(mc.mtype):
        code = _MEMBER_TYPES.get(_norm_enum(t))
        if code is None:
            bad = mc.mtype.index(t)
            raise WireError(f"member {int(m_ids[bad])}: unknown member_type {t!r}")
        type_code[t] = code
    m_type = np.fromiter((type_code[t] for t in mc.mtype), np.int32, nm)
    m_roll = np.radians(mc.rotation)
    m_prop = np.full(nm, -1, np.int32)
    for i, sid in enumerate(mc.section):
        if sid is not None:
            if int(sid) not in sec_row:
                raise WireError(f"member {int(m_ids[i])}: unknown section {sid}")
            m_prop[i] = sec_row[int(sid)]
        elif m_type[i] != MT_RIGID:
            raise WireError(f"member {int(m_ids[i])}: non-rigid member without section")
    m_section_raw = [sec_raw.get(int(sid)) if sid is not None else None for sid in mc.section]
    # sections that send their members to the modifier path: shear-centre offset (coupling on) or a product of inertia
    sec_special = set()
    for sid, (s_raw, _mat) in sec_raw.items():
        if (opts.include_shear_center_coupling and (float(s_raw.get("y_s") or 0.0) != 0.0 or float(s_raw.get("z_s") or 0.0) != 0.0)) \
                or sec_eiyz.get(sid, 0.0) != 0.0 or not np.isfinite(float(s_raw.get("y_s") or 0.0) + float(s_raw.get("z_s") or 0.0)):
            sec_special.add(sid)
    m_rel = np.full(nm, -1, np.int32)
    rel_rows: list = []
    hinge_curves: list = []
    for i in range(nm):
        sid = mc.section[i]
        if mc.start_hinge[i] is None and mc.end_hinge[i] is None and not mc.start_offset[i] and not mc.end_offset[i] \
                and (sid is None or int(sid) not in sec_special):
            continue
        t = int(m_type[i])
        # shear-centre eccentricity (pydantic_models.py:1106-1111; toggle 1358-1361; `mirror` negates it, 1521-1524)
        ys = zs = 0.0
        if sid is not None and opts.include_shear_center_coupling and t not in (MT_TRUSS, MT_RIGID):
            s_raw = sec_raw[int(sid)][0]
            sgn = -1.0 if mc.mirror[i] else 1.0
            ys, zs = sgn * float(s_raw.get("y_s") or 0.0), sgn * float(s_raw.get("z_s") or 0.0)
            if not (np.isfinite(ys) and np.isfinite(zs)):
                raise WireError(f"section {sid}: y_s / z_s must be finite")
        # y-z bending coupling of sections with a product of inertia (bending members only)
        eiyz = sec_eiyz.get(int(sid), 0.0) if (sid is not None and t == MT_NORMAL) else 0.0
        # end offsets (MemberEndOffset, pydantic_models.py:525-531, global X Y Z): the elastic axis runs between the
        # offset points, each tied to its node by a rigid arm
        offs = np.zeros(6)
        for e, o in enumerate((mc.start_offset[i], mc.end_offset[i])):
            if o:
                offs[3 * e:3 * e + 3] = [float(o.get(ax) or 0.0) for ax in ("X", "Y", "Z")]
        if not np.isfinite(offs).all():
            raise WireError(f"member {int(m_ids[i])}: end offsets must be finite")
        if mc.start_hinge[i] is not None or mc.end_hinge[i] is not None or ys != 0.0 or zs != 0.0 or offs.any() or eiyz != 0.0:
            m_rel[i] = len(rel_rows)
            rel_rows.append(np.concatenate([hinge_row(mc.start_hinge[i]), hinge_row(mc.end_hinge[i]), [ys, zs], offs, [eiyz]]))
            for e, hid in enumerate((mc.start_hinge[i], mc.end_hinge[i])):
                h = hinges.get(int(hid)) if hid is not None else None
                for d, f in enumerate(_HINGE_CURVES if h else ()):
                    if h.get(f) and h[f].get("points"):
                        hinge_curves.append((i, int(m_rel[i]), 6 * e + d) + hinge_curve(int(hid), f, h[f]))
    rel = np.array(rel_rows, np.float64).reshape(len(rel_rows), REL_STRIDE)
    hinge_caps: list = []
    for i in range(nm):
        for e, hid in enumerate((mc.start_hinge[i], mc.end_hinge[i])):
            h = hinges.get(int(hid)) if hid is not None else None
            for comp, f in enumerate(_HINGE_CAPACITIES[:6] if h else ()):
                if h.get(f) is not None and np.isfinite(float(h[f])) and float(h[f]) > 0.0:
                    hinge_caps.append((i, e, comp, float(h[f]), int(hid)))
    # MemberHinge.rotational_release_warp (pydantic_models.py:2078): spring stiffness between the warping of the member end and
    # that of its node, as the other release fields; None = continuous.  Acted on in _warping_nodes
    m_warp_rel = np.full((nm, 2), RIGID_RELEASE)
    for i in range(nm):
        for e, hid in enumerate((mc.start_hinge[i], mc.end_hinge[i])):
            h = hinges.get(int(hid)) if hid is not None else None
            if h is not None and h.get("rotational_release_warp") is not None:
                m_warp_rel[i, e] = max(float(h["rotational_release_warp"]), 0.0)
    # cables: elastic catenary between the end nodes; hinges, offsets and section eccentricities do not apply to them
    c_member = np.nonzero(m_type == MT_CABLE)[0].astype(np.int32)
    for i in c_member:
        if m_rel[i] >= 0:
            m_rel[i] = -1
            warnings.append(("W_CABLE_MODIFIERS_IGNORED", f"member {int(m_ids[i])}: hinges, end offsets and shear-centre offsets "
                             "do not apply to a cable"))
    c_chord = np.linalg.norm(xyz[m_n2[c_member]] - xyz[m_n1[c_member]], axis=1)
    c_ea = np.array([sc["prop"][m_prop[i]][0] for i in c_member], np.float64)
    pre = mc.pretension if mc.pretension is not None else [None] * nm
    given = mc.unstretched_length if mc.unstretched_length is not None else [None] * nm
    c_l0 = c_chord.copy()
    for j, i in enumerate(c_member):
        # `unstretched_length` wins over `pretension`: L0 = L_chord / (1 + pretension / EA); neither: the chord
        if given[i] is not None:
            c_l0[j] = float(given[i])
        elif pre[i] is not None:
            c_l0[j] = c_chord[j] / (1.0 + float(pre[i]) / c_ea[j])
        if not (np.isfinite(c_l0[j]) and c_l0[j] > 0.0 and c_ea[j] > 0.0):
            raise WireError(f"member {int(m_ids[i])}: a cable needs a positive unstretched length and axial stiffness")
    # own weight per unit length, "carried intrinsically in the catenary formulation" (1337-1340): the `weight` override when
    # non-zero, else density * area * |gravity_factor| when self weight is enabled
    g_dir = np.asarray(opts.gravity_direction, np.float64)
    g_dir = g_dir / (np.linalg.norm(g_dir) or 1.0)
    c_w = np.zeros((len(c_member), 3))
    for j, i in enumerate(c_member):
        w = abs(float(mc.weight[i] or 0.0))
        if w == 0.0 and opts.enable_self_weight:
            w = sc["prop_rho_a"][m_prop[i]] * abs(opts.gravity_factor)
        c_w[j] = w * g_dir
    if nm:
        axis = xyz[m_n2] - xyz[m_n1]
        with_rec = np.nonzero(m_rel >= 0)[0]
        if len(with_rec):
            axis[with_rec] += rel[m_rel[with_rec], REL_O2:REL_O2 + 3] - rel[m_rel[with_rec], REL_O1:REL_O1 + 3]
        length = np.linalg.norm(axis, axis=1)
        if (length <= 0.0).any():
            raise WireError(f"member {int(m_ids[np.nonzero(length <= 0.0)[0][0]])}: zero length (start and end of the member axis coincide)")
    # Member.reference_member / reference_node (unity checks across members, fers_core/examples/101_...py:61-67)
    m_ref_member = np.full(nm, -1, np.int32)
    m_ref_node = np.full(nm, -1, np.int32)
    for i in range(nm):
        rm, rn = mc.ref_member[i], mc.ref_node[i]
        if rm is None and rn is None:
            continue
        rm = rm.get("id") if isinstance(rm, dict) else rm
        rn = rn.get("id") if isinstance(rn, dict) else rn
        if rm is not None and int(rm) in member_index:
            m_ref_member[i] = member_index[int(rm)]
        if rn is not None and int(rn) in node_index:
            m_ref_node[i] = node_index[int(rn)]
    member_sets = {}
    for ms in (model.get("member_sets") or []):
        idx = []
        for mid in (ms.get("member_ids") or ms.get("members") or []):
            mid = mid["id"] if isinstance(mid, dict) else mid
            if int(mid) in member_index:
                idx.append(member_index[int(mid)])
        member_sets[int(ms["id"])] = idx
    return dict(m_n1=m_n1, m_n2=m_n2, m_type=m_type, m_roll=m_roll, m_prop=m_prop, m_rel=m_rel, rel=rel, member_index=member_index,
                member_ids=m_ids,
                hinge_curves=hinge_curves, m_section_raw=m_section_raw, m_ref_member=m_ref_member, m_ref_node=m_ref_node,
                member_sets=member_sets, c_member=c_member, c_l0=c_l0, c_ea=c_ea, c_w=c_w, m_warp_rel=m_warp_rel,
                hinge_caps=hinge_caps)


def _parse_plates(model: dict, mats: dict, nd: dict, warnings: list) -> dict:
    """Plate elements (pydantic_models.py:1700-1726); plate surfaces that arrive without elements are meshed here
    (`mesh_plate_surfaces`: ear clipping over the boundary nodes, the reference's own client-side rule)."""
    node_index, xyz = nd["node_index"], nd["xyz"]
    plates = list(model.get("plate_elements") or [])
    meshed = {p.get("source_surface_id") for p in plates}
    todo = [surf for surf in (model.get("plate_surfaces") or []) if surf.get("id") not in meshed]
    if todo:
        plates += mesh_plate_surfaces(todo, plates, nd, warnings)
        xyz = nd["xyz"]                                  # the mesher may have added boundary nodes
    p_ids_l, p_nodes_l, p_data_l, p_flags_l, p_surface_l = [], [], [], [], []
    for p in plates:
        nids = [nid["id"] if isinstance(nid, dict) else nid for nid in p["node_ids"]]
        mat = mats.get(int(p["material"])) if not isinstance(p.get("material"), dict) else p["material"]
        if mat is None:
            raise WireError(f"plate {p['id']}: unknown material")
        E, G = float(mat["e_mod"]), float(mat["g_mod"])
        nu = E / (2.0 * G) - 1.0
        th = float(p["thickness"])
        if not (np.isfinite([E, G, th]).all() and E > 0.0 and G > 0.0 and th > 0.0):
            raise WireError(f"plate {p['id']}: e_mod, g_mod and thickness must be positive numbers")
        beh = _norm_enum(p.get("behavior") or "Shell")
        flags = {"shell": 3, "membrane": 1, "plate": 2, "bending": 2}.get(beh, 3)
        if _norm_enum(p.get("plane_state") or "PlaneStress") == "planestrain":
            flags |= 4
        sm = p.get("stiffness_modifiers") or {}
        mods = [1.0 if sm.get(k) is None else float(sm[k]) for k in ("membrane", "bending", "shear")]
        if _norm_enum(p.get("theory") or "Auto") == "kirchhoff" and flags & 2:
            # PlateTheory (pydantic_models.py:816-819, field 1723): Kirchhoff -> discrete Kirchhoff triangle (DKT, no
            # transverse shear deformation); Auto / Mindlin -> DSG3.  The reference calls its own DKT element "currently
            # unreliable" (tests/nafems/reference_examples.py:16-18): this one is pinned on thin-plate theory instead
            flags |= 8
            if not any(code == "W_KIRCHHOFF_NO_SHEAR_FORCES" for code, _ in warnings):
                warnings.append(("W_KIRCHHOFF_NO_SHEAR_FORCES", "theory=Kirchhoff: thin-plate (DKT) elements carry no transverse shear "
                                 "strain; their results hold zero shear forces q_x, q_y (moments and deflections are thin-plate values)"))
        offset = float(p.get("offset") or 0.0)
        if not np.isfinite(offset):
            raise WireError(f"plate {p['id']}: offset must be a number")
        lx_dir = _vec3(p["local_x_direction"], "plate.local_x_direction") if p.get("local_x_direction") else None
        ortho = [0.0] * 7
        op = mat.get("orthotropic_plate")
        if op:
            # OrthotropicPlateMaterial (pydantic_models.py:663-681): E_x, E_y along the plate's local x / y, G_xy, nu_xy
            # ("nu_yx is derived from e_x nu_yx = e_y nu_xy"); the transverse shear moduli of the Mindlin part default
            # to G_xy when omitted (the engine that defines the default is closed: stated)
            ex, ey, gxy, nxy = (float(op[k]) for k in ("e_x", "e_y", "g_xy", "nu_xy"))
            gxz = float(op["g_xz"]) if op.get("g_xz") is not None else gxy
            gyz = float(op["g_yz"]) if op.get("g_yz") is not None else gxy
            if not (np.isfinite([ex, ey, gxy, nxy, gxz, gyz]).all() and min(ex, ey, gxy, gxz, gyz) > 0.0 and nxy * nxy * ey / ex < 1.0):
                raise WireError(f"plate {p['id']}: orthotropic plate material needs positive moduli and nu_xy^2 e_y / e_x < 1")
            ortho = [1.0, ex, ey, gxy, nxy, gxz, gyz]
            if flags & 4:
                warnings.append(("W_ORTHOTROPIC_PLANE_STRAIN", f"plate {p['id']}: plane strain with an orthotropic plate material "
                                 "needs out-of-plane constants the schema does not carry; the plane-stress law is used"))
        tris = [nids] if len(nids) == 3 else [[nids[0], nids[1], nids[2]], [nids[0], nids[2], nids[3]]]
        if len(nids) == 4:
            warnings.append(("W_QUAD_SPLIT", f"plate {p['id']}: QUAD4 split into two TRI3"))
        for t3 in tris:
            try:
                p_nodes_l.append([node_index[int(v)] for v in t3])
            except KeyError as e:
                raise WireError(f"plate {p['id']}: unknown node {e}") from None
            p_ids_l.append(int(p["id"]))
            p_surface_l.append(-1 if p.get("source_surface_id") is None else int(p["source_surface_id"]))
            theta = 0.0
            if lx_dir is not None:
                theta = plate_local_x_angle(xyz[p_nodes_l[-1]], lx_dir)
                if theta is None:
                    theta = 0.0
                    warnings.append(("W_PLATE_LOCAL_X_NORMAL", f"plate {p['id']}: local_x_direction is normal to the "
                                     "plate (or zero); default local axes kept"))
            p_data_l.append([float(p["thickness"]), E, nu, mods[0], mods[1], mods[2], theta] + ortho + [offset, 0.0])
            p_flags_l.append(flags)
    p_ids = np.array(p_ids_l, np.int64)
    p_nodes = np.array(p_nodes_l, np.int32).reshape(len(p_ids_l), 3)
    p_data = np.array(p_data_l, np.float64).reshape(len(p_ids_l), NPDATA)
    p_flags = np.array(p_flags_l, np.int32)
    if len(p_ids):
        x3 = xyz[p_nodes]
        area2 = np.linalg.norm(np.cross(x3[:, 1] - x3[:, 0], x3[:, 2] - x3[:, 0]), axis=1)
        if (area2 <= 0.0).any():
            raise WireError(f"plate {int(p_ids[np.nonzero(area2 <= 0.0)[0][0]])}: degenerate element (zero area)")
    plate_index = {}
    for i, pid in enumerate(p_ids_l):
        plate_index.setdefault(pid, []).append(i)
    return dict(plates=plates, p_ids=p_ids, p_ids_l=p_ids_l, p_nodes=p_nodes, p_data=p_data, p_flags=p_flags, plate_index=plate_index,
                p_surface=np.array(p_surface_l, np.int64))


def _load_records(c: dict, key: str):
    """One load list of a load case: the wire records, or columns prepared by from_fers (``_columns`` entry)."""
    cols = (c.get("_columns") or {}).get(key)
    return cols if cols is not None else list(c.get(key) or [])


def _parse_load_cases(analysis: dict, opts: Options, nd: dict, mb: dict, sc: dict, pl: dict, mcols: MemberColumns,
                      warnings: list) -> list:
    """Load cases (pydantic_models.py:2048-2059)."""
    node_index, member_index, xyz = nd["node_index"], mb["member_index"], nd["xyz"]
    m_prop, m_type = mb["m_prop"], mb["m_type"]
    plates, p_ids_l, p_nodes, plate_index = pl["plates"], pl["p_ids_l"], pl["p_nodes"], pl["plate_index"]
    load_cases = []
    cable_row = {int(m): j for j, m in enumerate(mb["c_member"])}
    raw_cases = list(analysis.get("load_cases") or [])
    sw_case_id = None
    if opts.enable_self_weight:
        sw_case_id = opts.self_weight_load_case_id
        if sw_case_id is None or all(int(c["id"]) != sw_case_id for c in raw_cases):
            sw_case_id = (max([int(c["id"]) for c in raw_cases], default=0) + 1) if sw_case_id is None else sw_case_id
            raw_cases.append({"id": sw_case_id, "name": "Self weight"})
    for c in raw_cases:
        lca = LoadCaseArrays(id=int(c["id"]), name=str(c.get("name", c["id"])))
        nod_node, nod_vec = [], []
        for key, lo in (("nodal_loads", 0), ("nodal_moments", 3)):
            recs = _load_records(c, key)
            if isinstance(recs, dict):                     # columns: node ids, magnitudes, directions [n,3]
                n = len(recs["node"])
                v = np.zeros((n, 6))
                v[:, lo:lo + 3] = recs["magnitude"][:, None] * recs["direction"]
                try:
                    nod_node += [node_index[int(i)] for i in recs["node"]]
                except KeyError as e:
                    raise WireError(f"load case {lca.id}: nodal load on unknown node {e}") from None
                nod_vec += list(v)
                continue
            for nl in recs:
                v = np.zeros(6)
                v[lo:lo + 3] = float(nl["magnitude"]) * _vec3(nl["direction"], key[:-1] + ".direction")
                nod_node.append(_node_of(nl, node_index)); nod_vec.append(v)
        dl_member, dl_data = [], []
        recs = _load_records(c, "distributed_loads")
        if isinstance(recs, dict):                         # columns: member ids, q1, q2, directions [n,3], a, b
            a, b = recs["start_frac"], recs["end_frac"]
            if not ((0.0 <= a) & (a <= b) & (b <= 1.0)).all():
                raise WireError(f"load case {lca.id}: distributed load needs 0<=start_frac<=end_frac<=1")
            try:
                dl_member += [member_index[int(i)] for i in recs["member"]]
            except KeyError as e:
                raise WireError(f"load case {lca.id}: distributed load on unknown member {e}") from None
            dl_data += list(np.concatenate([recs["magnitude"][:, None] * recs["direction"],
                                            recs["end_magnitude"][:, None] * recs["direction"], a[:, None], b[:, None]], axis=1))
        else:
            for dl in recs:
                mi = _member_of(dl, member_index)
                d = _vec3(dl["direction"], "distributed_load.direction")
                q1 = float(dl["magnitude"])
                q2 = float(dl["end_magnitude"]) if dl.get("end_magnitude") is not None else q1
                a = float(dl.get("start_frac", 0.0) or 0.0)
                b = float(dl.get("end_frac", 1.0) if dl.get("end_frac") is not None else 1.0)
                if not (0.0 <= a <= b <= 1.0):
                    raise WireError(f"distributed load {dl.get('id')}: need 0<=start_frac<=end_frac<=1")
                dl_member.append(mi)
                dl_data.append(np.concatenate([q1 * d, q2 * d, [a, b]]))
        if opts.enable_self_weight and int(c["id"]) == sw_case_id:
            g = np.asarray(opts.gravity_direction, np.float64)
            g = g / (np.linalg.norm(g) or 1.0)
            for i in range(len(mcols.ids)):
                if m_prop[i] < 0 or m_type[i] == MT_CABLE:
                    continue
                w = float(mcols.weight[i] or 0.0)
                if w == 0.0:
                    w = sc["prop_rho_a"][m_prop[i]] * abs(opts.gravity_factor)
                if w != 0.0:
                    dl_member.append(i)
                    dl_data.append(np.concatenate([w * g, w * g, [0.0, 1.0]]))
        pl_member, pl_data = [], []
        for pl_ in (c.get("member_point_loads") or []):
            row = np.zeros(8)
            row[:3] = float(pl_["magnitude"]) * _vec3(pl_["direction"], "member_point_load.direction")
            row[6] = float(pl_["position"])
            row[7] = 1.0 if _norm_enum(pl_.get("axes") or "global") == "local" else 0.0
            pl_member.append(_member_of(pl_, member_index)); pl_data.append(row)
        for pm_ in (c.get("member_point_moments") or []):
            row = np.zeros(8)
            row[3:6] = float(pm_["magnitude"]) * _vec3(pm_["direction"], "member_point_moment.direction")
            row[6] = float(pm_["position"])
            row[7] = 1.0 if _norm_enum(pm_.get("axes") or "global") == "local" else 0.0
            pl_member.append(_member_of(pm_, member_index)); pl_data.append(row)
        pp_plate, pp_vec = [], []
        for pp in (c.get("plate_pressures") or []):
            tgt = pp.get("target") or {}
            d = _vec3(pp["direction"], "plate_pressure.direction")
            ids = []
            # tagged union {"kind": "Element", "plate_element_id"} | {"kind": "Surface", "surface_id"}
            # (pydantic_models.py:763-787)
            if tgt.get("plate_element_id") is not None:
                ids += plate_index.get(int(tgt["plate_element_id"]), [])
            if tgt.get("surface_id") is not None:
                sid_ = int(tgt["surface_id"])
                ids += [i for i, p in enumerate(plates_flat_surface(plates, p_ids_l)) if p == sid_]
            for i in ids:
                w = 1.0
                if pp.get("projected"):      # pressure on the area projected onto `direction` (e.g. snow, 1733-1736)
                    x3 = xyz[p_nodes[i]]
                    nrm = np.cross(x3[1] - x3[0], x3[2] - x3[0])
                    dn = np.linalg.norm(d)
                    w = abs(float(nrm @ d)) / (np.linalg.norm(nrm) * dn) if dn > 0 else 0.0
                pp_plate.append(i); pp_vec.append(w * float(pp["magnitude"]) * d)
        for sl in (c.get("surface_loads") or []):
            if p_nodes.shape[0]:
                _surface_load_to_plates(sl, xyz, p_nodes, pp_plate, pp_vec, warnings)
            else:
                _surface_load_to_members(sl, xyz, mb, dl_member, dl_data, warnings)
        # uniform full-length loads on cables belong to the catenary (load per unstretched length); anything else on a cable
        # is carried to its end nodes like on a pin-ended member
        cab_q = np.zeros((len(cable_row), 3))
        if cable_row and dl_member:
            keep_m, keep_d = [], []
            for mi, row in zip(dl_member, dl_data):
                j = cable_row.get(int(mi))
                if j is None:
                    keep_m.append(mi); keep_d.append(row)
                elif row[6] == 0.0 and row[7] == 1.0 and np.array_equal(row[0:3], row[3:6]):
                    cab_q[j] += row[0:3]
                else:
                    keep_m.append(mi); keep_d.append(row)
                    warnings.append(("W_CABLE_LOAD_LUMPED", f"load case {lca.id}: a partial or varying distributed load on cable "
                                     f"{int(mcols.ids[mi])} is carried to its end nodes"))
            dl_member, dl_data = keep_m, keep_d
        lca.cab_q = cab_q
        lca.nod_node = np.array(nod_node, np.int32)
        lca.nod_vec = np.array(nod_vec, np.float64).reshape(len(nod_node), 6)
        lca.dl_member = np.array(dl_member, np.int32)
        lca.dl_data = np.array(dl_data, np.float64).reshape(len(dl_member), 8)
        lca.pl_member = np.array(pl_member, np.int32)
        lca.pl_data = np.array(pl_data, np.float64).reshape(len(pl_member), 8)
        lca.pp_plate = np.array(pp_plate, np.int32)
        lca.pp_vec = np.array(pp_vec, np.float64).reshape(len(pp_plate), 3)
        for arr in (lca.nod_vec, lca.dl_data, lca.pl_data, lca.pp_vec):
            if not np.isfinite(arr).all():
                raise WireError(f"load case {lca.id}: non-finite load value")
        load_cases.append(lca)
    return load_cases


def _parse_combinations(combos: list, load_cases: list):
    """Combinations (pydantic_models.py:491-505)."""
    lc_index = {lc.id: i for i, lc in enumerate(load_cases)}
    combo_ids = np.fromiter((int(c["id"]) for c in combos), np.int64, len(combos))
    combo_names = [str(c.get("name", c["id"])) for c in combos]
    combo_factors = np.zeros((len(combos), len(load_cases)), np.float64)
    for ci, c in enumerate(combos):
        lcf = c.get("load_case_factors") or c.get("load_cases_factors") or []
        if isinstance(lcf, dict):
            lcf = [{"load_case_id": k, "factor": v} for k, v in lcf.items()]
        for f in lcf:
            lid = int(f["load_case_id"])
            if lid not in lc_index:
                raise WireError(f"combination {c['id']}: unknown load case {lid}")
            combo_factors[ci, lc_index[lid]] += float(f["factor"])
    return combo_ids, combo_names, combo_factors


def _warping_nodes(nd: dict, mb: dict, sc: dict, warnings: list) -> dict:
    """Warping torsion: auxiliary nodes carrying psi."""
    node_ids, xyz, node_support_id, bc_code, bc_k = nd["node_ids"], nd["xyz"], nd["node_support_id"], nd["bc_code"], nd["bc_k"]
    m_prop, m_type, m_rel, rel, m_n1, m_n2, prop = mb["m_prop"], mb["m_type"], mb["m_rel"], mb["rel"], mb["m_n1"], mb["m_n2"], sc["prop"]
    nn = len(node_ids)
    warp_node = np.full(nn, -1, np.int32)
    w_member = np.zeros(0, np.int32)
    w_aux = np.zeros((0, 2), np.int32)
    has_w = (m_prop >= 0) & ~np.isin(m_type, (MT_TRUSS, MT_RIGID))
    has_w &= np.where(m_prop >= 0, prop[np.maximum(m_prop, 0), 7] if len(prop) else 0.0, 0.0) > 0.0
    if has_w.any():
        w_member = np.nonzero(has_w)[0].astype(np.int32)
        for i in w_member:
            if m_rel[i] >= 0 and rel[m_rel[i], REL_O1:REL_O2 + 3].any():
                warnings.append(("W_WARPING_OFFSET", f"member {int(mb['member_ids'][i])}: warping couplings follow the node-to-node axis, not the offset axis"))
        touched = np.unique(np.concatenate([m_n1[w_member], m_n2[w_member]]))
        warp_node[touched] = nn + np.arange(len(touched), dtype=np.int32)
        w_aux = np.stack([warp_node[m_n1[w_member]], warp_node[m_n2[w_member]]], axis=1).astype(np.int32)
        # warping released at a member end (MemberHinge.rotational_release_warp = 0): that end warps freely, whatever its node
        # does -- it gets an auxiliary node of its own (unrestrained psi: the bimoment vanishes there)
        owner = list(touched)                       # real node every auxiliary node belongs to
        free_psi = [False] * len(touched)
        for k, i in enumerate(w_member):
            for e in range(2):
                kw = mb["m_warp_rel"][i, e]
                if kw == 0.0:
                    w_aux[k, e] = nn + len(owner)
                    owner.append(int(m_n1[i] if e == 0 else m_n2[i]))
                    free_psi.append(True)
                elif kw > 0.0:
                    warnings.append(("W_WARP_RELEASE_SPRING", f"member {int(mb['member_ids'][i])}: a warping release with a spring "
                                     "stiffness is not modelled; the warping of this end stays continuous with its node"))
        touched = np.asarray(owner, np.int64)
        na = len(touched)
        aux_ids = -1 - node_ids[touched]
        aux_ids[len(free_psi) - sum(free_psi):] -= (1 << 40) * (1 + np.arange(sum(free_psi)))     # private ones: unique, negative
        node_ids = np.concatenate([node_ids, aux_ids])        # aux ids are negative: never on the wire
        xyz = np.concatenate([xyz, xyz[touched]])
        node_support_id = np.concatenate([node_support_id, node_support_id[touched]])
        aux_code = np.full((na, 6), BC_FIXED, np.uint8)
        aux_code[:, 0] = np.where(nd["warp_fixed"][touched] & ~np.asarray(free_psi), BC_FIXED, BC_FREE)
        bc_code = np.concatenate([bc_code, aux_code])
        bc_k = np.concatenate([bc_k, np.zeros((na, 6))])
    return dict(n_real=nn, warp_node=warp_node, w_member=w_member, w_aux=w_aux, node_ids=node_ids, xyz=xyz,
                node_support_id=node_support_id, bc_code=bc_code, bc_k=bc_k)


def member_axis(model, m: int) -> np.ndarray:
    """Vector from the start to the end of member m's elastic axis (end offsets included)."""
    d = model.xyz[model.m_n2[m]] - model.xyz[model.m_n1[m]]
    r = int(model.m_rel[m])
    if r >= 0:
        d = d + model.rel[r, REL_O2:REL_O2 + 3] - model.rel[r, REL_O1:REL_O1 + 3]
    return d


def i_yz_metres4(i_y: float, i_z: float, i_yz):
    """Section.i_yz in m^4, and a note for the warnings (or None).

    The schema states the unit of this one field as "mm^4 in JSON, scaled to m^4 by solver" (pydantic_models.py:1078-1082)
    while i_y / i_z travel in m^4 on the same record.  Any real section has I_yz^2 < I_y I_z, and a value in mm^4 is 1e12
    times its value in m^4, so the two readings never both pass that bound for a physical section: a value that breaks the
    bound as m^4 is mm^4 (the schema's reading, scaled by 1e-12); a value that already satisfies it in the units of
    i_y / i_z is used as it is, with a note."""
    v = float(i_yz or 0.0)
    if v == 0.0 or not np.isfinite(v):
        return v, None
    if v * v >= i_y * i_z:
        return v * 1.0e-12, None
    return v, ("i_yz is consistent with i_y / i_z as given (m^4) and is used unscaled; the schema states mm^4 for this field "
               "(that reading would make it 1e-12 of this value)")


def section_bending_inertias(i_y: float, i_z: float, i_yz, angle_deg):
    """(I_y, I_z, I_yz) about the member's local y and z axes, and a note for the warnings (or None).

    Section.i_yz / principal_axis_angle (pydantic_models.py:1078-1090).  I_z = int y^2, I_y = int z^2, I_yz = int y z dA in
    the member's local axes, so the bending energy is E/2 int (I_z v''^2 + I_y w''^2 + 2 I_yz v'' w'') dx.
    * i_yz given, no angle (or an angle that agrees with the triple: tan 2a = 2 I_yz / (I_z - I_y) to 1 degree):
      i_y / i_z are centroidal values, used with the coupling as they are;
    * angle given with i_yz absent or zero -- or inconsistent with the triple: i_y / i_z are principal values; the principal
      axes are the local axes turned by the angle in the sense of `rotation_angle` (y_p = c y + s z), which gives
      I_z = c^2 I_1 + s^2 I_2,  I_y = s^2 I_1 + c^2 I_2,  I_yz = s c (I_1 - I_2)  with I_1 = i_z, I_2 = i_y.
    The engine that defines the signs is closed: convention stated, unpinned."""
    iyz = float(i_yz or 0.0)
    ang = float(angle_deg or 0.0)
    if not (np.isfinite(iyz) and np.isfinite(ang)):
        raise WireError("section: i_yz / principal_axis_angle must be finite")
    if iyz != 0.0:
        if ang == 0.0:
            return i_y, i_z, iyz, None
        implied = 0.5 * np.degrees(np.arctan2(2.0 * iyz, i_z - i_y))
        diff = (ang - implied + 45.0) % 90.0 - 45.0          # principal directions repeat every 90 degrees
        if abs(diff) <= 1.0:
            return i_y, i_z, iyz, None
        note = "principal_axis_angle does not match i_y, i_z, i_yz: i_y / i_z taken as principal values, i_yz ignored"
    elif ang == 0.0:
        return i_y, i_z, 0.0, None
    else:
        note = None
    c, s_ = np.cos(np.radians(ang)), np.sin(np.radians(ang))
    return s_ * s_ * i_z + c * c * i_y, c * c * i_z + s_ * s_ * i_y, s_ * c * (i_z - i_y), note


def plate_local_x_angle(x3: np.ndarray, direction: np.ndarray):
    """PlateElement.local_x_direction (pydantic_models.py:1703) as the angle that turns the default plate frame (local z =
    normal over the node order, local x = global X projected into the plane, global Y when the normal is along X;
    csrc/fers_elem.cuh plate_geom) about the normal so that local x is the projection of ``direction`` into the plane.
    None when the direction has no in-plane part."""
    n = np.cross(x3[1] - x3[0], x3[2] - x3[0])
    nn = np.linalg.norm(n)
    if not nn > 0.0:
        return 0.0                          # degenerate plates are rejected further down
    ez = n / nn
    gx = np.array([0.0, 1.0, 0.0]) if abs(ez[0]) > 1.0 - 1e-6 else np.array([1.0, 0.0, 0.0])
    ex = gx - (gx @ ez) * ez
    ex /= np.linalg.norm(ex)
    ey = np.cross(ez, ex)
    d = np.asarray(direction, np.float64)
    a, b = float(d @ ex), float(d @ ey)
    if not np.hypot(a, b) > 1e-9 * max(np.linalg.norm(d), 1e-300):
        return None
    theta = float(np.arctan2(b, a))
    return 0.0 if abs(theta) < 1e-15 else theta


def plates_flat_surface(plates, p_ids_l):
    """source_surface_id per flattened TRI3 row (quads contribute two rows)."""
    by_id = {int(p["id"]): p.get("source_surface_id") for p in plates}
    return [(-1 if by_id[pid] is None else int(by_id[pid])) for pid in p_ids_l]


# ---- plate surfaces that arrive without elements -------------------------------------------------------------------
def densify_polygon(pts: np.ndarray, size) -> np.ndarray:
    """Boundary of a surface with its edges cut into ceil(length / size) equal segments (fers_core/plates/mesh.py:72-94, the
    reference's ``mesh.target_size`` rule: the boundary is refined, no interior nodes are made); consecutive duplicates
    dropped (:97-107, tolerance 1e-9)."""
    if size is None or not (float(size) > 0.0):
        return pts
    out = []
    n = len(pts)
    for i in range(n):
        a, b = pts[i], pts[(i + 1) % n]
        seg = max(1, int(np.ceil(np.linalg.norm(b - a) / float(size))))
        for k in range(seg):
            out.append(a + (b - a) * (k / seg))
    keep = []
    for q in out:
        if not keep or np.linalg.norm(q - keep[-1]) > 1e-9:
            keep.append(q)
    if len(keep) > 1 and np.linalg.norm(keep[0] - keep[-1]) <= 1e-9:
        keep.pop()
    return np.array(keep, float)


def _append_nodes(nd: dict, ids: list, xyz: np.ndarray) -> None:
    """Nodes made by the surface mesher join the node table after the wire nodes: unsupported, free (out-of-plane DOFs held
    in a 2D model, as for every node)."""
    n = len(ids)
    first = len(nd["node_ids"])
    nd["node_ids"] = np.concatenate([nd["node_ids"], np.asarray(ids, np.int64)])
    nd["xyz"] = np.concatenate([nd["xyz"], np.asarray(xyz, np.float64).reshape(n, 3)])
    code = np.zeros((n, 6), np.uint8)
    if nd.get("two_d"):
        code[:, 2:5] = BC_FIXED
    nd["bc_code"] = np.concatenate([nd["bc_code"], code])
    nd["bc_k"] = np.concatenate([nd["bc_k"], np.zeros((n, 6))])
    nd["node_support_id"] = np.concatenate([nd["node_support_id"], np.full(n, -1, np.int64)])
    nd["warp_fixed"] = np.concatenate([nd["warp_fixed"], np.zeros(n, bool)])
    for k, i in enumerate(ids):
        nd["node_index"][int(i)] = first + k


def mesh_plate_surfaces(surfaces: list, plates: list, nd: dict, warnings: list) -> list:
    """PlateSurface records (pydantic_models.py:1752-1790) -> TRI3 plate-element records, the rule of the reference's
    client-side mesher (fers_core/plates/mesh.py:12-31, :72-94, :106-148; plates/platesurface.py:180-237; fers/fers.py:
    1071-1086): the boundary polygon is refined to ``mesh.target_size`` (edges cut into equal segments), projected into
    its plane, oriented counter-clockwise and ear-clipped; a generated vertex that coincides (9 decimals) with a boundary
    node of the surface -- or with a node generated for a surface of the same ``mesh_group_id`` -- reuses it, the others
    become new nodes numbered after the largest node id; element ids continue after the largest plate id.  No interior
    nodes are made (the reference makes none either); ``mesh.divisions`` and openings are noted, not meshed."""
    node_index = nd["node_index"]
    next_id = max([int(p["id"]) for p in plates], default=0) + 1
    next_node = int(max(nd["node_ids"].max() if len(nd["node_ids"]) else 0, 0)) + 1
    group_cache: dict = {}
    out = []
    for surf in surfaces:
        xyz = nd["xyz"]
        nids = [int(v["id"] if isinstance(v, dict) else v) for v in surf.get("boundary_node_ids") or []]
        for v in nids:
            if v not in node_index:
                raise WireError(f"plate surface {surf.get('id')}: unknown boundary node {v}")
        pts = xyz[[node_index[v] for v in nids]]
        if len(nids) > 1 and np.linalg.norm(pts[0] - pts[-1]) <= 1e-9:
            nids, pts = nids[:-1], pts[:-1]
        if len(nids) < 3:
            raise WireError(f"plate surface {surf.get('id')}: needs at least three boundary nodes")
        mesh = surf.get("mesh") or {}
        key = lambda q: (round(float(q[0]), 9), round(float(q[1]), 9), round(float(q[2]), 9))       # noqa: E731
        gid = surf.get("mesh_group_id")
        cache = group_cache.setdefault(gid, {}) if gid is not None else {}
        for v, q in zip(nids, pts):
            cache[key(q)] = v
        dense = densify_polygon(pts, mesh.get("target_size"))
        tris = ear_clip_polygon(dense, f"plate surface {surf.get('id')}")
        # node ids in the order the triangles first meet a new vertex (platesurface.py:207-216)
        ids, new_ids, new_xyz = [cache.get(key(q)) for q in dense], [], []
        for tri in tris:
            for i in tri:
                if ids[i] is None:
                    v = cache.get(key(dense[i]))
                    if v is None:
                        v = next_node
                        next_node += 1
                        cache[key(dense[i])] = v
                        new_ids.append(v); new_xyz.append(dense[i])
                    ids[i] = v
        if new_ids:
            _append_nodes(nd, new_ids, np.array(new_xyz))
        if mesh.get("divisions"):
            warnings.append(("W_SURFACE_MESH_DIVISIONS_IGNORED", f"plate surface {surf.get('id')}: mesh.divisions is not applied "
                             "(the boundary is refined to mesh.target_size only)"))
        if surf.get("openings"):
            warnings.append(("W_SURFACE_OPENINGS_IGNORED", f"plate surface {surf.get('id')}: openings are not cut out of the mesh"))
        for (i, j, k) in tris:
            out.append({"id": next_id, "node_ids": [ids[i], ids[j], ids[k]], "material": surf["material"],
                        "thickness": surf["thickness"], "behavior": surf.get("behavior"), "theory": surf.get("theory"),
                        "plane_state": surf.get("plane_state"), "offset": surf.get("offset"),
                        "stiffness_modifiers": surf.get("stiffness_modifiers"), "local_x_direction": surf.get("local_x_direction"),
                        "source_surface_id": surf.get("id"), "classification": surf.get("classification")})
            next_id += 1
    return out


def ear_clip_polygon(pts: np.ndarray, what: str = "polygon") -> list:
    """Triangles (index triples into ``pts`` [n,3]) of a simple planar polygon by ear clipping, counter-clockwise about the
    polygon normal (fers_core/plates/mesh.py:47-66 plane basis, :106-148 ear clipping, tolerance 1e-9)."""
    tol = 1e-9
    o = pts[0]
    u = nrm = None
    for i in range(1, len(pts) - 1):
        ea, eb = pts[i] - o, pts[i + 1] - o
        if u is None and np.linalg.norm(ea) > tol:
            u = ea / np.linalg.norm(ea)
        c = np.cross(ea, eb)
        if np.linalg.norm(c) > tol:
            nrm = c / np.linalg.norm(c)
            break
    if u is None or nrm is None:
        raise WireError(f"{what}: degenerate polygon")
    if (np.abs((pts - o) @ nrm) > 1e-7).any():
        raise WireError(f"{what}: polygon must be planar")
    v = np.cross(nrm, u)
    v /= np.linalg.norm(v)
    p2 = np.stack([(pts - o) @ u, (pts - o) @ v], axis=1)
    idx = list(range(len(pts)))
    area = 0.5 * sum(p2[i, 0] * p2[(i + 1) % len(pts), 1] - p2[(i + 1) % len(pts), 0] * p2[i, 1] for i in range(len(pts)))
    if area < 0.0:
        idx.reverse()

    def cross2(a, b):
        return a[0] * b[1] - a[1] * b[0]

    def inside(p, a, b, c):
        v0, v1, v2 = c - a, b - a, p - a
        d00, d01, d02, d11, d12 = v0 @ v0, v0 @ v1, v0 @ v2, v1 @ v1, v1 @ v2
        den = d00 * d11 - d01 * d01
        if abs(den) <= tol:
            return False
        uu, vv = (d11 * d02 - d01 * d12) / den, (d00 * d12 - d01 * d02) / den
        return uu >= -tol and vv >= -tol and uu + vv <= 1.0 + tol
    tris = []
    while len(idx) > 3:
        for off, cur in enumerate(idx):
            prev, nxt = idx[off - 1], idx[(off + 1) % len(idx)]
            a, b, c = p2[prev], p2[cur], p2[nxt]
            if cross2(b - a, c - b) <= tol:
                continue
            if any(q not in (prev, cur, nxt) and inside(p2[q], a, b, c) for q in idx):
                continue
            tris.append((prev, cur, nxt))
            del idx[off]
            break
        else:
            raise WireError(f"{what}: polygon could not be triangulated (use a simple, non-self-intersecting polygon)")
    tris.append((idx[0], idx[1], idx[2]))
    return tris


def _node_of(item: dict, node_index: dict) -> int:
    nid = item.get("node")
    nid = nid["id"] if isinstance(nid, dict) else nid
    if int(nid) not in node_index:
        raise WireError(f"load {item.get('id')}: unknown node {nid}")
    return node_index[int(nid)]


def _member_of(item: dict, member_index: dict) -> int:
    mid = item.get("member")
    mid = mid["id"] if isinstance(mid, dict) else mid
    if int(mid) not in member_index:
        raise WireError(f"load {item.get('id')}: unknown member {mid}")
    return member_index[int(mid)]


def _surface_load_to_plates(sl, xyz, p_nodes, pp_plate, pp_vec, warnings):
    """SurfaceLoad polygon -> per-plate traction (pydantic_models.py:1962-1967).

    A plate receives the full pressure when its centroid lies inside the polygon (evaluated in
    the polygon's own plane); plates cut by the polygon edge are not split (SURVEY.md A.9 [G]).
    """
    if p_nodes.shape[0] == 0:
        warnings.append(("W_SURFACE_LOAD_NO_PLATES", f"surface load {sl.get('id')}: no plate elements to load"))
        return
    poly = np.array([_vec3(v, "surface_load.polygon") for v in sl["polygon"]])
    d = _vec3(sl["direction"], "surface_load.direction")
    q = float(sl["magnitude"])
    o = poly[0]
    nrm = np.zeros(3)
    for i in range(len(poly)):
        nrm += np.cross(poly[i] - o, poly[(i + 1) % len(poly)] - o)
    nn = np.linalg.norm(nrm)
    if nn == 0.0:
        raise WireError(f"surface load {sl.get('id')}: degenerate polygon")
    nrm /= nn
    e1 = poly[1] - o
    e1 /= np.linalg.norm(e1)
    e2 = np.cross(nrm, e1)
    p2 = np.stack([(poly - o) @ e1, (poly - o) @ e2], 1)
    cen = xyz[p_nodes].mean(axis=1)
    c2 = np.stack([(cen - o) @ e1, (cen - o) @ e2], 1)
    off = np.abs((cen - o) @ nrm)
    span = max(np.ptp(p2[:, 0]), np.ptp(p2[:, 1]), 1e-12)
    inside = np.zeros(len(cen), bool)
    x, y = c2[:, 0], c2[:, 1]
    j = len(p2) - 1
    for i in range(len(p2)):
        xi, yi, xj, yj = p2[i, 0], p2[i, 1], p2[j, 0], p2[j, 1]
        cond = ((yi > y) != (yj > y)) & (x < (xj - xi) * (y - yi) / (yj - yi + 1e-300) + xi)
        inside ^= cond
        j = i
    inside &= off <= 1e-6 * span + 1e-9
    for i in np.nonzero(inside)[0]:
        pp_plate.append(int(i)); pp_vec.append(q * d)


def _surface_load_to_members(sl, xyz, mb, dl_member, dl_data, warnings):
    """SurfaceLoad on a model without plates (pydantic_models.py:1962-1967): the area load travels one way to the members
    that lie in the polygon's plane along ``distribution_direction`` (fers_core/loads/surfaceload.py:24-45; the only
    in-repo model, tests/functionality/test_surface_loads.py:8-37, has two members along X under a polygon with
    distribution_direction = X).  Every such member line takes the strip of the polygon between the mid-lines to its
    neighbours (the polygon edge beyond the outermost lines); along the member the strip width is piecewise linear, so
    the load becomes trapezoid member loads between the break points (polygon vertices and strip-boundary crossings).
    Total applied force = magnitude x polygon area when the members span the polygon.  Semantics stated, unpinned."""
    poly = np.array([_vec3(v, "surface_load.polygon") for v in sl["polygon"]])
    if len(poly) < 3:
        raise WireError(f"surface load {sl.get('id')}: polygon needs at least three vertices")
    d = _vec3(sl["direction"], "surface_load.direction")
    q = float(sl["magnitude"])
    dd = sl.get("distribution_direction")
    o = poly[0]
    nrm = np.zeros(3)
    for i in range(len(poly)):
        nrm += np.cross(poly[i] - o, poly[(i + 1) % len(poly)] - o)
    if np.linalg.norm(nrm) == 0.0:
        raise WireError(f"surface load {sl.get('id')}: degenerate polygon")
    nrm /= np.linalg.norm(nrm)
    m_n1, m_n2 = mb["m_n1"], mb["m_n2"]
    span = max(np.ptp(poly, axis=0).max(), 1e-12)
    tolp = 1e-6 * span + 1e-9
    in_plane = (np.abs((xyz[m_n1] - o) @ nrm) <= tolp) & (np.abs((xyz[m_n2] - o) @ nrm) <= tolp) if len(m_n1) else np.zeros(0, bool)
    if dd is None:
        # no direction given: the direction most in-plane members follow
        cand = np.nonzero(in_plane)[0]
        if len(cand):
            ax = xyz[m_n2[cand]] - xyz[m_n1[cand]]
            ax /= np.linalg.norm(ax, axis=1)[:, None]
            ref = ax[0]
            dd = ref
    if dd is None:
        warnings.append(("W_SURFACE_LOAD_NO_PLATES", f"surface load {sl.get('id')}: no plate elements and no members in its plane"))
        return
    e_s = _vec3(dd, "surface_load.distribution_direction")
    e_s = e_s - (e_s @ nrm) * nrm
    if np.linalg.norm(e_s) == 0.0:
        raise WireError(f"surface load {sl.get('id')}: distribution_direction is normal to the polygon")
    e_s /= np.linalg.norm(e_s)
    e_t = np.cross(nrm, e_s)
    P = np.stack([(poly - o) @ e_s, (poly - o) @ e_t], axis=1)                  # polygon in (s, t)
    axis = xyz[m_n2] - xyz[m_n1] if len(m_n1) else np.zeros((0, 3))
    L = np.linalg.norm(axis, axis=1)
    par = in_plane & (np.abs(np.abs((axis / np.maximum(L, 1e-300)[:, None]) @ e_s) - 1.0) <= 1e-6)
    members = np.nonzero(par)[0]
    if len(members) == 0:
        warnings.append(("W_SURFACE_LOAD_NO_PLATES", f"surface load {sl.get('id')}: no plate elements and no members along its "
                         "distribution direction"))
        return
    t_of = ((xyz[m_n1[members]] - o) @ e_t)
    lines = np.unique(np.round(t_of / tolp).astype(np.int64))                       # distinct member lines
    t_lines = np.array([t_of[np.round(t_of / tolp).astype(np.int64) == k].mean() for k in lines])
    t_min, t_max = P[:, 1].min(), P[:, 1].max()
    bounds = np.concatenate([[min(t_min, t_lines[0])], 0.5 * (t_lines[1:] + t_lines[:-1]), [max(t_max, t_lines[-1])]])

    def width(sv, lo, hi):
        """Length of {t in [lo, hi] : (sv, t) inside the polygon} (even-odd rule over the crossings of the line s = sv)."""
        ts = []
        n = len(P)
        for i in range(n):
            (s1, t1), (s2, t2) = P[i], P[(i + 1) % n]
            if (s1 > sv) != (s2 > sv):
                ts.append(t1 + (sv - s1) * (t2 - t1) / (s2 - s1))
        ts.sort()
        return sum(max(0.0, min(b, hi) - max(a, lo)) for a, b in zip(ts[0::2], ts[1::2]))
    for li, tl in enumerate(t_lines):
        lo, hi = bounds[li], bounds[li + 1]
        # break points along s: polygon vertices and the crossings of the polygon edges with the strip boundaries
        brk = list(P[:, 0])
        n = len(P)
        for i in range(n):
            (s1, t1), (s2, t2) = P[i], P[(i + 1) % n]
            for tb in (lo, hi):
                if (t1 > tb) != (t2 > tb):
                    brk.append(s1 + (tb - t1) * (s2 - s1) / (t2 - t1))
        on_line = members[np.abs(t_of - tl) <= 2 * tolp]
        for m in on_line:
            s_a, s_b = float((xyz[m_n1[m]] - o) @ e_s), float((xyz[m_n2[m]] - o) @ e_s)
            s_lo, s_hi = min(s_a, s_b), max(s_a, s_b)
            cuts = sorted({s_lo, s_hi, *[b for b in brk if s_lo < b < s_hi]})
            for c0, c1 in zip(cuts[:-1], cuts[1:]):
                eps = 1e-9 * (c1 - c0)
                w0, w1 = width(c0 + eps, lo, hi), width(c1 - eps, lo, hi)
                if w0 == 0.0 and w1 == 0.0:
                    continue
                f0, f1 = (c0 - s_a) / (s_b - s_a), (c1 - s_a) / (s_b - s_a)
                if f0 > f1:                                                  # member runs against e_s
                    f0, f1, w0, w1 = f1, f0, w1, w0
                dl_member.append(int(m))
                dl_data.append(np.concatenate([q * w0 * d, q * w1 * d, [f0, f1]]))
