"""Synthetic workloads of BASELINE.json (configs[3] and configs[4]) built directly as arrays.

cfg4: regular 3D space frame, bays 6.0 m x 6.0 m, storey 3.5 m, Y up, node index
``i + (nx+1)*(k + (nz+1)*j)`` (i along X fastest, k along Z, j = level), base level all-fixed;
columns HEB300, beams IPE400 (properties from the reference catalogue
fers_core/sections/steel_sections.generated.json); 8 load cases (G, Q checkerboard, Wx, Wz and
four notional cases) and ``n_combos`` combinations whose factors are drawn with
``numpy.random.default_rng(20260119)`` from {0, 0.7, 1.0, 1.35, 1.5}.
"""
from __future__ import annotations

import numpy as np

from . import wire

E_STEEL, G_STEEL, RHO_STEEL = 210e9, 81e9, 7850.0
# A, Iy, Iz, J, a_sy, a_sz (m^2, m^4) -- reference catalogue values
HEB300 = (0.014911962508337631, 8.562957407497643e-05, 0.00025171852141282227, 1.8899425072297412e-06,
          0.010276356568545326, 0.0032263438840005967)
IPE400 = (0.008448888430980666, 1.3178706258205786e-05, 0.00023136476234289616, 5.087647357490124e-07,
          0.004533525707121986, 0.003379320620928902)
COMBO_SEED = 20260119
COMBO_LEVELS = np.array([0.0, 0.7, 1.0, 1.35, 1.5])


def _prop_row(sec, timoshenko: bool) -> list:
    A, Iy, Iz, J, asy, asz = sec
    return [E_STEEL * A, E_STEEL * Iy, E_STEEL * Iz, G_STEEL * J,
            G_STEEL * asy if timoshenko else 0.0, G_STEEL * asz if timoshenko else 0.0, 0.0, 0.0]


def grid_frame(nx: int, nz: int, ny: int, bay: float = 6.0, storey: float = 3.5, n_combos: int = 64,
               timoshenko: bool = False, order: str = "LINEAR", solve_loadcases: bool = False) -> wire.ModelArrays:
    """nx x nz bays, ny storeys.  Returns a ModelArrays (ids = index + 1)."""
    px, pz, py = nx + 1, nz + 1, ny + 1
    nn = px * pz * py
    i, k, j = np.meshgrid(np.arange(px), np.arange(pz), np.arange(py), indexing="ij")
    idx = (i + px * (k + pz * j)).reshape(-1)
    xyz = np.zeros((nn, 3))
    xyz[idx, 0] = (i * bay).reshape(-1)
    xyz[idx, 1] = (j * storey).reshape(-1)
    xyz[idx, 2] = (k * bay).reshape(-1)

    def nid(ii, kk, jj):
        return ii + px * (kk + pz * jj)

    # columns
    ii, kk, jj = np.meshgrid(np.arange(px), np.arange(pz), np.arange(ny), indexing="ij")
    c1, c2 = nid(ii, kk, jj).reshape(-1), nid(ii, kk, jj + 1).reshape(-1)
    # beams along X and Z on levels 1..ny
    ii, kk, jj = np.meshgrid(np.arange(nx), np.arange(pz), np.arange(1, py), indexing="ij")
    bx1, bx2 = nid(ii, kk, jj).reshape(-1), nid(ii + 1, kk, jj).reshape(-1)
    bx_i, bx_k = ii.reshape(-1), kk.reshape(-1)
    ii, kk, jj = np.meshgrid(np.arange(px), np.arange(nz), np.arange(1, py), indexing="ij")
    bz1, bz2 = nid(ii, kk, jj).reshape(-1), nid(ii, kk + 1, jj).reshape(-1)
    bz_i, bz_k = ii.reshape(-1), kk.reshape(-1)
    n1 = np.concatenate([c1, bx1, bz1]); n2 = np.concatenate([c2, bx2, bz2])
    is_beam = np.concatenate([np.zeros(len(c1), bool), np.ones(len(bx1) + len(bz1), bool)])
    checker = np.concatenate([np.zeros(len(c1), bool), (bx_i + bx_k) % 2 == 0, (bz_i + bz_k) % 2 == 0])
    order_ = np.lexsort((n2, n1))                 # members sorted by start node: assembly locality
    n1, n2, is_beam, checker = n1[order_], n2[order_], is_beam[order_], checker[order_]
    nm = len(n1)

    bc_code = np.zeros((nn, 6), np.uint8)
    base = np.arange(px * pz)                     # level j = 0
    bc_code[base] = wire.BC_FIXED
    node_support_id = np.full(nn, -1, np.int64)
    node_support_id[base] = 1

    prop = np.array([_prop_row(HEB300, timoshenko), _prop_row(IPE400, timoshenko)])
    m_prop = np.where(is_beam, 1, 0).astype(np.int32)

    # ---- load cases -----------------------------------------------------------------
    beams = np.nonzero(is_beam)[0].astype(np.int32)

    def udl_case(cid, name, members, q):
        lc = wire.LoadCaseArrays(id=cid, name=name)
        lc.dl_member = members.astype(np.int32)
        row = np.array([0.0, q, 0.0, 0.0, q, 0.0, 0.0, 1.0])
        lc.dl_data = np.tile(row, (len(members), 1))
        return lc

    def nodal_case(cid, name, nodes, vec):
        lc = wire.LoadCaseArrays(id=cid, name=name)
        lc.nod_node = nodes.astype(np.int32)
        lc.nod_vec = np.tile(np.asarray(vec, np.float64), (len(nodes), 1))
        return lc

    lvl = idx // (px * pz)
    ii_all = idx % px
    kk_all = (idx // px) % pz
    above = lvl >= 1
    cases = [
        udl_case(1, "G", beams, -25e3),
        udl_case(2, "Q", np.nonzero(is_beam & checker)[0], -15e3),
        nodal_case(3, "Wx", idx[above & (ii_all == 0)], (12e3, 0, 0, 0, 0, 0)),
        nodal_case(4, "Wz", idx[above & (kk_all == 0)], (0, 0, 12e3, 0, 0, 0)),
        nodal_case(5, "N+x", idx[above], (1.5e3, 0, 0, 0, 0, 0)),
        nodal_case(6, "N-x", idx[above], (-1.5e3, 0, 0, 0, 0, 0)),
        nodal_case(7, "N+z", idx[above], (0, 0, 1.5e3, 0, 0, 0)),
        nodal_case(8, "N-z", idx[above], (0, 0, -1.5e3, 0, 0, 0)),
    ]
    rng = np.random.default_rng(COMBO_SEED)
    factors = COMBO_LEVELS[rng.integers(0, len(COMBO_LEVELS), size=(n_combos, len(cases)))]
    factors[:, 0] = np.where(factors[:, 0] == 0.0, 1.0, factors[:, 0])   # gravity always present

    opts = wire.Options(order=order, nonlinear_method="P_DELTA", solve_loadcases=solve_loadcases)
    return wire.ModelArrays(
        node_ids=np.arange(1, nn + 1, dtype=np.int64), xyz=xyz, node_support_id=node_support_id,
        bc_code=bc_code, bc_k=np.zeros((nn, 6)),
        m_ids=np.arange(1, nm + 1, dtype=np.int64), m_n1=n1.astype(np.int32), m_n2=n2.astype(np.int32),
        m_type=np.zeros(nm, np.int32), m_roll=np.zeros(nm), m_prop=m_prop, m_rel=np.full(nm, -1, np.int32),
        prop=prop, prop_rho_a=np.array([RHO_STEEL * HEB300[0], RHO_STEEL * IPE400[0]]), rel=np.zeros((0, wire.REL_STRIDE)),
        p_ids=np.zeros(0, np.int64), p_nodes=np.zeros((0, 3), np.int32), p_data=np.zeros((0, wire.NPDATA)),
        p_flags=np.zeros(0, np.int32), load_cases=cases,
        combo_ids=np.arange(1, n_combos + 1, dtype=np.int64),
        combo_names=[f"C{c + 1:02d}" for c in range(n_combos)], combo_factors=factors,
        imperfections=[], member_sets={}, options=opts,
    )


def to_wire_document(model: wire.ModelArrays) -> dict:
    """Serialise an array model back to the reference wire form (small models / tests)."""
    sup = {"id": 1, "classification": None, "warping_condition": None}
    for d in wire.DOF_NAMES:
        sup[d] = {"condition_type": "Fixed", "stiffness": None}
    nodes = [{"id": int(model.node_ids[i]), "classification": "", "X": float(model.xyz[i, 0]),
              "Y": float(model.xyz[i, 1]), "Z": float(model.xyz[i, 2]),
              "nodal_support": (1 if model.node_support_id[i] >= 0 else None)} for i in range(model.n_nodes)]
    mats = [{"id": 1, "name": "S355", "e_mod": E_STEEL, "g_mod": G_STEEL, "density": RHO_STEEL, "yield_stress": 355e6}]
    secs = []
    for s in range(model.prop.shape[0]):
        EA, EIy, EIz, GJ, GAsy, GAsz = model.prop[s, :6]
        sec = {"id": s + 1, "name": f"S{s + 1}", "material": 1, "area": EA / E_STEEL, "i_y": EIy / E_STEEL,
               "i_z": EIz / E_STEEL, "j": GJ / G_STEEL}
        if GAsy > 0:
            sec["a_sy"] = GAsy / G_STEEL
        if GAsz > 0:
            sec["a_sz"] = GAsz / G_STEEL
        secs.append(sec)
    members = [{"id": int(model.m_ids[m]), "start_node_id": int(model.node_ids[model.m_n1[m]]),
                "end_node_id": int(model.node_ids[model.m_n2[m]]), "section": int(model.m_prop[m]) + 1,
                "rotation_angle": float(np.degrees(model.m_roll[m])), "member_type": "Normal", "weight": 0.0,
                "start_hinge": None, "end_hinge": None} for m in range(model.n_members)]
    cases = []
    for lc in model.load_cases:
        c = {"id": lc.id, "name": lc.name, "nodal_loads": [], "nodal_moments": [], "distributed_loads": [],
             "member_point_loads": [], "member_point_moments": [], "surface_loads": [], "plate_pressures": []}
        for q, (n, v) in enumerate(zip(lc.nod_node, lc.nod_vec)):
            f = v[:3]
            mag = float(np.linalg.norm(f))
            if mag > 0:
                c["nodal_loads"].append({"id": q + 1, "node": int(model.node_ids[n]), "magnitude": mag,
                                         "direction": [float(x) for x in f / mag], "load_type": "force"})
        for q, (m, row) in enumerate(zip(lc.dl_member, lc.dl_data)):
            q1 = row[0:3]
            mag = float(np.linalg.norm(q1))
            d = q1 / mag
            c["distributed_loads"].append({"id": q + 1, "member": int(model.m_ids[m]), "magnitude": mag,
                                           "end_magnitude": float(row[3:6] @ d), "direction": [float(x) for x in d],
                                           "start_frac": float(row[6]), "end_frac": float(row[7])})
        cases.append(c)
    combos = [{"id": int(model.combo_ids[c]), "name": model.combo_names[c],
               "load_case_factors": [{"load_case_id": lc.id, "factor": float(model.combo_factors[c, li])}
                                     for li, lc in enumerate(model.load_cases)]}
              for c in range(model.combo_ids.shape[0])]
    o = model.options
    return {
        "schema_version": 1,
        "settings": {"unit_settings": {"system": "metric", "lengthUnit": "m", "forceUnit": "N"},
                     "general_info": {"project_name": "synthetic", "author": "synth", "version": "1.0"}},
        "model": {"nodes": nodes, "members": members, "member_sets": [], "materials": mats, "sections": secs,
                  "nodal_supports": [sup], "member_hinges": [], "plate_elements": [], "plate_surfaces": []},
        "analysis": {"options": {"solve_loadcases": o.solve_loadcases, "tolerance": o.tolerance,
                                 "max_iterations": o.max_iterations, "dimensionality": "3D",
                                 "order": "Linear" if o.order == "LINEAR" else "Nonlinear",
                                 "nonlinear_method": "P_DELTA", "rigid_strategy": "RigidMember",
                                 "include_shear_deformation": True, "include_warping": False,
                                 "include_shear_center_coupling": False},
                     "load_cases": cases, "load_combinations": combos, "imperfection_cases": [], "unity_checks": []},
        "results": None,
    }


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: batch of parametric frames, sway imperfection, second-order P-Delta
# ---------------------------------------------------------------------------------------------
# A, Iy, Iz, J of a few catalogue sections (fers_core/sections/steel_sections.generated.json)
BATCH_SECTIONS = (
    (0.0078, 2.0e-05, 5.7e-05, 6.0e-07),     # HEB200-like column
    (0.0106, 3.9e-05, 1.13e-04, 1.0e-06),    # HEB240-like column
    HEB300[:4],
    (0.0039, 2.8e-06, 3.9e-05, 1.3e-07),     # IPE240-like beam
    (0.0054, 6.0e-06, 8.4e-05, 2.0e-07),     # IPE300-like beam
    (0.0073, 1.0e-05, 1.63e-04, 3.7e-07),    # IPE360-like beam
)
BATCH_SEED = 7
PORTAL_NODES = 45            # 3 x 3 plan x 5 levels


def portal_frame(bay: float, storey: float, col_sec: int, beam_sec: int, gravity: float, wind: float,
                 tolerance: float = 1e-8, order: str = "NONLINEAR") -> wire.ModelArrays:
    """One 2x2-bay, 4-storey 3D frame (45 nodes, 270 DOF) with a 1/200 sway imperfection about Z and a
    single combination (gravity UDL on the beams + lateral nodal loads), P-Delta consistent."""
    px = pz = 3
    py = 5
    nn = px * pz * py
    i, k, j = np.meshgrid(np.arange(px), np.arange(pz), np.arange(py), indexing="ij")
    idx = (i + px * (k + pz * j)).reshape(-1)
    xyz = np.zeros((nn, 3))
    xyz[idx, 0] = (i * bay).reshape(-1); xyz[idx, 1] = (j * storey).reshape(-1); xyz[idx, 2] = (k * bay).reshape(-1)

    def nid(ii, kk, jj):
        return ii + px * (kk + pz * jj)
    ii, kk, jj = np.meshgrid(np.arange(px), np.arange(pz), np.arange(py - 1), indexing="ij")
    c1, c2 = nid(ii, kk, jj).reshape(-1), nid(ii, kk, jj + 1).reshape(-1)
    ii, kk, jj = np.meshgrid(np.arange(px - 1), np.arange(pz), np.arange(1, py), indexing="ij")
    bx1, bx2 = nid(ii, kk, jj).reshape(-1), nid(ii + 1, kk, jj).reshape(-1)
    ii, kk, jj = np.meshgrid(np.arange(px), np.arange(pz - 1), np.arange(1, py), indexing="ij")
    bz1, bz2 = nid(ii, kk, jj).reshape(-1), nid(ii, kk + 1, jj).reshape(-1)
    n1 = np.concatenate([c1, bx1, bz1]); n2 = np.concatenate([c2, bx2, bz2])
    is_beam = np.concatenate([np.zeros(len(c1), bool), np.ones(len(bx1) + len(bz1), bool)])
    o = np.lexsort((n2, n1))
    n1, n2, is_beam = n1[o], n2[o], is_beam[o]
    nm = len(n1)
    bc = np.zeros((nn, 6), np.uint8); bc[: px * pz] = wire.BC_FIXED
    sup = np.full(nn, -1, np.int64); sup[: px * pz] = 1
    prop = np.array([[E_STEEL * s[0], E_STEEL * s[1], E_STEEL * s[2], G_STEEL * s[3], 0, 0, 0, 0] for s in BATCH_SECTIONS])
    m_prop = np.where(is_beam, beam_sec, col_sec).astype(np.int32)
    lc = wire.LoadCaseArrays(id=1, name="Design")
    beams = np.nonzero(is_beam)[0].astype(np.int32)
    lc.dl_member = beams
    q = -20e3 * gravity
    lc.dl_data = np.tile(np.array([0.0, q, 0.0, 0.0, q, 0.0, 0.0, 1.0]), (len(beams), 1))
    lvl = idx // (px * pz)
    top = idx[(lvl >= 1) & (idx % px == 0)]
    lc.nod_node = top.astype(np.int32)
    lc.nod_vec = np.tile(np.array([wind, 0, 0.3 * wind, 0, 0, 0.0]), (len(top), 1))
    opts = wire.Options(order=order, nonlinear_method="P_DELTA", pdelta_formulation="CONSISTENT", tolerance=tolerance,
                        max_iterations=30, solve_loadcases=False)
    imp = [{"imperfection_case_id": 1, "load_combinations": [1],
            "sway_imperfections": [{"magnitude": 1.0 / 200.0, "axis": [0, 0, 1], "height_direction": [0, 1, 0],
                                    "reference_point": [0, 0, 0]}], "translation_imperfections": []}]
    return wire.ModelArrays(
        node_ids=np.arange(1, nn + 1, dtype=np.int64), xyz=xyz, node_support_id=sup, bc_code=bc, bc_k=np.zeros((nn, 6)),
        m_ids=np.arange(1, nm + 1, dtype=np.int64), m_n1=n1.astype(np.int32), m_n2=n2.astype(np.int32),
        m_type=np.zeros(nm, np.int32), m_roll=np.zeros(nm), m_prop=m_prop, m_rel=np.full(nm, -1, np.int32), prop=prop,
        prop_rho_a=np.array([RHO_STEEL * s[0] for s in BATCH_SECTIONS]), rel=np.zeros((0, wire.REL_STRIDE)),
        p_ids=np.zeros(0, np.int64), p_nodes=np.zeros((0, 3), np.int32), p_data=np.zeros((0, wire.NPDATA)),
        p_flags=np.zeros(0, np.int32), load_cases=[lc], combo_ids=np.array([1], np.int64), combo_names=["ULS"],
        combo_factors=np.ones((1, 1)), imperfections=imp, member_sets={}, options=opts)


def portal_params(n_frames: int, seed: int = BATCH_SEED) -> list:
    """Parameters of the batch: bay 4-8 m, storey 3-4 m, section indices, gravity factor 0.5-1.5."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_frames):
        out.append(dict(bay=float(rng.uniform(4.0, 8.0)), storey=float(rng.uniform(3.0, 4.0)),
                        col_sec=int(rng.integers(0, 3)), beam_sec=int(rng.integers(3, 6)),
                        gravity=float(rng.uniform(0.5, 1.5)), wind=float(rng.uniform(2e3, 8e3))))
    return out


def concat_models(models: list) -> wire.ModelArrays:
    """Concatenate equally-shaped models into one block-diagonal model (frame-major numbering).
    All models must share the section table, options, load-case layout and imperfection case."""
    m0 = models[0]
    nn, nm = m0.n_nodes, m0.n_members
    nf = len(models)

    def cat(name, node_offset=False, member_offset=False):
        parts = []
        for f, m in enumerate(models):
            a = getattr(m, name)
            parts.append(a + (f * nn if node_offset else 0) if node_offset else a)
        return np.concatenate(parts, 0)
    cases = []
    for li, lc0 in enumerate(m0.load_cases):
        lc = wire.LoadCaseArrays(id=lc0.id, name=lc0.name)
        lc.nod_node = np.concatenate([m.load_cases[li].nod_node + f * nn for f, m in enumerate(models)]).astype(np.int32)
        lc.nod_vec = np.concatenate([m.load_cases[li].nod_vec for m in models], 0)
        lc.dl_member = np.concatenate([m.load_cases[li].dl_member + f * nm for f, m in enumerate(models)]).astype(np.int32)
        lc.dl_data = np.concatenate([m.load_cases[li].dl_data for m in models], 0)
        lc.pl_member = np.concatenate([m.load_cases[li].pl_member + f * nm for f, m in enumerate(models)]).astype(np.int32)
        lc.pl_data = np.concatenate([m.load_cases[li].pl_data for m in models], 0)
        cases.append(lc)
    return wire.ModelArrays(
        node_ids=np.arange(1, nf * nn + 1, dtype=np.int64), xyz=cat("xyz"),
        node_support_id=cat("node_support_id"), bc_code=cat("bc_code"), bc_k=cat("bc_k"),
        m_ids=np.arange(1, nf * nm + 1, dtype=np.int64), m_n1=cat("m_n1", node_offset=True).astype(np.int32),
        m_n2=cat("m_n2", node_offset=True).astype(np.int32), m_type=cat("m_type"), m_roll=cat("m_roll"),
        m_prop=cat("m_prop"), m_rel=cat("m_rel"), prop=m0.prop.copy(), prop_rho_a=m0.prop_rho_a.copy(), rel=m0.rel.copy(),
        p_ids=np.zeros(0, np.int64), p_nodes=np.zeros((0, 3), np.int32), p_data=np.zeros((0, wire.NPDATA)),
        p_flags=np.zeros(0, np.int32), load_cases=cases, combo_ids=m0.combo_ids.copy(), combo_names=list(m0.combo_names),
        combo_factors=m0.combo_factors.copy(), imperfections=list(m0.imperfections), member_sets={}, options=m0.options)


def portal_batch(n_frames: int, seed: int = BATCH_SEED, tolerance: float = 1e-8):
    """(concatenated model, list of per-frame models) for configs[4]."""
    frames = [portal_frame(tolerance=tolerance, **p) for p in portal_params(n_frames, seed)]
    return concat_models(frames), frames
