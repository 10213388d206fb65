"""Result arrays and their serialisation to the reference's ``ResultsBundle`` wire form.

Output contract: ``fers_core/types/pydantic_models.py`` -- ``ResultsBundle`` 2171-2185,
``Results`` 1875-1890, ``MemberResult`` 1594-1635, ``NodeForces`` 642-649,
``NodeDisplacement`` 632-639, ``ReactionNodeResult`` 980-983, ``PlateResult`` 1743-1749.
Conventions: result sets keyed by NAME, entities by decimal id string;
reactions for every supported node with all six components; member end forces in the stiffness
(nodal) sign convention, global and local; interior section forces at x_frac 0.25/0.5/0.75.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

SECTION_FRACS = (0.25, 0.5, 0.75)
_F = ("fx", "fy", "fz", "mx", "my", "mz")
_D = ("dx", "dy", "dz", "rx", "ry", "rz")
_PR = ("nx", "ny", "nxy", "mx", "my", "mxy", "qx", "qy")


@dataclass
class ResultArrays:
    """All result sets of one document as dense arrays; leading axis = rhs column."""
    plan: object
    U: np.ndarray          # [nrhs,Nn,6]   global node displacements
    R: np.ndarray          # [nrhs,Nn,6]   reactions (zero rows for unsupported nodes)
    f_local: np.ndarray    # [nrhs,Nm,12]  member end forces, local frame, nodal convention
    f_global: np.ndarray   # [nrhs,Nm,12]
    u_local: np.ndarray    # [nrhs,Nm,12]  member end displacements in the local frame
    sec: np.ndarray        # [nrhs,Nm,3,6] internal forces at SECTION_FRACS, local frame
    lmax: np.ndarray       # [nrhs,Nm,6]
    lmin: np.ndarray
    gmax: np.ndarray
    gmin: np.ndarray
    p_res: np.ndarray      # [nrhs,Npl,8]  nx ny nxy mx my mxy qx qy (element local frame)
    p_nodal: np.ndarray    # [nrhs,Npl,3,6] corner forces, global
    p_cdg: np.ndarray      # [nrhs,Npl,6]  centroid displacement global
    p_cdl: np.ndarray      # [nrhs,Npl,6]  centroid displacement local
    p_centroid: np.ndarray  # [Npl,3]
    iterations: np.ndarray  # [nrhs] int (-1: linear)
    converged: np.ndarray   # [nrhs] bool
    residual: np.ndarray    # [nrhs]
    method: str = "Linear"
    solve_ms: float | None = None
    warnings: list = field(default_factory=list)
    bw: np.ndarray = None            # [nrhs,Nm,2] bimoment at the member ends (nodal force at the warping DOF)
    unity: object = None             # unity.UnityAccumulator while the analysis runs (None: no checks)
    unity_xyz: object = None

    @classmethod
    def empty(cls, model, plan) -> "ResultArrays":
        n, nn, nm, npl = plan.n_rhs, model.n_nodes, model.n_members, model.n_plates
        ns = len(SECTION_FRACS)
        z = np.zeros
        return cls(plan=plan, U=z((n, nn, 6)), R=z((n, nn, 6)), f_local=z((n, nm, 12)), f_global=z((n, nm, 12)),
                   u_local=z((n, nm, 12)), sec=z((n, nm, ns, 6)), lmax=z((n, nm, 6)), lmin=z((n, nm, 6)),
                   gmax=z((n, nm, 6)), gmin=z((n, nm, 6)), p_res=z((n, npl, 8)), p_nodal=z((n, npl, 3, 6)),
                   p_cdg=z((n, npl, 6)), p_cdl=z((n, npl, 6)), p_centroid=z((npl, 3)),
                   iterations=np.full(n, -1, np.int64), converged=np.ones(n, bool), residual=z(n), bw=z((n, nm, 2)))


def _forces(v, bw=0.0) -> dict:
    d = {k: float(x) for k, x in zip(_F, v)}
    d["bw"] = float(bw)
    return d


def _disp(v) -> dict:
    d = {k: float(x) for k, x in zip(_D, v)}
    d["warp"] = 0.0
    return d


def results_to_wire(model, res: ResultArrays) -> dict:
    """``ResultsBundle`` dict (pydantic_models.py:2171-2185) from result arrays."""
    plan = res.plan
    bundle = {"loadcases": {}, "loadcombinations": {},
              "unity_check_results": res.unity.results() if res.unity is not None else []}
    n_real = model.n_real
    sup_nodes = np.nonzero(model.node_support_id[:n_real] >= 0)[0]
    node_keys = [str(int(i)) for i in model.node_ids]
    warp_node = model.warp_node if len(model.warp_node) == n_real else np.full(n_real, -1)

    def warp_of(arr, col, i):
        """Value at the warping DOF of real node i (DOF 0 of its auxiliary node), 0 without one."""
        a = int(warp_node[i])
        return float(arr[col, a, 0]) if a >= 0 else 0.0
    mem_keys = [str(int(i)) for i in model.m_ids]
    warn = [{"code": c, "message": m} for c, m in (list(model.warnings) + list(res.warnings))]
    for col in range(plan.n_rhs):
        disp = {}
        for i in range(n_real):
            disp[node_keys[i]] = _disp(res.U[col, i])
            disp[node_keys[i]]["warp"] = warp_of(res.U, col, i)
        reac = {}
        for i in sup_nodes:
            reac[node_keys[i]] = {
                "location": {"X": float(model.xyz[i, 0]), "Y": float(model.xyz[i, 1]), "Z": float(model.xyz[i, 2])},
                "nodal_forces": _forces(res.R[col, i], warp_of(res.R, col, i)),
                "support_id": int(model.node_support_id[i]),
            }
        members = {}
        for m in range(model.n_members):
            fl, fg, ul = res.f_local[col, m], res.f_global[col, m], res.u_local[col, m]
            sec = [{"x_frac": float(x), "forces": _forces(res.sec[col, m, si])} for si, x in enumerate(SECTION_FRACS)]
            series = ([{"x_frac": 0.0, "forces": _forces(fl[0:6])}] + sec
                      + [{"x_frac": 1.0, "forces": _forces(-fl[6:12])}])
            b1, b2 = (float(res.bw[col, m, 0]), float(res.bw[col, m, 1])) if res.bw is not None else (0.0, 0.0)
            members[mem_keys[m]] = {
                "start_node_forces": _forces(fg[0:6], b1), "end_node_forces": _forces(fg[6:12], b2),
                "maximums": _forces(res.gmax[col, m]), "minimums": _forces(res.gmin[col, m]),
                "local_start_forces": _forces(fl[0:6], b1), "local_end_forces": _forces(fl[6:12], b2),
                "local_maximums": _forces(res.lmax[col, m]), "local_minimums": _forces(res.lmin[col, m]),
                "local_displacement_start_node": _disp(ul[0:6]), "local_displacement_end_node": _disp(ul[6:12]),
                "section_forces": sec, "internal_force_series": series,
            }
        r = {
            "name": plan.names[col],
            "result_type": {plan.kinds[col]: int(plan.ids[col])},
            "displacement_nodes": disp, "reaction_nodes": reac, "member_results": members,
            "summary": {"total_displacements": n_real, "total_member_forces": model.n_members,
                        "total_reaction_forces": int(len(sup_nodes))},
            "errors_and_warnings": {"errors": [], "warnings": warn},
            "solver_diagnostics": {
                "analysis_method": res.method,
                "converged": bool(res.converged[col]),
                "iterations": None if res.iterations[col] < 0 else int(res.iterations[col]),
                "residual_norm": None if res.iterations[col] < 0 else float(res.residual[col]),
                "solve_time_ms": None if res.solve_ms is None else float(res.solve_ms),
            },
        }
        if model.n_plates:
            plates = {}
            for e in range(model.n_plates):
                key = str(int(model.p_ids[e]))
                nod = {node_keys[int(n)]: _forces(res.p_nodal[col, e, k]) for k, n in enumerate(model.p_nodes[e])}
                if key in plates:   # second half of a split QUAD4: accumulate corner forces
                    for nk, f in nod.items():
                        if nk in plates[key]["nodal_forces_global"]:
                            for c in f:
                                plates[key]["nodal_forces_global"][nk][c] += f[c]
                        else:
                            plates[key]["nodal_forces_global"][nk] = f
                    continue
                c = res.p_centroid[e]
                plates[key] = {
                    "plate_id": int(model.p_ids[e]),
                    "centroid": {"X": float(c[0]), "Y": float(c[1]), "Z": float(c[2])},
                    "centroid_displacement_global": _disp(res.p_cdg[col, e]),
                    "centroid_displacement_local": _disp(res.p_cdl[col, e]),
                    "nodal_forces_global": nod,
                    "resultants": {k: float(v) for k, v in zip(_PR, res.p_res[col, e])},
                }
            r["plate_results"] = plates
            r["summary"]["total_plate_forces"] = len(plates)
        bundle["loadcases" if plan.kinds[col] == "Loadcase" else "loadcombinations"][plan.names[col]] = r
    return bundle
