"""Result arrays and their serialisation to the reference's ``ResultsBundle`` wire form.

Output contract: ``fers_core/types/pydantic_models.py`` -- ``ResultsBundle`` 2171-2185,
``Results`` 1875-1890, ``MemberResult`` 1594-1635, ``NodeForces`` 642-649,
``NodeDisplacement`` 632-639, ``ReactionNodeResult`` 980-983, ``PlateResult`` 1743-1749.
Conventions: result sets keyed by NAME, entities by decimal id string;
reactions for every supported node with all six components; member end forces in the stiffness
(nodal) sign convention, global and local; interior section forces at x_frac 0.25/0.5/0.75.
"""
from __future__ import annotations

from collections.abc import Mapping
from dataclasses import dataclass, field

import numpy as np

from .hostlogic import DEFLECTED_STATIONS, member_deflected_shape

SECTION_FRACS = (0.25, 0.5, 0.75)
_STATIONS = tuple(float(x) for x in np.linspace(0.0, 1.0, DEFLECTED_STATIONS))
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


class _Context:
    """What every column of a result bundle shares: keys, supported nodes, warnings, plate rows per plate id."""

    def __init__(self, model, res: ResultArrays):
        self.model, self.res, self.plan = model, res, res.plan
        self.n_real = model.n_real
        self.sup_nodes = np.nonzero(model.node_support_id[:self.n_real] >= 0)[0]
        self.node_keys = [str(int(i)) for i in model.node_ids]
        self.mem_keys = [str(int(i)) for i in model.m_ids]
        self.warp_node = model.warp_node if len(model.warp_node) == self.n_real else np.full(self.n_real, -1)
        self.warn = [{"code": c, "message": m} for c, m in (list(model.warnings) + list(res.warnings))]
        self.plate_rows: dict = {}                       # plate id key -> TRI3 rows (a split QUAD4 has two)
        for e in range(model.n_plates):
            self.plate_rows.setdefault(str(int(model.p_ids[e])), []).append(e)


class _Column:
    """One load case / combination: every wire sub-document is built from the result arrays on request."""

    def __init__(self, ctx: _Context, col: int):
        self.c, self.col = ctx, col
        self._shape = None

    def deflected_shape(self) -> np.ndarray:
        """[Nm, stations, 3] global displacements along every member of this column (built once, on first use)."""
        if self._shape is None:
            self._shape = member_deflected_shape(self.c.model, self.c.res.U[self.col])
        return self._shape

    def _warp(self, arr, i) -> float:
        """Value at the warping DOF of real node i (DOF 0 of its auxiliary node), 0 without one."""
        a = int(self.c.warp_node[i])
        return float(arr[self.col, a, 0]) if a >= 0 else 0.0

    def node_disp(self, i: int) -> dict:
        d = _disp(self.c.res.U[self.col, i])
        d["warp"] = self._warp(self.c.res.U, i)
        return d

    def reaction(self, i: int) -> dict:
        model, res = self.c.model, self.c.res
        return {
            "location": {"X": float(model.xyz[i, 0]), "Y": float(model.xyz[i, 1]), "Z": float(model.xyz[i, 2])},
            "nodal_forces": _forces(res.R[self.col, i], self._warp(res.R, i)),
            "support_id": int(model.node_support_id[i]),
        }

    def member(self, m: int) -> dict:
        res, col = self.c.res, self.col
        fl, fg, ul = res.f_local[col, m], res.f_global[col, m], res.u_local[col, m]
        sec = [{"x_frac": float(x), "forces": _forces(res.sec[col, m, si])} for si, x in enumerate(SECTION_FRACS)]
        series = ([{"x_frac": 0.0, "forces": _forces(fl[0:6])}] + sec
                  + [{"x_frac": 1.0, "forces": _forces(-fl[6:12])}])
        b1, b2 = (float(res.bw[col, m, 0]), float(res.bw[col, m, 1])) if res.bw is not None else (0.0, 0.0)
        extra = {}
        if self.c.model.options.include_member_deflected_shape:
            # whole table when a document is being built eagerly, this member alone under the lazy view
            shape = self._shape[m] if self._shape is not None else member_deflected_shape(self.c.model, res.U[col], [m])[0]
            extra["member_displacements"] = [{"x_frac": float(x), "displacement": [float(v) for v in d]}
                                             for x, d in zip(_STATIONS, shape)]
        return {
            "start_node_forces": _forces(fg[0:6], b1), "end_node_forces": _forces(fg[6:12], b2),
            "maximums": _forces(res.gmax[col, m]), "minimums": _forces(res.gmin[col, m]),
            "local_start_forces": _forces(fl[0:6], b1), "local_end_forces": _forces(fl[6:12], b2),
            "local_maximums": _forces(res.lmax[col, m]), "local_minimums": _forces(res.lmin[col, m]),
            "local_displacement_start_node": _disp(ul[0:6]), "local_displacement_end_node": _disp(ul[6:12]),
            "section_forces": sec, "internal_force_series": series, **extra,
        }

    def plate(self, key: str) -> dict:
        model, res, col = self.c.model, self.c.res, self.col
        rows = self.c.plate_rows[key]
        e0 = rows[0]
        nodal: dict = {}
        for e in rows:                                   # second half of a split QUAD4: corner forces accumulate
            for k, n in enumerate(model.p_nodes[e]):
                f = _forces(res.p_nodal[col, e, k])
                nk = self.c.node_keys[int(n)]
                if nk in nodal:
                    for comp in f:
                        nodal[nk][comp] += f[comp]
                else:
                    nodal[nk] = f
        c = res.p_centroid[e0]
        return {
            "plate_id": int(model.p_ids[e0]),
            "centroid": {"X": float(c[0]), "Y": float(c[1]), "Z": float(c[2])},
            "centroid_displacement_global": _disp(res.p_cdg[col, e0]),
            "centroid_displacement_local": _disp(res.p_cdl[col, e0]),
            "nodal_forces_global": nodal,
            "resultants": {k: float(v) for k, v in zip(_PR, res.p_res[col, e0])},
        }

    def summary(self) -> dict:
        s = {"total_displacements": self.c.n_real, "total_member_forces": self.c.model.n_members,
             "total_reaction_forces": int(len(self.c.sup_nodes))}
        if self.c.model.n_plates:
            s["total_plate_forces"] = len(self.c.plate_rows)
        return s

    def diagnostics(self) -> dict:
        res, col = self.c.res, self.col
        return {
            "analysis_method": res.method,
            "converged": bool(res.converged[col]),
            "iterations": None if res.iterations[col] < 0 else int(res.iterations[col]),
            "residual_norm": None if res.iterations[col] < 0 else float(res.residual[col]),
            "solve_time_ms": None if res.solve_ms is None else float(res.solve_ms),
        }

    def document(self, mapping=None) -> dict:
        """The wire ``Results`` document of this column; ``mapping(keys, items, getter)`` builds the per-entity maps
        (plain dicts by default, :class:`LazyMap` for :func:`lazy_bundle`)."""
        c, plan, col = self.c, self.c.plan, self.col
        if mapping is None:
            if c.model.options.include_member_deflected_shape:
                self.deflected_shape()

            def mapping(keys, items, get):
                return {k: get(i) for k, i in zip(keys, items)}
        nodes = range(c.n_real)
        r = {
            "name": plan.names[col],
            "result_type": {plan.kinds[col]: int(plan.ids[col])},
            "displacement_nodes": mapping([c.node_keys[i] for i in nodes], nodes, self.node_disp),
            "reaction_nodes": mapping([c.node_keys[i] for i in c.sup_nodes], c.sup_nodes, self.reaction),
            "member_results": mapping(c.mem_keys, range(c.model.n_members), self.member),
            "summary": self.summary(),
            "errors_and_warnings": {"errors": [], "warnings": c.warn},
            "solver_diagnostics": self.diagnostics(),
        }
        if c.model.n_plates:
            keys = list(c.plate_rows)
            r["plate_results"] = mapping(keys, keys, self.plate)
        return r


def results_to_wire(model, res: ResultArrays) -> dict:
    """``ResultsBundle`` dict (pydantic_models.py:2171-2185) from result arrays."""
    plan = res.plan
    ctx = _Context(model, res)
    bundle = {"loadcases": {}, "loadcombinations": {},
              "unity_check_results": res.unity.results() if res.unity is not None else []}
    for col in range(plan.n_rhs):
        bundle["loadcases" if plan.kinds[col] == "Loadcase" else "loadcombinations"][plan.names[col]] = _Column(ctx, col).document()
    return bundle


class LazyMap(Mapping):
    """Read-only mapping whose values are built on access (one node / member / plate / column at a time)."""

    def __init__(self, keys, items, get):
        self._index = {k: i for k, i in zip(keys, items)}
        self._get = get

    def __getitem__(self, key):
        return self._get(self._index[key])

    def __iter__(self):
        return iter(self._index)

    def __len__(self):
        return len(self._index)

    def to_dict(self) -> dict:
        return {k: _plain(v) for k, v in self.items()}


def _plain(v):
    if isinstance(v, LazyMap):
        return v.to_dict()
    if isinstance(v, dict):
        return {k: _plain(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_plain(x) for x in v]
    return v


def lazy_bundle(model, res: ResultArrays) -> LazyMap:
    """The same ``ResultsBundle`` document as :func:`results_to_wire`, as nested read-only mappings over the
    result arrays: ``bundle["loadcombinations"][name]["member_results"]["17"]["maximums"]["mz"]`` touches one
    member of one column.  configs[3] would be ~70 GB as JSON; the arrays are 10 GB and a
    view costs nothing until it is read.  ``.to_dict()`` materialises any level."""
    plan = res.plan
    ctx = _Context(model, res)

    def column(col):
        return _Column(ctx, col).document(mapping=LazyMap)
    groups = {}
    for kind, name in (("Loadcase", "loadcases"), ("Loadcombination", "loadcombinations")):
        cols = [c for c in range(plan.n_rhs) if plan.kinds[c] == kind]
        groups[name] = LazyMap([plan.names[c] for c in cols], cols, column)
    unity = res.unity.results() if res.unity is not None else []
    top = {"loadcases": groups["loadcases"], "loadcombinations": groups["loadcombinations"], "unity_check_results": unity}
    return LazyMap(list(top), list(top), top.__getitem__)


# ---- JSON text straight from the arrays (the boundary of fers_core/fers/fers.py:300 returns a string) -------------
_SLOT, _ISLOT = "\x01", "\x02"           # number / integer slots of fers_b200_json_records templates
_F_T = "{" + ",".join(f'"{k}":{_SLOT}' for k in _F + ("bw",)) + "}"
_D_T = "{" + ",".join(f'"{k}":{_SLOT}' for k in _D + ("warp",)) + "}"
_MEMBER_T = ("{" + ",".join(f'"{k}":{_F_T}' for k in (
    "start_node_forces", "end_node_forces", "maximums", "minimums", "local_start_forces", "local_end_forces",
    "local_maximums", "local_minimums"))
    + f',"local_displacement_start_node":{_D_T},"local_displacement_end_node":{_D_T}'
    + ',"section_forces":[' + ",".join(f'{{"x_frac":{float(x)!r},"forces":{_F_T}}}' for x in SECTION_FRACS) + "]"
    + ',"internal_force_series":[' + ",".join(f'{{"x_frac":{float(x)!r},"forces":{_F_T}}}' for x in (0.0,) + SECTION_FRACS + (1.0,))
    + "]}")
_SHAPE_T = (',"member_displacements":[' + ",".join(f'{{"x_frac":{float(x)!r},"displacement":[{_SLOT},{_SLOT},{_SLOT}]}}' for x in _STATIONS) + "]}")
_REACTION_T = f'{{"location":{{"X":{_SLOT},"Y":{_SLOT},"Z":{_SLOT}}},"nodal_forces":{_F_T},"support_id":{_ISLOT}}}'


def _json_records(template: str, keys_blob: bytes, values: np.ndarray) -> str:
    """One '{"key": record, ...}' map through the library's writer (shortest round-trip numbers)."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    vals = np.ascontiguousarray(values, np.float64)
    n, slots = vals.shape
    t = template.encode()
    bound = lib.fers_b200_json_records_bound(len(t), slots, n, len(keys_blob))
    out = np.empty(bound, np.uint8)
    written = C.c_size_t()
    _lib.check(lib.fers_b200_json_records(t, len(t), slots, n, keys_blob, vals.ctypes.data, out.ctypes.data, bound,
                                          C.byref(written)), "json_records")
    return out[: written.value].tobytes().decode()


def results_to_json(model, res: ResultArrays) -> str:
    """The ``ResultsBundle`` document of :func:`results_to_wire` as JSON text, written from the arrays without the
    intermediate dicts: ``json.loads(results_to_json(m, r)) == results_to_wire(m, r)``."""
    import json
    plan = res.plan
    ctx = _Context(model, res)
    n_real, nm = ctx.n_real, model.n_members
    blob = lambda keys: ("\0".join(keys) + "\0").encode() if keys else b""     # noqa: E731
    node_blob = blob(ctx.node_keys[:n_real])
    sup_blob = blob([ctx.node_keys[i] for i in ctx.sup_nodes])
    mem_blob = blob(ctx.mem_keys)
    wn = np.asarray(ctx.warp_node[:n_real], np.int64)
    has_w = wn >= 0

    def warp(arr, col, nodes):
        w = np.zeros(len(nodes))
        sel = has_w[nodes]
        w[sel] = arr[col, wn[nodes][sel], 0]
        return w[:, None]
    all_nodes = np.arange(n_real)
    groups = {"loadcases": [], "loadcombinations": []}
    for col in range(plan.n_rhs):
        v = _Column(ctx, col)
        disp = _json_records(_D_T, node_blob, np.hstack([res.U[col, :n_real], warp(res.U, col, all_nodes)])) if n_real else "{}"
        sup = ctx.sup_nodes
        if len(sup):
            reac = _json_records(_REACTION_T, sup_blob,
                                 np.hstack([model.xyz[sup], res.R[col, sup], warp(res.R, col, sup),
                                            model.node_support_id[sup][:, None].astype(np.float64)]))
        else:
            reac = "{}"
        if nm:
            fl, fg, ul, sec = res.f_local[col], res.f_global[col], res.u_local[col], res.sec[col]
            z = np.zeros((nm, 1))
            bw = res.bw[col] if res.bw is not None else np.zeros((nm, 2))
            b1, b2 = bw[:, :1], bw[:, 1:2]
            secs = [x for si in range(len(SECTION_FRACS)) for x in (sec[:, si], z)]
            mv = np.hstack([fg[:, 0:6], b1, fg[:, 6:12], b2, res.gmax[col], z, res.gmin[col], z,
                            fl[:, 0:6], b1, fl[:, 6:12], b2, res.lmax[col], z, res.lmin[col], z,
                            ul[:, 0:6], z, ul[:, 6:12], z] + secs + [fl[:, 0:6], z] + secs + [-fl[:, 6:12], z])
            template = _MEMBER_T
            if model.options.include_member_deflected_shape:
                template = _MEMBER_T[:-1] + _SHAPE_T
                mv = np.hstack([mv, v.deflected_shape().reshape(nm, -1)])
            mem = _json_records(template, mem_blob, mv)
        else:
            mem = "{}"
        parts = [f'"name":{json.dumps(plan.names[col])}',
                 f'"result_type":{json.dumps({plan.kinds[col]: int(plan.ids[col])})}',
                 f'"displacement_nodes":{disp}', f'"reaction_nodes":{reac}', f'"member_results":{mem}',
                 f'"summary":{json.dumps(v.summary())}',
                 f'"errors_and_warnings":{json.dumps({"errors": [], "warnings": ctx.warn})}',
                 f'"solver_diagnostics":{json.dumps(v.diagnostics())}']
        if model.n_plates:
            parts.append('"plate_results":' + json.dumps({k: v.plate(k) for k in ctx.plate_rows}))
        groups["loadcases" if plan.kinds[col] == "Loadcase" else "loadcombinations"].append(
            json.dumps(plan.names[col]) + ":{" + ",".join(parts) + "}")
    unity = res.unity.results() if res.unity is not None else []
    return ('{"loadcases":{' + ",".join(groups["loadcases"]) + '},"loadcombinations":{' + ",".join(groups["loadcombinations"])
            + '},"unity_check_results":' + json.dumps(unity) + "}")
