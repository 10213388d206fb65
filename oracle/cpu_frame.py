"""CPU ORACLE / CPU BASELINE driver for plain 3D frames -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python side of oracle/cpu_frame.cpp (C++ + OpenMP + the OpenBLAS inside the scipy wheel): the configs[3] subset of
oracle/fers_oracle.py restated for speed, so that the CPU arm of bench.py can run BASELINE.json configs[3] ITSELF
(692 106 DOF x 64 combinations) on the host cores instead of a 20x smaller sample.  Same phases as one pass of the GPU hot
path: load vectors -> element stiffness + assembly + boundary conditions -> one factorisation shared by all right-hand
sides (fers_core/types/pydantic_models.py:1216-1219) -> 64-RHS substitution -> member forces and reactions.

The elimination tree (ordering + symbolic analysis) is the product's host-side analysis phase, ``chol.CholPlan`` -- plain
integer arrays, no arithmetic of the path; everything numeric happens in libfers_cpu.so.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this module.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import time

import numpy as np

from fers_core_b200 import chol, wire
from . import build as _build

_LIB = None


class FrontPlan(C.Structure):
    _fields_ = [("n_nodes", C.c_int), ("n_fronts", C.c_int), ("n_levels", C.c_int), ("reserved", C.c_int),
                ("fronts", C.c_void_p), ("nodes", C.c_void_p), ("child", C.c_void_p), ("inv_off", C.c_void_p),
                ("inv", C.c_void_p), ("level_ptr", C.c_void_p), ("level_fronts", C.c_void_p), ("kscatter", C.c_void_p),
                ("kscatter_ptr", C.c_void_p), ("pos", C.c_void_p),
                ("factor_doubles", C.c_longlong), ("schur_doubles0", C.c_longlong), ("schur_doubles1", C.c_longlong),
                ("rhs_rows", C.c_longlong)]


def openblas_path() -> str:
    import scipy
    hits = sorted(glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so*")))
    if not hits:
        raise RuntimeError("the OpenBLAS shipped with scipy was not found (scipy.libs/libscipy_openblas*.so)")
    return os.path.abspath(hits[0])


def load():
    global _LIB
    if _LIB is None:
        lib = C.CDLL(_build.build())
        lib.fers_cpu_last_error.restype = C.c_char_p
        if lib.fers_cpu_init(openblas_path().encode()) != 0:
            raise RuntimeError(lib.fers_cpu_last_error().decode())
        _LIB = lib
    return _LIB


def _ck(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what}: {load().fers_cpu_last_error().decode()}")


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def threads() -> int:
    return int(load().fers_cpu_threads())


def set_threads(n: int) -> None:
    load().fers_cpu_set_threads(int(n))


class CpuFrame:
    """One plain frame model on the host: pattern, property table, boundary conditions."""

    def __init__(self, model: wire.ModelArrays):
        self.lib = load()
        m = self.model = model
        if m.n_plates or len(m.w_member) or np.any(m.m_rel >= 0) or np.any(~np.isin(m.m_type, (wire.MT_NORMAL, wire.MT_TRUSS))):
            raise ValueError("cpu_frame handles Normal / Truss members without modifier records only (the configs[3] subset)")
        if np.any(m.m_prop < 0):
            raise ValueError("cpu_frame: every member needs a section")
        self.xyz = np.ascontiguousarray(m.xyz, np.float64)
        self.n1 = np.ascontiguousarray(m.m_n1, np.int32); self.n2 = np.ascontiguousarray(m.m_n2, np.int32)
        self.m_prop = np.ascontiguousarray(m.m_prop, np.int32); self.m_type = np.ascontiguousarray(m.m_type, np.int32)
        self.roll = np.ascontiguousarray(m.m_roll, np.float64)
        self.prop = np.ascontiguousarray(m.prop, np.float64)
        self.rowptr, self.colidx = chol.host_pattern(m.n_nodes, self.n1, self.n2)
        self.rowptr = np.ascontiguousarray(self.rowptr, np.int32); self.colidx = np.ascontiguousarray(self.colidx, np.int32)
        self.nnzb = int(len(self.colidx))
        self.diag_slot = np.empty(m.n_nodes, np.int32)
        for_rows = np.repeat(np.arange(m.n_nodes, dtype=np.int32), np.diff(self.rowptr))
        self.diag_slot[for_rows[self.colidx == for_rows]] = np.nonzero(self.colidx == for_rows)[0].astype(np.int32)
        self.vals = np.empty(self.nnzb * 36, np.float64)
        self.plan = None
        self.timings: dict = {}

    # ---- stages (names follow fers_core_b200.engine.Engine) --------------------------------------------------------
    def assemble(self) -> None:
        m = self.model
        _ck(self.lib.fers_cpu_assemble(m.n_nodes, m.n_members, _p(self.xyz), _p(self.n1), _p(self.n2), _p(self.m_prop),
                                       _p(self.m_type), _p(self.roll), _p(self.prop), _p(self.rowptr), _p(self.colidx),
                                       _p(self.vals)), "assemble")
        # boundary conditions (oracle.fers_oracle.constrain): Fixed eliminated, springs on the diagonal, free DOFs without
        # stiffness restrained automatically
        code = m.bc_code.reshape(-1)
        self.kspring = np.where(code == wire.BC_SPRING, m.bc_k.reshape(-1), 0.0)
        diag = self.vals.reshape(-1, 36)[self.diag_slot][:, ::7].reshape(-1) + self.kspring
        self.fixed = np.ascontiguousarray(((code == wire.BC_FIXED) | (diag == 0.0)).astype(np.uint8))
        self.has_springs = bool(self.kspring.any())

    def load_cases(self):
        m = self.model
        nlc = len(m.load_cases)
        F = np.zeros((max(nlc, 1), m.n_dof)); peq = np.zeros((max(nlc, 1), m.n_members * 12))
        for li, lc in enumerate(m.load_cases):
            if len(lc.pp_plate):
                raise ValueError("cpu_frame: plate loads are outside the subset")
            nn_, nv = np.ascontiguousarray(lc.nod_node, np.int32), np.ascontiguousarray(lc.nod_vec, np.float64)
            dm, dd = np.ascontiguousarray(lc.dl_member, np.int32), np.ascontiguousarray(lc.dl_data, np.float64)
            pm, pd = np.ascontiguousarray(lc.pl_member, np.int32), np.ascontiguousarray(lc.pl_data, np.float64)
            _ck(self.lib.fers_cpu_load_case(m.n_nodes, m.n_members, _p(self.xyz), _p(self.n1), _p(self.n2), _p(self.m_prop),
                                            _p(self.m_type), _p(self.roll), _p(self.prop), len(nn_), _p(nn_), _p(nv), len(dm),
                                            _p(dm), _p(dd), len(pm), _p(pm), _p(pd), _p(F[li]), _p(peq[li])), "load_case")
        return F, peq

    def combine(self, per_case: np.ndarray, factors: np.ndarray) -> np.ndarray:
        nlc, rows = per_case.shape
        fac = np.ascontiguousarray(factors, np.float64)
        out = np.empty((rows, fac.shape[0]))
        _ck(self.lib.fers_cpu_combine(C.c_longlong(rows), nlc, fac.shape[0], _p(per_case), _p(fac), _p(out)), "combine")
        return out

    def make_plan(self):
        """Ordering + symbolic analysis (the product's host-side analysis phase) and the host buffers of the factor."""
        if self.plan is not None:
            return
        m = self.model
        restrained = (m.bc_code.reshape(-1, 6) == wire.BC_FIXED).all(axis=1)
        self._cplan = chol.CholPlan(self.rowptr, self.colidx, m.xyz, node_restrained=None if restrained.all() else restrained)
        ex = self._keep = self._cplan.export()
        info = self._cplan.info
        perm = self._cplan.perm()
        # perm lists the eliminated nodes in elimination order, then -1 (fully restrained nodes are left out of the tree)
        pos = np.full(m.n_nodes, -1, np.int32)
        n_elim = int((perm >= 0).sum())
        pos[perm[:n_elim]] = np.arange(n_elim, dtype=np.int32)
        self._pos = pos
        self._arr = {k: np.ascontiguousarray(ex[k]) for k in ("fronts", "nodes", "child", "inv_off", "inv", "level_fronts", "kscatter")}
        self._level_ptr = np.ascontiguousarray(ex["level_ptr"], np.int32)
        self._ks_ptr = np.ascontiguousarray(ex["kscatter_ptr"], np.int64)
        p = FrontPlan()
        p.n_nodes, p.n_fronts, p.n_levels = m.n_nodes, info.n_fronts, info.n_levels
        for k in ("fronts", "nodes", "child", "inv_off", "inv", "level_fronts", "kscatter"):
            setattr(p, k, self._arr[k].ctypes.data)
        p.level_ptr = self._level_ptr.ctypes.data; p.kscatter_ptr = self._ks_ptr.ctypes.data; p.pos = pos.ctypes.data
        p.factor_doubles, p.rhs_rows = info.factor_doubles, info.rhs_rows
        p.schur_doubles0, p.schur_doubles1 = info.schur_doubles
        self.plan = p
        self.factor = np.empty(info.factor_doubles + m.n_dof, np.float64)
        self.schur = [np.empty(max(int(s), 1), np.float64) for s in info.schur_doubles]
        self.flops = info.flops

    def factorize(self) -> None:
        self.make_plan()
        info = C.c_int(0)
        _ck(self.lib.fers_cpu_mf_factor(C.byref(self.plan), _p(self.vals), _p(self.diag_slot), _p(self.fixed),
                                        _p(self.kspring) if self.has_springs else None, _p(self.factor), _p(self.schur[0]),
                                        _p(self.schur[1]), C.byref(info)), "mf_factor")
        if info.value:
            raise RuntimeError(f"stiffness matrix is not positive definite (pivot {info.value} in elimination order)")

    def solve(self, f: np.ndarray) -> np.ndarray:
        nr = f.shape[1]
        work = np.empty((self.plan.rhs_rows + self.model.n_dof) * nr)
        u = np.empty_like(f)
        _ck(self.lib.fers_cpu_mf_solve(C.byref(self.plan), _p(self.factor), _p(self.fixed), nr, _p(f), _p(u), _p(work)), "mf_solve")
        return u

    def spmm(self, x: np.ndarray) -> np.ndarray:
        y = np.empty_like(x)
        _ck(self.lib.fers_cpu_spmm(self.model.n_nodes, x.shape[1], _p(self.rowptr), _p(self.colidx), _p(self.vals), _p(x), _p(y)), "spmm")
        return y

    def recover(self, u: np.ndarray, f: np.ndarray, peq: np.ndarray) -> dict:
        m = self.model
        nr = u.shape[1]
        r = self.spmm(u)
        r -= f
        r *= (m.node_support_id >= 0).repeat(6)[:, None]
        fl = np.empty((m.n_members, 12, nr)); fg = np.empty_like(fl); ul = np.empty_like(fl)
        _ck(self.lib.fers_cpu_recover_members(m.n_members, nr, _p(self.xyz), _p(self.n1), _p(self.n2), _p(self.m_prop),
                                              _p(self.m_type), _p(self.roll), _p(self.prop), _p(u), _p(peq), _p(fl), _p(fg), _p(ul)),
            "recover_members")
        return {"R": r, "f_local": fl, "f_global": fg, "u_local": ul}

    def run_linear(self, factors: np.ndarray) -> dict:
        """One pass of the linear hot path for the columns ``factors`` [ncols, Nlc] (twin of Engine.run_linear)."""
        t = [time.perf_counter()]

        def lap(name):
            t.append(time.perf_counter()); self.timings[name] = t[-1] - t[-2]
        f_case, peq_case = self.load_cases(); lap("load_cases_s")
        self.assemble(); lap("assemble_s")
        f = self.combine(f_case, factors); peq = self.combine(peq_case, factors); lap("combine_s")
        first = self.plan is None
        self.make_plan()
        if first:
            lap("plan_s")
        self.factorize(); lap("factor_s")
        u = self.solve(f); lap("solve_s")
        # residual of the assembled K on the free DOFs (same check as the engine's direct path)
        ku = self.spmm(u)
        if self.has_springs:
            ku += self.kspring[:, None] * u
        free = self.fixed == 0
        num = np.linalg.norm((ku - f)[free], axis=0); den = np.linalg.norm(f[free], axis=0)
        self.relres = float(np.max(np.where(den > 0, num / np.where(den > 0, den, 1.0), 0.0)))
        lap("residual_s")
        rec = self.recover(u, f, peq); lap("recover_s")
        rec["U"] = u
        return rec
