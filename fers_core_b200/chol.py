"""Host handle of the sparse multifrontal Cholesky of libfers_b200.so (include/fers_b200.h,
"sparse direct solve").  The reference engine's first-order path is "a batch first-order solve that
shares one factorization across many load cases" (fers_core/types/pydantic_models.py:1216-1219);
``CholPlan`` is the analysis phase of that factorisation (ordering + symbolic, done once per
sparsity pattern on the host), ``CholSolver`` owns the device buffers and runs the numeric
factorisation / substitution kernels.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib

FRONT_DTYPE = np.dtype([("n1", "<i4"), ("n2", "<i4"), ("nodes_off", "<i4"), ("child_begin", "<i4"),
                        ("child_end", "<i4"), ("parent", "<i4"), ("pad0", "<i4"), ("pad1", "<i4"),
                        ("panel_off", "<i8"), ("schur_off", "<i8"), ("dinv_off", "<i8"), ("rhs_off", "<i8")])
KSCATTER_DTYPE = np.dtype([("front", "<i4"), ("k", "<i4"), ("bi", "<i4"), ("bj", "<i4")])
CHOL_NB = 64
DEFAULT_LEAF_NODES = 48


@dataclass
class PlanInfo:
    n_fronts: int
    n_levels: int
    max_front_nodes: int
    factor_doubles: int
    schur_doubles: tuple
    rhs_rows: int
    index_bytes: int
    flops: float
    nnz_l: int
    n_kscatter: int
    root_nodes: int


def host_pattern(n_nodes: int, p1: np.ndarray, p2: np.ndarray, p_nodes: np.ndarray | None = None):
    """(rowptr, colidx) of the block-CSR node pattern on the host -- integer for integer what
    fers_b200_pattern_build leaves on the device (tests/test_gpu_parity.py checks it)."""
    lib = _lib.load()
    a = np.ascontiguousarray(p1, np.int32)
    b = np.ascontiguousarray(p2, np.int32)
    pn = np.zeros(0, np.int32) if p_nodes is None else np.ascontiguousarray(p_nodes, np.int32).reshape(-1)
    npl = len(pn) // 3
    cap = n_nodes + 2 * len(a) + 6 * npl
    rowptr = np.empty(n_nodes + 1, np.int32)
    colidx = np.empty(max(cap, 1), np.int32)
    nnzb = C.c_int()
    _lib.check(lib.fers_b200_host_pattern(n_nodes, len(a), a.ctypes.data if len(a) else None, b.ctypes.data if len(b) else None,
                                          npl, pn.ctypes.data if npl else None, rowptr.ctypes.data, colidx.ctypes.data, cap,
                                          C.byref(nnzb)), "host_pattern")
    return rowptr, colidx[: nnzb.value]


class CholPlan:
    """Nested-dissection ordering + symbolic factorisation of a block-CSR node pattern."""

    def __init__(self, rowptr: np.ndarray, colidx: np.ndarray, xyz: np.ndarray, leaf_nodes: int = DEFAULT_LEAF_NODES,
                 node_restrained: np.ndarray | None = None):
        self.lib = _lib.load()
        self.n_nodes = int(len(rowptr) - 1)
        rp = np.ascontiguousarray(rowptr, np.int32)
        ci = np.ascontiguousarray(colidx, np.int32)
        xy = np.ascontiguousarray(xyz, np.float64)
        h = C.c_void_p()
        mask = None
        if node_restrained is not None and np.any(node_restrained):
            mask = np.ascontiguousarray(node_restrained, np.uint8)
            assert mask.shape == (self.n_nodes,)
        _lib.check(self.lib.fers_b200_chol_plan_create(self.n_nodes, rp.ctypes.data, ci.ctypes.data, xy.ctypes.data,
                                                       int(leaf_nodes), None if mask is None else mask.ctypes.data,
                                                       C.byref(h)), "chol_plan_create")
        self.handle = h
        out = (C.c_longlong * 12)()
        _lib.check(self.lib.fers_b200_chol_plan_info(self.handle, out, 12), "chol_plan_info")
        self.info = PlanInfo(n_fronts=out[0], n_levels=out[1], max_front_nodes=out[2], factor_doubles=out[3],
                             schur_doubles=(out[4], out[5]), rhs_rows=out[6], index_bytes=out[7],
                             flops=float(out[8]) * 1e6, nnz_l=out[9], n_kscatter=out[10], root_nodes=out[11])

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.fers_b200_chol_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def perm(self) -> np.ndarray:
        """Elimination position -> node."""
        p = np.empty(self.n_nodes, np.int32)
        _lib.check(self.lib.fers_b200_chol_plan_perm(self.handle, p.ctypes.data), "chol_plan_perm")
        return p

    def levels(self) -> np.ndarray:
        """[n_levels, 4]: fronts, max own nodes, max front nodes, Mflop."""
        nl = self.info.n_levels
        out = (C.c_longlong * (4 * nl))()
        _lib.check(self.lib.fers_b200_chol_plan_levels(self.handle, out, nl), "chol_plan_levels")
        return np.array(out[:], np.int64).reshape(nl, 4)

    def export(self) -> dict:
        """Host copy of every index array the device kernels use (tests, diagnostics)."""
        buf = np.zeros(self.info.index_bytes, np.uint8)
        offs = (C.c_longlong * 8)()
        nl = self.info.n_levels
        lp = np.zeros(nl + 1, np.int32)
        kp = np.zeros(nl + 1, np.int64)
        _lib.check(self.lib.fers_b200_chol_plan_export(self.handle, buf.ctypes.data, offs, lp.ctypes.data, kp.ctypes.data),
                   "chol_plan_export")
        nf = self.info.n_fronts
        fronts = np.frombuffer(buf, FRONT_DTYPE, nf, offs[0])
        n_nodes_list = int((fronts["n1"] + fronts["n2"]).sum())
        n_child = int(offs[7])
        inv_off = np.frombuffer(buf, np.int64, n_child, offs[3])
        n_inv = 0
        for f in fronts:
            n_inv += int(f["n1"] + f["n2"]) * int(f["child_end"] - f["child_begin"])
        return dict(
            fronts=fronts,
            nodes=np.frombuffer(buf, np.int32, n_nodes_list, offs[1]),
            child=np.frombuffer(buf, np.int32, n_child, offs[2]),
            inv_off=inv_off,
            inv=np.frombuffer(buf, np.int32, n_inv, offs[4]),
            level_fronts=np.frombuffer(buf, np.int32, nf, offs[5]),
            kscatter=np.frombuffer(buf, KSCATTER_DTYPE, self.info.n_kscatter, offs[6]),
            level_ptr=lp, kscatter_ptr=kp, raw=buf,
        )


class CholSolver:
    """Device side of the sparse direct solve: owns index / factor / work buffers (torch tensors) and
    calls the numeric kernels.  One instance per (plan, device)."""

    def __init__(self, plan: CholPlan, device, max_nrhs: int = 64):
        import torch
        self.torch = torch
        self.plan = plan
        self.lib = plan.lib
        self.dev = device
        self.max_nrhs = max_nrhs
        self.index = torch.empty(plan.info.index_bytes, dtype=torch.uint8, device=device)
        self.factor_bytes = self.lib.fers_b200_chol_factor_bytes(plan.handle)
        self.work_bytes = self.lib.fers_b200_chol_work_bytes(plan.handle, max_nrhs)
        self.factor = torch.empty(self.factor_bytes, dtype=torch.uint8, device=device)
        self.work = torch.empty(self.work_bytes, dtype=torch.uint8, device=device)
        self.info_dev = torch.zeros(1, dtype=torch.int32, device=device)
        st = torch.cuda.current_stream(device).cuda_stream
        _lib.check(self.lib.fers_b200_chol_upload(plan.handle, self.index.data_ptr(), st), "chol_upload")
        self.factored = False

    @property
    def stream(self) -> int:
        return self.torch.cuda.current_stream(self.dev).cuda_stream

    def factorize(self, d_vals, d_diag_slot, d_fixed, d_kspring=None, check: bool = True) -> int:
        """Numeric factorisation; returns info (0 = positive definite)."""
        info = C.c_int(0)
        _lib.check(self.lib.fers_b200_chol_factor(
            self.plan.handle, self.index.data_ptr(), d_vals.data_ptr(), d_diag_slot.data_ptr(), d_fixed.data_ptr(),
            None if d_kspring is None else d_kspring.data_ptr(), self.factor.data_ptr(), self.work.data_ptr(),
            self.work_bytes, self.info_dev.data_ptr(), C.byref(info) if check else None, self.stream), "chol_factor")
        self.factored = True
        return info.value

    def solve(self, d_b, d_x, d_fixed, accumulate: bool = False) -> None:
        """x (+)= K^-1 b for [n_dof, nrhs] device tensors (nrhs even, <= max_nrhs)."""
        nrhs = int(d_b.shape[1])
        _lib.check(self.lib.fers_b200_chol_solve(
            self.plan.handle, self.index.data_ptr(), self.factor.data_ptr(), d_fixed.data_ptr(), nrhs, d_b.data_ptr(),
            d_x.data_ptr(), 1 if accumulate else 0, self.work.data_ptr(), self.work_bytes, self.stream), "chol_solve")

    def profile(self) -> dict:
        out = (C.c_double * 20)()
        _lib.check(self.lib.fers_b200_chol_profile(out, 20), "chol_profile")
        keys = ["init_ms", "scatter_ms", "potrf_ms", "trsm_ms", "update_ms", "syrk_ms", "factor_ms", "solve_ms",
                "factor_launches", "solve_launches", "syrk_sum_ms", "factor_sum_ms", "solve_sum_ms", "n_factor",
                "n_solve", "n_syrk_launches", "oz_ms", "oz_sum_ms", "n_oz_fronts", "oz_setting"]
        return dict(zip(keys, out[:]))
