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


class TreeParallelSolver(CholSolver):
    """The sparse direct solve of ONE model on two GPUs (strong scaling; SURVEY.md 8e only asks for sharding by
    combinations -- the reference describes a single shared factorisation, pydantic_models.py:1216-1219).

    The root front of the nested dissection has two children: rank r factorises the subtree under child r
    (fers_b200_chol_plan_partition drops the other fronts from this rank's level lists; every offset stays that of the whole
    plan), rank 1 sends the Schur complement of its top front to rank 0 over NVLink (NCCL send / recv; gloo in tests), rank 0
    extend-adds both and factorises the root.  The substitution mirrors it: forward on the two subtrees, rank 1's boundary
    right-hand side to rank 0, root forward + backward on rank 0, root solution to rank 1, backward on the subtrees, the two
    halves of the permuted solution exchanged -- both ranks end with the whole solution.  What crosses the link per
    factorisation: one [6 n2]^2 Schur complement (0.96 GB at configs[3]); per substitution: 2 x 5.6 MB + the two halves of
    the solution."""

    def __init__(self, plan: CholPlan, device, rank: int, world: int, group=None, max_nrhs: int = 64):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank, self.peer = int(rank), 1 - int(rank)
        fronts = plan.export()["fronts"].copy()          # the whole plan, before this rank's fronts are selected
        out = (C.c_longlong * 12)()
        _lib.check(plan.lib.fers_b200_chol_plan_partition(plan.handle, int(rank), int(world), out, 12), "chol_plan_partition")
        self.root, self.c0, self.c1, self.ls = (int(out[i]) for i in range(4))
        self.sub = ((int(out[4]), int(out[5])), (int(out[6]), int(out[7])))
        self.root_pos = (int(out[8]), int(out[8] + out[9]))
        self.fronts_kept, self.k_blocks_kept = int(out[10]), int(out[11])
        f1 = fronts[self.c1]
        m2 = 6 * int(f1["n2"])
        arena = plan.info.schur_doubles[0] if ((self.ls - 1) & 1) else 0
        self._schur = (arena + int(f1["schur_off"]), m2 * m2)                     # doubles into the work buffer
        self._rhs_c1 = (int(f1["rhs_off"]), 6 * int(f1["n1"] + f1["n2"]))         # rows of the front-local store W
        self._rhs_rows = int(plan.info.rhs_rows)
        super().__init__(plan, device, max_nrhs)
        self.wk = self.work[: (self.work_bytes // 8) * 8].view(self.torch.float64)
        self.link_bytes = {"factor": 8 * self._schur[1]}

    def _send(self, t):
        self.dist.send(t, self.peer, group=self.group)

    def _recv(self, t):
        self.dist.recv(t, self.peer, group=self.group)

    def factorize(self, d_vals, d_diag_slot, d_fixed, d_kspring=None, check: bool = True) -> int:
        ks = None if d_kspring is None else d_kspring.data_ptr()

        def levels(lo, hi):
            _lib.check(self.lib.fers_b200_chol_factor_levels(
                self.plan.handle, self.index.data_ptr(), d_vals.data_ptr(), d_diag_slot.data_ptr(), d_fixed.data_ptr(), ks,
                self.factor.data_ptr(), self.work.data_ptr(), self.work_bytes, self.info_dev.data_ptr(), lo, hi, self.stream),
                "chol_factor_levels")
        levels(0, self.ls)
        off, n = self._schur
        if self.rank == 1:
            self._send(self.wk[off: off + n])
        else:
            self._recv(self.wk[off: off + n])
            levels(self.ls, self.ls + 1)
        self.factored = True
        if not check:
            return 0
        info = self.info_dev.clone()
        self.dist.all_reduce(info, op=self.dist.ReduceOp.MAX, group=self.group)
        return int(info.item())

    def solve(self, d_b, d_x, d_fixed, accumulate: bool = False) -> None:
        nrhs = int(d_b.shape[1])

        def phase(bits, lo, hi):
            _lib.check(self.lib.fers_b200_chol_solve_phase(
                self.plan.handle, self.index.data_ptr(), self.factor.data_ptr(), d_fixed.data_ptr(), nrhs, d_b.data_ptr(),
                d_x.data_ptr(), 1 if accumulate else 0, self.work.data_ptr(), self.work_bytes, bits, lo, hi, self.stream),
                "chol_solve_phase")
        y0 = self._rhs_rows * nrhs                                         # the permuted solution Y follows W
        rows = lambda a, b: self.wk[y0 + 6 * a * nrhs: y0 + 6 * b * nrhs]
        w_c1 = self.wk[self._rhs_c1[0] * nrhs: (self._rhs_c1[0] + self._rhs_c1[1]) * nrhs]
        phase(1, 0, self.ls)
        if self.rank == 1:
            self._send(w_c1)
            self._recv(rows(*self.root_pos))
        else:
            self._recv(w_c1)
            phase(3, self.ls, self.ls + 1)
            self._send(rows(*self.root_pos))
        phase(2, 0, self.ls)
        mine, theirs = rows(*self.sub[self.rank]), rows(*self.sub[self.peer])
        ops = [self.dist.P2POp(self.dist.isend, mine, self.peer, group=self.group),
               self.dist.P2POp(self.dist.irecv, theirs, self.peer, group=self.group)]
        if self.rank == 1:
            ops.reverse()
        for r in self.dist.batch_isend_irecv(ops):
            r.wait()
        phase(4, 0, self.ls + 1)
        self.link_bytes["solve"] = 8 * (w_c1.numel() + rows(*self.root_pos).numel() + mine.numel() + theirs.numel())
