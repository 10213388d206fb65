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


def tree_owner(k: int, d: int, j: int) -> int:
    """Rank that owns the upper front (depth d, index j) of a 2^k-rank tree: the leftmost rank below it."""
    return j << (k - d)


def tree_schedule(k: int, ls: int, top: list) -> dict:
    """The order of work and messages of the tree-parallel solve, the same on every rank (a rank skips the messages it is not
    part of; level calls on levels where it owns nothing are empty).  ``top[d][j]`` = front id of the upper front (d, j).
    Steps: ("levels" | "fwd" | "bwd", first level, one past last level); ("schur" | "w", front, depth, src, dst): the Schur
    complement / the front-local right-hand side of a right child to the owner of its parent; ("anc", [fronts], src, dst): the
    solution rows of an upper front and of its ancestors to the owner of its right child; ("gather",): every rank's own rows
    to everyone; ("unpermute",)."""
    factor, solve = [("levels", 0, ls - k + 1)], [("fwd", 0, ls - k + 1)]
    for d in range(k - 1, -1, -1):
        for j in range(1 << d):
            fid, src, dst = top[d + 1][2 * j + 1], tree_owner(k, d + 1, 2 * j + 1), tree_owner(k, d, j)
            factor.append(("schur", fid, d + 1, src, dst))
            solve.append(("w", fid, d + 1, src, dst))
        factor.append(("levels", ls - d, ls - d + 1))
        solve.append(("fwd", ls - d, ls - d + 1))
    for d in range(k):
        solve.append(("bwd", ls - d, ls - d + 1))
        for j in range(1 << d):
            chain = [top[dd][j >> (d - dd)] for dd in range(d, -1, -1)]
            solve.append(("anc", chain, tree_owner(k, d, j), tree_owner(k, d + 1, 2 * j + 1)))
    solve += [("bwd", 0, ls - k + 1), ("gather",), ("unpermute",)]
    return {"factor": factor, "solve": solve}


class TreeParallelSolver(CholSolver):
    """The sparse direct solve of ONE model on 2^k GPUs (strong scaling; SURVEY.md 8e only asks for sharding by
    combinations -- the reference describes a single shared factorisation, pydantic_models.py:1216-1219).

    The top k levels of the nested-dissection front tree are a binary tree.  Rank s factorises the subtree under its s-th
    depth-k front; an upper front (depth d, index j) belongs to the leftmost rank below it, j << (k - d)
    (fers_b200_chol_plan_partition drops everything else from this rank's level lists; every offset stays that of the whole
    plan).  Going up, the owner of a right child sends its Schur complement to the owner of the parent -- which holds the left
    child -- over NVLink (NCCL send / recv; gloo in tests); the parent is extend-added and factorised there.  The
    substitution mirrors it: forward on the subtrees, boundary right-hand sides up the tree, backward from the root with the
    solution of the ancestors handed down to the right children, and one all_gather of the rows each rank owns -- every rank
    ends with the whole solution.  Across the links per factorisation: one [6 n2]^2 Schur complement per upper front
    (0.96 GB for the root of configs[3])."""

    def __init__(self, plan: CholPlan, device, rank: int, world: int, group=None, max_nrhs: int = 64):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank, self.world = int(rank), int(world)
        fronts = plan.export()["fronts"].copy()          # the whole plan, before this rank's fronts are selected
        n_out = 4 + 4 * self.world
        out = (C.c_longlong * n_out)()
        _lib.check(plan.lib.fers_b200_chol_plan_partition(plan.handle, self.rank, self.world, out, n_out), "chol_plan_partition")
        k = self.k = int(out[0])
        self.ls = int(out[1])
        self.fronts_kept, self.k_blocks_kept = int(out[2]), int(out[3])
        n_top = (2 << k) - 1
        ids = [int(out[4 + t]) for t in range(n_top)]
        self.top = [[ids[(1 << d) - 1 + j] for j in range(1 << d)] for d in range(k + 1)]     # top[d][j] = front id
        lo = [int(out[4 + n_top + s]) for s in range(self.world)]
        self._f = fronts
        sd0 = int(plan.info.schur_doubles[0])
        self._rhs_rows = int(plan.info.rhs_rows)

        def own(fid):                                    # [first, one past last) elimination position of a front's own nodes
            return int(fronts[fid]["pad0"]), int(fronts[fid]["pad0"] + fronts[fid]["n1"])
        self._own = own
        # rows of the permuted solution each rank owns: its subtree and the upper fronts it factorises
        self.owned = []
        for r in range(self.world):
            rng = [(lo[r], own(self.top[k][r])[1])]
            for d in range(k):
                for j in range(1 << d):
                    if self.owner(d, j) == r:
                        rng.append(own(self.top[d][j]))
            self.owned.append(rng)
        self._sd0 = sd0
        super().__init__(plan, device, max_nrhs)
        self.wk = self.work[: (self.work_bytes // 8) * 8].view(self.torch.float64)
        self.link_bytes = {"factor": 0, "solve": 0}

    def owner(self, d: int, j: int) -> int:
        return tree_owner(self.k, d, j)

    def _schur(self, fid: int, depth: int):
        """The Schur complement of front `fid` (tree depth `depth`) inside the work buffer, as a flat tensor."""
        m2 = 6 * int(self._f[fid]["n2"])
        off = (self._sd0 if ((self.ls - depth) & 1) else 0) + int(self._f[fid]["schur_off"])
        return self.wk[off: off + m2 * m2]

    def factorize(self, d_vals, d_diag_slot, d_fixed, d_kspring=None, check: bool = True) -> int:
        ks = None if d_kspring is None else d_kspring.data_ptr()
        sent = 0
        for step in tree_schedule(self.k, self.ls, self.top)["factor"]:
            if step[0] == "levels":
                _lib.check(self.lib.fers_b200_chol_factor_levels(
                    self.plan.handle, self.index.data_ptr(), d_vals.data_ptr(), d_diag_slot.data_ptr(), d_fixed.data_ptr(), ks,
                    self.factor.data_ptr(), self.work.data_ptr(), self.work_bytes, self.info_dev.data_ptr(), step[1], step[2],
                    self.stream), "chol_factor_levels")
            else:
                _, fid, depth, src, dst = step
                if self.rank == src:
                    t = self._schur(fid, depth)
                    self.dist.send(t, dst, group=self.group)
                    sent += 8 * t.numel()
                elif self.rank == dst:
                    self.dist.recv(self._schur(fid, depth), src, group=self.group)
        self.link_bytes["factor"] = sent
        self.factored = True
        if not check:
            return 0
        info = self.info_dev.clone()
        self.dist.all_reduce(info, op=self.dist.ReduceOp.MAX, group=self.group)
        return int(info.item())

    def solve(self, d_b, d_x, d_fixed, accumulate: bool = False) -> None:
        torch = self.torch
        nrhs = int(d_b.shape[1])

        def phase(bits, lo, hi):
            _lib.check(self.lib.fers_b200_chol_solve_phase(
                self.plan.handle, self.index.data_ptr(), self.factor.data_ptr(), d_fixed.data_ptr(), nrhs, d_b.data_ptr(),
                d_x.data_ptr(), 1 if accumulate else 0, self.work.data_ptr(), self.work_bytes, bits, lo, hi, self.stream),
                "chol_solve_phase")
        y0 = self._rhs_rows * nrhs                                         # the permuted solution Y follows W

        def rows(a, b):
            return self.wk[y0 + 6 * a * nrhs: y0 + 6 * b * nrhs]

        def w_block(fid):                                                   # front-local right-hand side rows of a front
            f = self._f[fid]
            return self.wk[int(f["rhs_off"]) * nrhs: (int(f["rhs_off"]) + 6 * int(f["n1"] + f["n2"])) * nrhs]
        sent = 0
        for step in tree_schedule(self.k, self.ls, self.top)["solve"]:
            kind = step[0]
            if kind == "fwd":
                phase(1, step[1], step[2])
            elif kind == "bwd":
                phase(2, step[1], step[2])
            elif kind == "w":                                               # boundary right-hand sides up the tree
                _, fid, _depth, src, dst = step
                if self.rank == src:
                    self.dist.send(w_block(fid), dst, group=self.group)
                    sent += 8 * w_block(fid).numel()
                elif self.rank == dst:
                    self.dist.recv(w_block(fid), src, group=self.group)
            elif kind == "anc":                                             # the ancestors' solution down the tree
                _, chain_fronts, src, dst = step
                if self.rank not in (src, dst):
                    continue
                chain = [self._own(fid) for fid in chain_fronts]
                if self.rank == src:
                    buf = torch.cat([rows(a, b) for a, b in chain])
                    self.dist.send(buf, dst, group=self.group)
                    sent += 8 * buf.numel()
                else:
                    buf = torch.empty(sum(6 * (b - a) * nrhs for a, b in chain), dtype=torch.float64, device=self.wk.device)
                    self.dist.recv(buf, src, group=self.group)
                    o = 0
                    for a, b in chain:
                        n = 6 * (b - a) * nrhs
                        rows(a, b).copy_(buf[o: o + n]); o += n
            elif kind == "gather":                                          # one all_gather of equally padded blocks
                width = max(sum(6 * (b - a) for a, b in rng) for rng in self.owned) * nrhs
                mine = torch.zeros(width, dtype=torch.float64, device=self.wk.device)
                o = 0
                for a, b in self.owned[self.rank]:
                    n = 6 * (b - a) * nrhs
                    mine[o: o + n] = rows(a, b); o += n
                allb = torch.empty(self.world * width, dtype=torch.float64, device=self.wk.device)
                self.dist.all_gather_into_tensor(allb, mine, group=self.group)
                for r in range(self.world):
                    if r == self.rank:
                        continue
                    o = r * width
                    for a, b in self.owned[r]:
                        n = 6 * (b - a) * nrhs
                        rows(a, b).copy_(allb[o: o + n]); o += n
                sent += 8 * (self.world - 1) * width
            else:
                phase(4, 0, self.ls + 1)
        self.link_bytes["solve"] = sent
