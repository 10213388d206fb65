"""CUDA engine: drives libfers_b200.so through its C ABI for one analysis.

Host-side mirror of the reference boundary ``fers_calculations.calculate_from_json`` /
``calculate_from_file`` (fers_core/fers/fers.py:300, :260).  PyTorch is used only for device
memory, streams and host<->device copies; every arithmetic step of the path is a kernel of
``libfers_b200.so`` (include/fers_b200.h).  There is no CPU fallback: without a CUDA device or
without the built library every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import json
import threading
import time
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, cable, chol, unity, wire
from .hostlogic import (
    INACTIVE_SCALE, effective_bc, imperfect_coordinates, member_frames_host, mpc_transform, pdelta_suppressed_axes, rhs_plan,
    rigid_property_row,
    unilateral_dofs, unilateral_update, curve_stiffness, curves_converged, hinge_curve_stiffness, rel_with_hinge_curves,
    hinge_capacity_notes,
)
from .results import ResultArrays, SECTION_FRACS, lazy_bundle, results_to_json, results_to_wire

DENSE_MAX_DOF = 8192          # models up to this many DOFs use the dense Cholesky path
SOLVERS = ("auto", "dense", "cholesky", "pcg")   # auto: dense up to DENSE_MAX_DOF, sparse Cholesky above
CHOL_REFINE_STEPS = 2         # at most this many iterative-refinement passes after the direct substitution
CHOL_REFINE_TOL = 1e-12       # refine while the relative residual max_c ||f_c - K u_c|| / ||f_c|| exceeds this
STAGE_MAX_BYTES = 256 << 20   # result fields up to this size travel through pinned staging buffers
CHOL_RESIDUAL_LIMIT = 1e-8    # relative residual above which the direct solve is reported as failed
PCG_TOL = 1e-12               # relative residual ||r|| / ||b|| per right-hand side
PCG_MAX_ITER = 200000
PCG_ALLOW_UNCONVERGED = False  # profiling captures only (bench.py --profile-iters)
NR_CHOICES = (8, 16, 32, 64)
ASSEMBLY_GATHER = 4           # assembly_variant: 4 = gather kernel (default); 0-3 = atomic scatter-add variants


CABLE_TOL = 1.0e-9          # relative change of the displacements between two Newton steps of the cable state
CABLE_MAX_ITER = 60


class EngineError(RuntimeError):
    pass


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise EngineError("fers_core_b200 needs a CUDA device (sm_100a); none is visible and there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise EngineError(f"fers_core_b200 runs on CUDA devices only, got {dev}")
    if dev.index is None:            # "cuda" without an index: the library must be set to the device torch is on
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _pad_nr(n: int) -> int:
    for c in NR_CHOICES:
        if n <= c:
            return c
    return NR_CHOICES[-1]


@dataclass
class SolveStats:
    method: str = "dense"
    iterations: int = 0
    relres: float = 0.0
    n_free: int = 0
    n_dead: int = 0
    nnzb: int = 0


class Engine:
    """One model resident on one GPU."""

    def __init__(self, model: wire.ModelArrays, device=None, assembly_variant: int = ASSEMBLY_GATHER, solver: str = "auto",
                 plan_ahead: bool = False, tree_parallel=None):
        if solver not in SOLVERS:
            raise EngineError(f"solver must be one of {SOLVERS}, got {solver!r}")
        self.solver = solver
        # (rank, world[, group]) of a torch.distributed group of two GPUs that hold the SAME model: the sparse factorisation
        # is split by the two subtrees under the root front (chol.TreeParallelSolver); every other stage runs on both
        self.tree_parallel = tree_parallel
        self._chol = None
        self._k_version = 0
        self._chol_version = -1
        self.model = model
        self.dev = _require_cuda(device)
        self.lib = _lib.load()
        torch.cuda.set_device(self.dev)
        _lib.check(self.lib.fers_b200_set_device(self.dev.index), "set_device")
        sm, maj, mnr, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        _lib.check(self.lib.fers_b200_device_info(self.dev.index, C.byref(sm), C.byref(maj), C.byref(mnr),
                                                  C.byref(mem)), "device_info")
        self.assembly_variant = assembly_variant
        self.stats = SolveStats()
        self.timings: dict = {}
        self.h2d_bytes = 0
        # nodes with all six DOFs restrained by PERMANENT supports: left out of the sparse factorisation
        self.h_node_restrained = (model.bc_code.reshape(-1, 6) == wire.BC_FIXED).all(axis=1)
        if self.h_node_restrained.all():
            self.h_node_restrained = None
        self._plan_thread = None
        self._plan_result = None
        if plan_ahead and self._wants_sparse_cholesky():
            # optional: the ordering / symbolic phase of the sparse direct solve (host, ~18 ms at configs[3]) starts now,
            # on a host thread, from a host-built copy of the pattern, and overlaps the uploads and the device set-up.
            # Off by default: the plan already uses every host core, so on the 16-core bench host the overlapped plan
            # took 29 ms instead of 17 and the set-up 12-19 ms instead of 5 -- no net gain (tools/e2e_probe.py).
            self._plan_thread = threading.Thread(target=self._plan_worker, daemon=True)
            self._plan_thread.start()
        self._upload_static()
        self._upload_load_cases()
        self._build_pattern()

    # ---- factorisation plan on a host thread ---------------------------------------------------
    def _wants_sparse_cholesky(self) -> bool:
        m = self.model
        if self.solver == "cholesky":
            return True
        return self.solver == "auto" and m.n_dof > DENSE_MAX_DOF

    def _pattern_pairs(self):
        """Node pairs of the pattern: members, then the five extra pairs of every warping member (see _build_pattern)."""
        m = self.model
        p1, p2 = m.m_n1, m.m_n2
        if len(m.w_member):
            s_, e_ = m.m_n1[m.w_member], m.m_n2[m.w_member]
            a1, a2 = m.w_aux[:, 0], m.w_aux[:, 1]
            p1 = np.concatenate([p1, np.stack([s_, s_, e_, e_, a1], axis=1).reshape(-1)])
            p2 = np.concatenate([p2, np.stack([a1, a2, a1, a2, a2], axis=1).reshape(-1)])
        return p1, p2

    def _plan_worker(self) -> None:
        try:
            m = self.model
            t0 = time.perf_counter()
            p1, p2 = self._pattern_pairs()
            rowptr, colidx = chol.host_pattern(m.n_nodes, p1, p2, m.p_nodes if m.n_plates else None)
            plan = chol.CholPlan(rowptr, colidx, m.xyz, node_restrained=self.h_node_restrained)
            self._plan_result = (plan, len(colidx), (time.perf_counter() - t0) * 1e3)
        except Exception as e:                      # surfaced by _solve_cholesky, which falls back to the inline plan
            self._plan_result = e

    # ---- helpers ---------------------------------------------------------------------------
    @property
    def d_ks(self):
        """Support-spring diagonal for the solve kernels, or None when the model has no springs
        (saves one dependent load per DOF row in the SpMM epilogue)."""
        return self.d_kspring if self.has_springs else None

    @property
    def stream(self) -> int:
        # read before every launch: also makes this engine's GPU the current one (the library launches on the current
        # device; another Engine of the same process may have switched it)
        if torch.cuda.current_device() != self.dev.index:
            torch.cuda.set_device(self.dev)
        return torch.cuda.current_stream(self.dev).cuda_stream

    def _t(self, a: np.ndarray, dtype) -> torch.Tensor:
        """Host array -> device tensor (through pinned memory for large arrays)."""
        a = np.ascontiguousarray(a, dtype=dtype)
        t = torch.from_numpy(a)
        self.h2d_bytes += a.nbytes
        if a.nbytes >= (1 << 20) and not t.is_pinned():
            t = t.pin_memory()
        return t.to(self.dev, non_blocking=True)

    def _empty(self, shape, dtype=torch.float64) -> torch.Tensor:
        return torch.empty(shape, dtype=dtype, device=self.dev)

    def _zeros(self, shape, dtype=torch.float64) -> torch.Tensor:
        return torch.zeros(shape, dtype=dtype, device=self.dev)

    # ---- set-up ----------------------------------------------------------------------------
    def _upload_static(self) -> None:
        m = self.model
        rigid = rigid_property_row(m.prop)
        table = np.vstack([m.prop, rigid[None, :]]) if m.prop.size else rigid[None, :]
        idx = np.where(m.m_prop >= 0, m.m_prop, table.shape[0] - 1)
        idx = np.where(m.m_type == wire.MT_RIGID, table.shape[0] - 1, idx).astype(np.int32)
        self.n_prop = int(table.shape[0])
        self.d_prop = self._t(table, np.float64)
        self.d_m_prop = self._t(idx, np.int32)
        self.d_n1 = self._t(m.m_n1, np.int32)
        self.d_n2 = self._t(m.m_n2, np.int32)
        self.d_type = self._t(m.m_type, np.int32)
        self.d_roll = self._t(m.m_roll, np.float64)
        self.has_rel = bool(np.any(m.m_rel >= 0))
        self.d_rel_idx = self._t(m.m_rel, np.int32) if self.has_rel else None
        if self.has_rel and (m.rel.ndim != 2 or m.rel.shape[1] != wire.REL_STRIDE):
            raise EngineError(f"ModelArrays.rel must be [Nr,{wire.REL_STRIDE}] (12 release springs, y_s, z_s, end offsets, E*I_yz), got {m.rel.shape}")
        self.d_rel = self._t(m.rel, np.float64) if self.has_rel else None
        # unilateral supports (PositiveOnly / NegativeOnly) start closed; engine._solve_nonlinear iterates their active set
        self.n_unilateral = int(len(unilateral_dofs(m)[0]))
        self.n_curves = len(m.support_curves)
        self.n_hinge_curves = len(m.hinge_curves)
        code, kk = effective_bc(m, np.ones(self.n_unilateral, bool))
        code = code.astype(np.uint8)
        self.has_springs = bool((code == wire.BC_SPRING).any()) or bool((unilateral_dofs(m)[2] > 0).any())
        self.d_bc_code = self._t(code.reshape(-1), np.uint8)
        self.d_bc_k = self._t(kk.reshape(-1), np.float64)
        self.d_supported = self._t((m.node_support_id >= 0).astype(np.uint8), np.uint8)
        # reactions are exactly zero off the supported nodes (k_reaction_finish): only their rows travel to the host
        self.sup_nodes = np.nonzero(m.node_support_id >= 0)[0]
        self.d_sup_nodes = torch.from_numpy(self.sup_nodes.astype(np.int64)).to(self.dev)
        self.d_p_nodes = self._t(m.p_nodes, np.int32) if m.n_plates else None
        self.d_p_data = self._t(m.p_data, np.float64) if m.n_plates else None
        self.d_p_flags = self._t(m.p_flags, np.int32) if m.n_plates else None
        self.d_xyz0 = self._t(m.xyz, np.float64)
        # catenary cables (cable.py): member rows; their tangents / right-hand-side corrections arrive per Newton step
        if int((m.m_type == wire.MT_CABLE).sum()) not in (0, len(m.c_member)) and not len(m.c_member):
            raise EngineError("ModelArrays: members of type Cable need the cable tables c_member / c_l0 / c_ea / c_w (wire.parse_document "
                              "fills them)")
        self.n_cables = int(len(m.c_member))
        self.d_c_member = self._t(m.c_member, np.int32) if self.n_cables else None

    def _upload_load_cases(self) -> None:
        """Load records of every case, resident on the device (inputs of fers_b200_load_case)."""
        self.d_lc = []
        for lc in self.model.load_cases:
            self.d_lc.append({k: (self._t(getattr(lc, k), dt) if len(getattr(lc, k)) else None)
                              for k, dt in (("nod_node", np.int32), ("nod_vec", np.float64), ("dl_member", np.int32),
                                            ("dl_data", np.float64), ("pl_member", np.int32), ("pl_data", np.float64),
                                            ("pp_plate", np.int32), ("pp_vec", np.float64))})

    def _build_pattern(self) -> None:
        m, lib = self.model, self.lib
        nn, nm, npl = m.n_nodes, m.n_members, m.n_plates
        # warping members couple the member ends with the auxiliary warping nodes: five extra node pairs each,
        # appended after the members for the pattern build only (their m_slot rows are the kernel's slot table)
        nw = self.n_warp = int(len(m.w_member))
        d_pn1, d_pn2, n_pairs = self.d_n1, self.d_n2, nm
        if nw:
            s_, e_ = m.m_n1[m.w_member], m.m_n2[m.w_member]
            a1, a2 = m.w_aux[:, 0], m.w_aux[:, 1]
            x1 = np.stack([s_, s_, e_, e_, a1], axis=1).reshape(-1)
            x2 = np.stack([a1, a2, a1, a2, a2], axis=1).reshape(-1)
            d_pn1 = self._t(np.concatenate([m.m_n1, x1]), np.int32)
            d_pn2 = self._t(np.concatenate([m.m_n2, x2]), np.int32)
            n_pairs = nm + 5 * nw
            self.d_w_member = self._t(m.w_member, np.int32)
            self.d_w_aux = self._t(m.w_aux, np.int32)
        cap = nn + 2 * n_pairs + 6 * npl
        self.d_rowptr = self._empty(nn + 1, torch.int32)
        self.d_colidx = self._empty(cap, torch.int32)
        self.d_m_slot = self._empty(max(n_pairs, 1) * 4, torch.int32)
        self.d_w_slot = self.d_m_slot[4 * nm:] if nw else None
        self.d_p_slot = self._empty(max(npl, 1) * 9, torch.int32)
        self.d_diag_slot = self._empty(nn, torch.int32)
        ws_bytes = lib.fers_b200_pattern_workspace(nn, n_pairs, npl)
        ws = self._empty(ws_bytes, torch.uint8)
        nnzb = C.c_int()
        _lib.check(lib.fers_b200_pattern_build(
            nn, n_pairs, _lib.ptr(d_pn1), _lib.ptr(d_pn2), npl, _lib.ptr(self.d_p_nodes),
            _lib.ptr(self.d_rowptr), _lib.ptr(self.d_colidx), cap, _lib.ptr(self.d_m_slot), _lib.ptr(self.d_p_slot),
            _lib.ptr(self.d_diag_slot), C.byref(nnzb), _lib.ptr(ws), ws_bytes, self.stream), "pattern_build")
        self.nnzb = nnzb.value
        self.stats.nnzb = self.nnzb
        self.d_colidx = self.d_colidx[: self.nnzb]
        self.d_vals = self._empty(self.nnzb * 36)
        # gather map of the default assembly kernel (half-members grouped by the block row they write)
        self.d_hm_ptr = self._empty(nn + 1, torch.int32)
        self.d_hm_list = self._empty(max(2 * nm, 1), torch.int32)
        gws_bytes = lib.fers_b200_gather_map_workspace(nn, nm)
        gws = self._empty(gws_bytes, torch.uint8)
        n_dup = C.c_int()
        _lib.check(lib.fers_b200_gather_map_build(
            nn, nm, self.nnzb, _lib.ptr(self.d_n1), _lib.ptr(self.d_n2), _lib.ptr(self.d_m_slot), _lib.ptr(self.d_hm_ptr),
            _lib.ptr(self.d_hm_list), C.byref(n_dup), _lib.ptr(gws), gws_bytes, self.stream), "gather_map_build")
        self.n_dup_half_members = n_dup.value
        # optional scratch of the gather kernel (frames + coefficients formed once per member): measured slower
        # than forming them per half-member at configs[3] (132 us against 106 us: the kernel is latency-bound, not
        # instruction-bound), so it is off by default
        self.d_member_cache = None
        nd = m.n_dof
        self.d_fixed = self._empty(nd, torch.uint8)
        self.d_free_index = self._empty(nd, torch.int32)
        self.d_kspring = self._empty(nd)
        self._bc_ws_bytes = lib.fers_b200_bc_workspace(nd)
        self._bc_ws = self._empty(self._bc_ws_bytes, torch.uint8)

    # ---- stages ----------------------------------------------------------------------------
    def set_unilateral(self, active: np.ndarray, curve_k: np.ndarray | None = None) -> None:
        """Support codes for the given active set of the unilateral supports and the current stiffness of the
        load-dependent springs (picked up by the next assemble())."""
        code, kk = effective_bc(self.model, active, curve_k)
        self.d_bc_code = self._t(code.astype(np.uint8).reshape(-1), np.uint8)
        self.d_bc_k = self._t(kk.reshape(-1), np.float64)

    def set_rel(self, rel: np.ndarray) -> None:
        """Member modifier records for the current stiffness of the load-dependent hinge springs (picked up by the
        next load_cases() / assemble() / recover())."""
        if rel.shape != self.model.rel.shape:
            raise EngineError(f"set_rel: expected {self.model.rel.shape}, got {rel.shape}")
        self.d_rel.copy_(torch.from_numpy(np.ascontiguousarray(rel, np.float64)))

    def assemble(self, d_xyz, d_scale=None, d_axial=None, kg_mode=0, kg_suppress=0, count: bool = True, cable_g=None) -> None:
        """Element stiffness + assembly + boundary conditions for one geometry / state.  ``cable_g`` [Ncable,3,3]: tangents of
        the catenary cables at the current state (their member rows carry scale 0 in ``d_scale``)."""
        m, lib, st = self.model, self.lib, self.stream
        self._k_version += 1
        n_rel = 0 if self.d_rel is None else self.d_rel.shape[0]
        if self.assembly_variant == ASSEMBLY_GATHER:
            # gather kernel: every block a member touches is written exactly once (no atomics, no zero pass);
            # blocks coupled only by plates are zeroed first
            if m.n_plates or self.n_warp:
                _lib.check(lib.fers_b200_zero(_lib.ptr(self.d_vals), self.d_vals.numel() * 8, st), "zero")
            _lib.check(lib.fers_b200_assemble_beams_gather(
                m.n_nodes, m.n_members, _lib.ptr(d_xyz), _lib.ptr(self.d_n1), _lib.ptr(self.d_n2), _lib.ptr(self.d_m_prop),
                _lib.ptr(self.d_type), _lib.ptr(self.d_roll), _lib.ptr(d_scale), _lib.ptr(self.d_rel_idx),
                _lib.ptr(self.d_prop), self.n_prop, _lib.ptr(self.d_rel), n_rel, _lib.ptr(d_axial), kg_mode, kg_suppress,
                _lib.ptr(self.d_m_slot), _lib.ptr(self.d_hm_ptr), _lib.ptr(self.d_hm_list), self.n_dup_half_members,
                _lib.ptr(self.d_diag_slot), _lib.ptr(self.d_member_cache), _lib.ptr(self.d_vals), st),
                "assemble_beams_gather")
        else:
            _lib.check(lib.fers_b200_zero(_lib.ptr(self.d_vals), self.d_vals.numel() * 8, st), "zero")
            _lib.check(lib.fers_b200_assemble_beams(
                m.n_members, _lib.ptr(d_xyz), _lib.ptr(self.d_n1), _lib.ptr(self.d_n2), _lib.ptr(self.d_m_prop),
                _lib.ptr(self.d_type), _lib.ptr(self.d_roll), _lib.ptr(d_scale), _lib.ptr(self.d_rel_idx),
                _lib.ptr(self.d_prop), self.n_prop, _lib.ptr(self.d_rel), n_rel,
                _lib.ptr(d_axial), kg_mode, kg_suppress, _lib.ptr(self.d_m_slot), _lib.ptr(self.d_vals),
                self.assembly_variant, st), "assemble_beams")
        if self.n_warp:
            _lib.check(lib.fers_b200_assemble_warping(
                self.n_warp, _lib.ptr(self.d_w_member), _lib.ptr(d_xyz), _lib.ptr(self.d_n1), _lib.ptr(self.d_n2),
                _lib.ptr(self.d_m_prop), _lib.ptr(d_scale), _lib.ptr(self.d_prop), _lib.ptr(self.d_w_slot),
                _lib.ptr(self.d_vals), st), "assemble_warping")
        if m.n_plates:
            _lib.check(lib.fers_b200_assemble_plates(
                m.n_plates, _lib.ptr(d_xyz), _lib.ptr(self.d_p_nodes), _lib.ptr(self.d_p_data), _lib.ptr(self.d_p_flags),
                _lib.ptr(self.d_p_slot), _lib.ptr(self.d_vals), st), "assemble_plates")
        if cable_g is not None and self.n_cables:
            _lib.check(lib.fers_b200_assemble_cables(self.n_cables, _lib.ptr(self.d_c_member), _lib.ptr(cable_g),
                                                     _lib.ptr(self.d_m_slot), _lib.ptr(self.d_vals), st), "assemble_cables")
        # count=False (re-assemblies inside an iteration): the free / dead DOF counts are not read back, no host sync
        n_free, n_dead = C.c_int(), C.c_int()
        _lib.check(lib.fers_b200_bc_build(
            m.n_nodes, _lib.ptr(self.d_bc_code), _lib.ptr(self.d_bc_k), _lib.ptr(self.d_vals), _lib.ptr(self.d_diag_slot),
            _lib.ptr(self.d_fixed), _lib.ptr(self.d_free_index), _lib.ptr(self.d_kspring), C.byref(n_free) if count else None,
            C.byref(n_dead) if count else None, _lib.ptr(self._bc_ws), self._bc_ws_bytes, st), "bc_build")
        if count:
            self.stats.n_free, self.stats.n_dead = n_free.value, n_dead.value

    def load_cases(self, d_xyz, d_scale=None):
        """Per load case: nodal load vector [Nlc,Ndof] and local equivalent loads [Nlc,Nm*12]."""
        m, lib, st = self.model, self.lib, self.stream
        nlc = len(m.load_cases)
        f = self._zeros((max(nlc, 1), m.n_dof))
        peq = self._zeros((max(nlc, 1), max(m.n_members, 1) * 12))
        for li, lc in enumerate(m.load_cases):
            t = self.d_lc[li]
            _lib.check(lib.fers_b200_load_case(
                m.n_nodes, m.n_members, _lib.ptr(d_xyz), _lib.ptr(self.d_n1), _lib.ptr(self.d_n2), _lib.ptr(self.d_m_prop),
                _lib.ptr(self.d_type), _lib.ptr(self.d_roll), _lib.ptr(d_scale), _lib.ptr(self.d_rel_idx),
                _lib.ptr(self.d_prop), _lib.ptr(self.d_rel),
                len(lc.nod_node), _lib.ptr(t["nod_node"]), _lib.ptr(t["nod_vec"]),
                len(lc.dl_member), _lib.ptr(t["dl_member"]), _lib.ptr(t["dl_data"]),
                len(lc.pl_member), _lib.ptr(t["pl_member"]), _lib.ptr(t["pl_data"]),
                len(lc.pp_plate), _lib.ptr(t["pp_plate"]), _lib.ptr(t["pp_vec"]), _lib.ptr(self.d_p_nodes),
                _lib.ptr(f[li]), _lib.ptr(peq[li]), st), "load_case")
        return f, peq

    def combine(self, per_case: torch.Tensor, factors: np.ndarray, nr: int) -> torch.Tensor:
        """[Nlc,rows] -> [rows,nr] with factors [ncols<=nr, Nlc] (zero-padded columns)."""
        nlc = per_case.shape[0]
        fac = np.zeros((nr, nlc))
        fac[: factors.shape[0]] = factors
        d_fac = self._t(fac, np.float64)
        out = self._empty((per_case.shape[1], nr))
        _lib.check(self.lib.fers_b200_combine(per_case.shape[1], nlc, nr, _lib.ptr(per_case), _lib.ptr(d_fac),
                                              _lib.ptr(out), self.stream), "combine")
        return out

    def solve(self, d_f: torch.Tensor, x0: torch.Tensor | None = None) -> torch.Tensor:
        """K_ff u = f for all columns of d_f [Ndof,nr]; returns u [Ndof,nr]."""
        m, lib, st = self.model, self.lib, self.stream
        nd, nr = d_f.shape
        method = self.solver
        if method == "auto":
            method = "dense" if nd <= DENSE_MAX_DOF else "cholesky"
        if method == "cholesky" and nr >= 2 and nr % 2 == 0:
            return self._solve_cholesky(d_f)
        if method == "dense" or (method == "cholesky" and nd <= DENSE_MAX_DOF):
            a = self._empty((nd, nd))
            _lib.check(lib.fers_b200_dense_expand(m.n_nodes, _lib.ptr(self.d_rowptr), _lib.ptr(self.d_colidx),
                                                  _lib.ptr(self.d_vals), _lib.ptr(self.d_fixed), _lib.ptr(self.d_ks),
                                                  _lib.ptr(a), st), "dense_expand")
            info_dev = self._zeros(1, torch.int32)
            info = C.c_int()
            _lib.check(lib.fers_b200_dense_cholesky(nd, 1, _lib.ptr(a), _lib.ptr(info_dev), C.byref(info), st),
                       "dense_cholesky")
            if info.value != 0:
                raise EngineError(f"stiffness matrix is not positive definite (leading minor {info.value}): "
                                  "the structure is a mechanism or insufficiently supported")
            u = d_f * (1 - self.d_fixed.to(torch.float64))[:, None]     # zero load rows of constrained DOFs
            u = u.contiguous()
            _lib.check(lib.fers_b200_dense_solve(nd, nr, 1, _lib.ptr(a), _lib.ptr(u), st), "dense_solve")
            self.stats.method = "dense"
            return u
        if nr not in NR_CHOICES:
            raise EngineError(f"PCG path needs nrhs in {NR_CHOICES}, got {nr}")
        minv = self._empty(m.n_nodes * 36)
        _lib.check(lib.fers_b200_block_jacobi(m.n_nodes, _lib.ptr(self.d_vals), _lib.ptr(self.d_diag_slot),
                                              _lib.ptr(self.d_fixed), _lib.ptr(self.d_ks), _lib.ptr(minv), st),
                   "block_jacobi")
        ws_bytes = lib.fers_b200_pcg_workspace(m.n_nodes, nr)
        ws = self._empty(ws_bytes, torch.uint8)
        u = x0 if x0 is not None else self._empty((nd, nr))
        iters = C.c_int()
        relres = (C.c_double * nr)()
        _lib.check(lib.fers_b200_pcg_solve(
            m.n_nodes, nr, _lib.ptr(self.d_rowptr), _lib.ptr(self.d_colidx), _lib.ptr(self.d_vals), _lib.ptr(self.d_fixed),
            _lib.ptr(self.d_ks), _lib.ptr(minv), _lib.ptr(d_f), _lib.ptr(u), 1 if x0 is not None else 0,
            PCG_TOL, PCG_MAX_ITER, 50, C.byref(iters), relres, _lib.ptr(ws), ws_bytes, st), "pcg_solve")
        self.stats.method = "pcg"
        self.stats.iterations = iters.value
        self.stats.relres = max(relres)
        if not (self.stats.relres <= PCG_TOL * 10) and not PCG_ALLOW_UNCONVERGED:
            raise EngineError(f"PCG did not converge: relative residual {self.stats.relres:.3e} after "
                              f"{iters.value} iterations")
        return u

    def _solve_cholesky(self, d_f: torch.Tensor) -> torch.Tensor:
        """Sparse multifrontal Cholesky (chol_plan.cu / chol_numeric.cu) + iterative refinement with
        the block-CSR SpMM.  The plan (ordering + symbolic) is built once per engine; the numeric
        factorisation is redone only when the matrix was reassembled."""
        m, lib, st = self.model, self.lib, self.stream
        nd, nr = d_f.shape
        if self._chol is None:
            t0 = time.perf_counter()
            plan = None
            if self._plan_thread is not None:
                self._plan_thread.join()
                self._plan_thread = None
                res, self._plan_result = self._plan_result, None
                if isinstance(res, tuple) and res[1] == self.nnzb:       # same pattern size as the device's
                    plan = res[0]
                    self.timings["chol_plan_thread_ms"] = res[2]
            if plan is None:
                rowptr = self.d_rowptr.cpu().numpy()
                colidx = self.d_colidx.cpu().numpy()
                plan = chol.CholPlan(rowptr, colidx, m.xyz, node_restrained=self.h_node_restrained)
            if self.tree_parallel is not None:
                self._chol = chol.TreeParallelSolver(plan, self.dev, *self.tree_parallel, max_nrhs=64)
            else:
                self._chol = chol.CholSolver(plan, self.dev, max_nrhs=64)
            self.timings["chol_plan_ms"] = (time.perf_counter() - t0) * 1e3          # what the pass waited for
        sol = self._chol
        if self._chol_version != self._k_version:
            info = sol.factorize(self.d_vals, self.d_diag_slot, self.d_fixed, self.d_ks)
            if info != 0:
                raise EngineError(f"stiffness matrix is not positive definite (pivot {info} in elimination order): "
                                  "the structure is a mechanism, insufficiently supported or buckled")
            self._chol_version = self._k_version
        u = self._empty((nd, nr))
        sol.solve(d_f, u, self.d_fixed)
        r = self._empty((nd, nr))
        ku = self._empty((nd, nr))
        norms = self._empty((2, nr))
        relres = float("nan")
        steps = 0
        while True:
            # true residual of the assembled K (block-CSR SpMM), all columns at once
            _lib.check(lib.fers_b200_spmm(m.n_nodes, nr, _lib.ptr(self.d_rowptr), _lib.ptr(self.d_colidx),
                                          _lib.ptr(self.d_vals), _lib.ptr(self.d_fixed), _lib.ptr(self.d_ks), _lib.ptr(u),
                                          _lib.ptr(ku), st), "spmm")
            _lib.check(lib.fers_b200_residual(nd, nr, _lib.ptr(d_f), _lib.ptr(ku), _lib.ptr(self.d_fixed), _lib.ptr(r),
                                              _lib.ptr(norms), st), "residual")
            nh = norms.cpu().numpy()
            with np.errstate(divide="ignore", invalid="ignore"):
                rel = np.sqrt(np.where(nh[1] > 0, nh[0] / np.where(nh[1] > 0, nh[1], 1.0), 0.0))
            relres = float(rel.max())
            # iterative refinement only while it is needed (same target as the PCG path) and still pays
            if relres <= CHOL_REFINE_TOL or steps >= CHOL_REFINE_STEPS or not np.isfinite(relres):
                break
            sol.solve(r, u, self.d_fixed, accumulate=True)
            steps += 1
        self.stats.method = "cholesky"
        self.stats.iterations = steps
        self.stats.relres = relres
        if not (relres <= CHOL_RESIDUAL_LIMIT):
            raise EngineError(f"sparse Cholesky solve failed: relative residual {relres:.3e} after refinement")
        return u

    def recover(self, d_xyz, d_u, d_f, d_peq, factors_pad, span, d_scale=None, d_axial_rhs=None, kg_mode=0,
                kg_suppress=0, detail=True, link_forces=None, cables=None) -> dict:
        """All recovery kernels for displacement columns d_u [Ndof,nr]."""
        m, lib, st = self.model, self.lib, self.stream
        nr = d_u.shape[1]
        nm = m.n_members
        out = {}
        r = self._empty((m.n_dof, nr))
        _lib.check(lib.fers_b200_reactions(m.n_nodes, nr, _lib.ptr(self.d_rowptr), _lib.ptr(self.d_colidx),
                                           _lib.ptr(self.d_vals), _lib.ptr(self.d_supported), _lib.ptr(d_u),
                                           _lib.ptr(d_f), _lib.ptr(r), st), "reactions")
        out["R"] = r
        out["sup_nodes"] = (self.sup_nodes, self.d_sup_nodes)
        if nm:
            fl = self._empty((nm, 12, nr)); fg = self._empty((nm, 12, nr)); ul = self._empty((nm, 12, nr))
            _lib.check(lib.fers_b200_recover_members(
                nm, nr, _lib.ptr(d_xyz), _lib.ptr(self.d_n1), _lib.ptr(self.d_n2), _lib.ptr(self.d_m_prop),
                _lib.ptr(self.d_type), _lib.ptr(self.d_roll), _lib.ptr(d_scale), _lib.ptr(self.d_rel_idx),
                _lib.ptr(self.d_prop), _lib.ptr(self.d_rel), _lib.ptr(d_axial_rhs), kg_mode, kg_suppress,
                _lib.ptr(d_u), _lib.ptr(d_peq), _lib.ptr(fl), _lib.ptr(fg), _lib.ptr(ul), st), "recover_members")
            sec_tables = (self.d_n1, self.d_n2, self.d_rel_idx, self.d_rel)
            if link_forces is not None:
                # LinearMpc (engine._MpcDevice): end forces of the rigid links; end forces of the members that sat on a slave
                # node carried from the master back to that node.  The section forces below are then formed on the ORIGINAL
                # connectivity and offsets (report tables), from these node forces.
                rows, lam_l, lam_g, ends, sec_tables = link_forces
                fl.index_copy_(0, rows, lam_l)
                fg.index_copy_(0, rows, lam_g)
                if ends is not None:
                    e_rows, e_end, e_arm, e_R = ends                     # member rows, 0 / 1, arm x_slave - x_master, member frames
                    g = fg.view(nm, 2, 6, nr)
                    F = g[e_rows, e_end, :3]
                    M = g[e_rows, e_end, 3:] - _MpcDevice._cross(e_arm, F)
                    g[e_rows, e_end, 3:] = M
                    fl.view(nm, 2, 2, 3, nr)[e_rows, e_end, 1] = torch.einsum("mij,mjr->mir", e_R, M)
            if cables is not None:
                # catenary cables: end forces of the converged state, formed on the host (_cable_report)
                fl.index_copy_(0, cables["rows"], cables["f_local"].expand(-1, -1, nr).contiguous())
                fg.index_copy_(0, cables["rows"], cables["f_global"].expand(-1, -1, nr).contiguous())
            out.update(f_local=fl, f_global=fg, u_local=ul)
            if self.n_warp:
                bw = self._empty((self.n_warp, 2, nr))
                _lib.check(lib.fers_b200_recover_warping(
                    self.n_warp, nr, _lib.ptr(self.d_w_member), _lib.ptr(self.d_w_aux), _lib.ptr(d_xyz), _lib.ptr(self.d_n1),
                    _lib.ptr(self.d_n2), _lib.ptr(self.d_m_prop), _lib.ptr(d_scale), _lib.ptr(self.d_prop), _lib.ptr(d_u),
                    _lib.ptr(fl), _lib.ptr(fg), _lib.ptr(bw), st), "recover_warping")
                out["bw"] = bw
            if detail:
                sp_ptr, sp_rows = span
                sec = self._empty((nm, 18, nr))
                mm = [self._empty((nm, 6, nr)) for _ in range(4)]
                d_fac = self._t(factors_pad, np.float64)
                _lib.check(lib.fers_b200_recover_sections(
                    nm, nr, factors_pad.shape[1], _lib.ptr(d_xyz), _lib.ptr(sec_tables[0]), _lib.ptr(sec_tables[1]),
                    _lib.ptr(self.d_roll), _lib.ptr(sp_ptr), _lib.ptr(sp_rows), _lib.ptr(d_fac), _lib.ptr(fl), _lib.ptr(fg),
                    _lib.ptr(sec_tables[2]), _lib.ptr(sec_tables[3]), _lib.ptr(sec), _lib.ptr(mm[0]), _lib.ptr(mm[1]), _lib.ptr(mm[2]), _lib.ptr(mm[3]), st),
                    "recover_sections")
                if cables is not None:
                    for dst, key in ((sec, "sec"), (mm[0], "lmax"), (mm[1], "lmin"), (mm[2], "gmax"), (mm[3], "gmin")):
                        dst.index_copy_(0, cables["rows"], cables[key].expand(-1, -1, nr).contiguous())
                out.update(sec=sec, lmax=mm[0], lmin=mm[1], gmax=mm[2], gmin=mm[3])
        if m.n_plates:
            npl = m.n_plates
            pres = self._empty((npl, 8, nr)); pnod = self._empty((npl, 18, nr)); pcd = self._empty((npl, 12, nr))
            _lib.check(lib.fers_b200_recover_plates(npl, nr, _lib.ptr(d_xyz), _lib.ptr(self.d_p_nodes),
                                                    _lib.ptr(self.d_p_data), _lib.ptr(self.d_p_flags), _lib.ptr(d_u),
                                                    _lib.ptr(pres), _lib.ptr(pnod), _lib.ptr(pcd), st), "recover_plates")
            out.update(p_res=pres, p_nodal=pnod, p_cd=pcd)
        return out

    def run_linear(self, factors: np.ndarray, detail: bool = False, span=None) -> dict:
        """One pass of the linear hot path for the right-hand sides ``factors`` [ncols<=64, Nlc]:
        load vectors -> element stiffness + assembly + BC -> multi-RHS solve -> recovery.
        Everything stays on the device; returns the result tensors (layout [.., nr])."""
        nr = _pad_nr(factors.shape[0])
        if self.n_cables:
            raise EngineError("run_linear: models with cables need the iteration of engine.analyze")
        if factors.shape[0] > 64:
            raise EngineError("run_linear handles at most 64 right-hand sides per pass")
        f_case, peq_case = self.load_cases(self.d_xyz0)
        self.assemble(self.d_xyz0)
        d_f = self.combine(f_case, factors, nr)
        d_peq = self.combine(peq_case, factors, nr)
        d_u = self.solve(d_f)
        facp = np.zeros((nr, factors.shape[1]))
        facp[: factors.shape[0]] = factors
        if detail and span is None:
            span = self.span_tables()
        rec = self.recover(self.d_xyz0, d_u, d_f, d_peq, facp, span, detail=detail)
        rec["U"] = d_u
        return rec

    def reactions_at_supports(self, r: torch.Tensor):
        """(supported node indices [ns] on the host, reactions [ns, 6, nr] on the device): the rows of the reaction
        multi-vector that are not identically zero -- what a caller downloads instead of the whole [Ndof, nr] array."""
        nr = r.shape[-1]
        return self.sup_nodes, r.view(self.model.n_nodes, 6, nr).index_select(0, self.d_sup_nodes)

    def run_linear_sharded(self, factors: np.ndarray, gather=("U", "R", "f_local"), detail: bool = False) -> dict:
        """run_linear for all ranks of the default torch.distributed group at once: combinations sharded over the ranks,
        result tensors gathered with one all_gather each (fers_core_b200.dist.run_linear_sharded)."""
        from . import dist as _dist
        return _dist.run_linear_sharded(self, factors, gather, detail)

    def span_tables(self):
        """Span loads per load case as CSR over members: (sp_ptr [Nlc*(Nm+1)] int32, sp_rows [*,9])."""
        m = self.model
        nm, nlc = m.n_members, len(m.load_cases)
        ptr = np.zeros((max(nlc, 1), nm + 1), np.int32)
        rows = []
        base = 0
        for li, lc in enumerate(m.load_cases):
            mem = np.concatenate([lc.dl_member, lc.pl_member]).astype(np.int64)
            dat = np.concatenate([np.concatenate([np.zeros((len(lc.dl_member), 1)), lc.dl_data], 1),
                                  np.concatenate([np.ones((len(lc.pl_member), 1)), lc.pl_data], 1)], 0)
            order = np.argsort(mem, kind="stable")
            cnt = np.bincount(mem, minlength=nm) if len(mem) else np.zeros(nm, np.int64)
            ptr[li, 1:] = base + np.cumsum(cnt)
            ptr[li, 0] = base
            base += len(mem)
            rows.append(dat[order])
        rows = np.concatenate(rows, 0) if rows else np.zeros((0, 9))
        if rows.shape[0] == 0:
            rows = np.zeros((1, 9))
        return self._t(ptr.reshape(-1), np.int32), self._t(rows, np.float64)


# ---------------------------------------------------------------------------------------------
def analyze(model: wire.ModelArrays, device=None, detail: bool = True, assembly_variant: int = ASSEMBLY_GATHER,
            stats_out: dict | None = None, solver: str = "auto") -> ResultArrays:
    """Full analysis of one document on the GPU (twin of oracle.fers_oracle.analyze)."""
    t_start = time.perf_counter()
    if len(model.c_member):
        # catenary cables: outside their own element (cable.py) the member rows behave as pin-ended bars without stiffness
        # (loads the catenary cannot take -- partial / point loads -- go to the end nodes as on a truss member)
        model = dataclasses.replace(model, m_type=np.where(model.m_type == wire.MT_CABLE, wire.MT_TRUSS, model.m_type).astype(np.int32))
    # RigidStrategy.LinearMpc: the engine solves the transformed model (slaves on rigid arms of their master, hostlogic.
    # mpc_transform) and reports on the original one; where the transformation does not apply the rigid members are stiff
    # members and the result says so
    model_t, mpc = mpc_transform(model)
    eng = Engine(model if model_t is None else model_t, device, assembly_variant, solver)
    opts = model.options
    plan = rhs_plan(model)
    out = ResultArrays.empty(model, plan)
    if model_t is None and opts.rigid_strategy == "LinearMpc" and np.any(model.m_type == wire.MT_RIGID):
        out.warnings.append(("W_LINEAR_MPC_AS_RIGID_MEMBER", f"rigid_strategy=LinearMpc not applied ({mpc}): rigid members are "
                             "modelled as stiff members (RigidMember strategy)"))
    mpc_dev = _MpcDevice(eng, model, mpc) if model_t is not None else None
    if plan.n_rhs == 0:
        return out
    second_order = opts.order == "NONLINEAR"
    unilateral = np.isin(model.m_type, (wire.MT_TENSION, wire.MT_COMPRESSION)) | (len(model.c_member) > 0)
    span = eng.span_tables()
    if model.unity_checks and detail:
        out.unity = unity.UnityAccumulator(model, plan, eng.dev)      # envelopes folded batch by batch in _store
    for grp in plan.groups:
        xyz = imperfect_coordinates(model, grp.imperfection) if grp.imperfection is not None else model.xyz
        d_xyz = eng._t(xyz, np.float64) if grp.imperfection is not None else eng.d_xyz0
        out.unity_xyz = eng.d_xyz0          # deflected positions are reported on the undeformed coordinates
        cols = list(grp.columns)
        fac = plan.factors[cols]
        if not (second_order or unilateral.any() or eng.n_unilateral or eng.n_curves or eng.n_hinge_curves):
            f_case, peq_case = eng.load_cases(d_xyz)
            eng.assemble(d_xyz)
            for c0 in range(0, len(cols), 64):
                sub = cols[c0:c0 + 64]
                nr = _pad_nr(len(sub))
                facp = np.zeros((nr, fac.shape[1])); facp[: len(sub)] = fac[c0:c0 + 64]
                d_f = eng.combine(f_case, fac[c0:c0 + 64], nr)
                d_peq = eng.combine(peq_case, fac[c0:c0 + 64], nr)
                d_u = eng.solve(d_f)
                links = None
                if mpc_dev is not None:
                    links = mpc_dev.finish(d_u, eng.recover(d_xyz, d_u, d_f, d_peq, facp, None, detail=False), facp)
                rec = eng.recover(d_xyz, d_u, d_f, d_peq, facp, span, detail=detail, link_forces=links)
                _store(out, model, sub, d_u, rec, xyz)
        else:
            for ci, col in enumerate(cols):
                _solve_nonlinear(eng, out, model, d_xyz, xyz, col, fac[ci:ci + 1], span, second_order, detail)
    out.method = "Linear" if not second_order else "P_DELTA"
    if second_order and opts.nonlinear_method == "COROTATIONAL":
        out.warnings.append(("W_COROTATIONAL_AS_PDELTA",
                             "corotational Newton-Raphson is not implemented; second-order effects use P-Delta"))
    torch.cuda.synchronize(eng.dev)
    if model.hinge_caps and model.n_members:
        out.warnings += hinge_capacity_notes(model, out.f_local, plan.names)
    out.solve_ms = (time.perf_counter() - t_start) * 1e3
    if stats_out is not None:
        stats_out.update(vars(eng.stats))
    return out


class _MpcDevice:
    """Device side of RigidStrategy.LinearMpc after the solve of the transformed model: displacements of the slave nodes from
    their masters, end forces of the rigid links from the equilibrium of the slave nodes (the elastic members that sat on a
    slave report their end forces at the master; carried back over the arm they, minus the nodal loads of the slave, are what
    the link must hold)."""

    def __init__(self, eng: Engine, model, info):
        dev = eng.dev
        t = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
        self.info = info
        self.slaves, self.master = t(info.slaves, torch.int64), t(info.master, torch.int64)
        self.arm = t(info.arm)
        self.att_member, self.att_end, self.att_slave = (t(a, torch.int64) for a in (info.att_member, info.att_end, info.att_slave))
        self.ext = t(info.ext)                                               # [ns, 6, Nlc]
        self.slave_pos = {int(s): i for i, s in enumerate(info.slaves)}
        rows = np.array([m for m, _, _ in info.links], np.int64)
        self.rows = t(rows, torch.int64)
        R, _, _ = member_frames_host(model, rows)                           # frames of the links (original geometry)
        self.R = t(R)
        self.far_is_end = [int(model.m_n2[m]) == far for m, _, far in info.links]
        self.link_arm = t(np.array([model.xyz[far] - model.xyz[near] for _, near, far in info.links]).reshape(-1, 3))
        # report tables: connectivity and modifier records of the ORIGINAL model for the section-force kernel
        has_rel = bool(np.any(model.m_rel >= 0))
        self.sec_tables = (t(model.m_n1, torch.int32), t(model.m_n2, torch.int32),
                           t(model.m_rel, torch.int32) if has_rel else None, t(model.rel) if has_rel else None)
        self.ends = None
        if len(info.att_member):
            Ra, _, _ = member_frames_host(model, info.att_member)
            self.ends = (self.att_member, self.att_end, self.arm.index_select(0, self.att_slave), t(Ra))

    @staticmethod
    def _cross(a, b):
        """a [3] or [n,3] (constant over the columns) x b [.., 3, nr]."""
        ax, ay, az = a[..., 0, None], a[..., 1, None], a[..., 2, None]
        return torch.stack([ay * b[..., 2, :] - az * b[..., 1, :], az * b[..., 0, :] - ax * b[..., 2, :],
                            ax * b[..., 1, :] - ay * b[..., 0, :]], dim=-2)

    def finish(self, d_u, rec, facp):
        nr = d_u.shape[1]
        info = self.info
        ns = len(info.slaves)
        # displacements of the slaves: theta_s = theta_m, u_s = u_m + theta_m x arm
        um = d_u.view(-1, 6, nr).index_select(0, self.master)
        us = um.clone()
        us[:, :3] += self._cross(-self.arm, um[:, 3:])                      # theta x r = -(r x theta)
        d_u.view(-1, 6, nr).index_copy_(0, self.slaves, us)
        # demand at the slave nodes
        rho = torch.zeros((ns, 6, nr), dtype=torch.float64, device=d_u.device)
        if len(info.att_member):
            ends = rec["f_global"].view(-1, 2, 6, nr)[self.att_member, self.att_end]          # at the master node
            F = ends[:, :3]
            M = ends[:, 3:] - self._cross(self.arm.index_select(0, self.att_slave), F)         # carried to the slave
            rho.index_add_(0, self.att_slave, torch.cat([F, M], dim=1))
        fac = torch.from_numpy(np.ascontiguousarray(facp, np.float64)).to(d_u.device)          # [nr, Nlc]
        rho -= torch.einsum("sql,rl->sqr", self.ext[:, :, : fac.shape[1]], fac)
        # up the forest, deepest link first
        sub = {s: rho[i].clone() for s, i in self.slave_pos.items()}
        lam_g = torch.zeros((len(info.links), 12, nr), dtype=torch.float64, device=d_u.device)
        for li, (_m, near, far) in enumerate(info.links):
            D = sub[far]
            carried = torch.cat([D[:3], D[3:] + self._cross(self.link_arm[li], D[:3])], dim=0)
            if self.far_is_end[li]:
                lam_g[li, :6], lam_g[li, 6:] = carried, -D
            else:
                lam_g[li, :6], lam_g[li, 6:] = -D, carried
            if near in sub:
                sub[near] = sub[near] + carried
        lam_l = torch.einsum("mij,mbjr->mbir", self.R, lam_g.view(-1, 4, 3, nr)).reshape(-1, 12, nr)
        return self.rows, lam_l, lam_g, self.ends, self.sec_tables


def _solve_nonlinear(eng: Engine, out, model, d_xyz, xyz, col, fac, span, second_order, detail):
    """Active-set (tension/compression-only members) around the P-Delta fixed point, one column."""
    opts = model.options
    nm = model.n_members
    uni = np.nonzero(np.isin(model.m_type, (wire.MT_TENSION, wire.MT_COMPRESSION)))[0]
    scale = np.ones(nm)
    n_cab = eng.n_cables
    if n_cab:
        # catenary cables live outside the beam assembly (their member rows carry no stiffness); load per unstretched length
        # of this column: own weight + the uniform member loads of the combination
        scale[model.c_member] = 0.0
        cab_q = model.c_w + sum(fac[0, li] * lc.cab_q for li, lc in enumerate(model.load_cases))
        c_i, c_j = model.m_n1[model.c_member], model.m_n2[model.c_member]
    u_host = np.zeros((model.n_nodes, 6))
    d_cab_g = None
    kg_mode = (1 if opts.pdelta_formulation == "CONSISTENT" else 2) if second_order else 0
    sup = 0
    for ax in pdelta_suppressed_axes(model):
        sup |= 1 << ax
    nr = 8 if model.n_dof > DENSE_MAX_DOF else 1
    facp = np.zeros((nr, fac.shape[1])); facp[0] = fac[0]
    total_it, conv, resn = 0, True, 0.0
    d_axial = None
    rec = None
    sup_active = np.ones(eng.n_unilateral, bool)
    curve_k = curve_stiffness(model, None)              # load-dependent springs start at their zero-force stiffness
    hinge_k = hinge_curve_stiffness(model, None)        # = the values model.rel carries
    if eng.n_hinge_curves:
        eng.set_rel(model.rel)                          # the previous column left its own state on the device
    settled = False
    for _outer in range(max(1, 2 * (len(uni) + eng.n_unilateral) + 2
                            + (max(opts.max_iterations, 20) if eng.n_curves or eng.n_hinge_curves else 0)
                            + (CABLE_MAX_ITER if n_cab else 0))):
        if eng.n_unilateral or eng.n_curves:
            eng.set_unilateral(sup_active, curve_k)
        d_scale = eng._t(scale, np.float64) if (len(uni) or n_cab) else None
        f_case, peq_case = eng.load_cases(d_xyz, d_scale)
        d_f = eng.combine(f_case, fac, nr)
        d_peq = eng.combine(peq_case, fac, nr)
        if n_cab:
            # Newton step in total form: (K + Kc) u = f - p(u_k) + Kc u_k, the cable state (p, G) from the host
            p_c, g_c, _a = cable.element_state(xyz[c_i] + u_host[c_i, :3], xyz[c_j] + u_host[c_j, :3], cab_q, model.c_l0, model.c_ea)
            gu = np.einsum("cij,cj->ci", g_c, u_host[c_i, :3] - u_host[c_j, :3])
            corr = np.concatenate([gu, -gu], axis=1) - p_c
            d_cab_g = eng._t(g_c, np.float64)
            _lib.check(eng.lib.fers_b200_cable_rhs(n_cab, _lib.ptr(eng.d_c_member), _lib.ptr(eng.d_n1), _lib.ptr(eng.d_n2),
                                                   _lib.ptr(eng._t(corr, np.float64)), _lib.ptr(d_f), nr, 0, eng.stream), "cable_rhs")
        eng.assemble(d_xyz, d_scale, cable_g=d_cab_g)
        d_u = eng.solve(d_f)
        d_axial = None
        if second_order and nm:
            conv = False
            for _it in range(max(1, opts.max_iterations)):
                rec = eng.recover(d_xyz, d_u, d_f, d_peq, facp, span, d_scale, d_axial, kg_mode, sup, detail=False)
                ax_rhs = eng._empty((nm, nr))
                _lib.check(eng.lib.fers_b200_member_axial(nm, nr, _lib.ptr(rec["f_local"]), _lib.ptr(ax_rhs), eng.stream),
                           "member_axial")
                d_axial = ax_rhs
                d_ax1 = ax_rhs[:, 0].contiguous()
                eng.assemble(d_xyz, d_scale, d_ax1, kg_mode, sup, cable_g=d_cab_g)
                d_un = eng.solve(d_f)
                total_it += 1
                den = float(torch.linalg.norm(d_un[:, 0]))
                resn = float(torch.linalg.norm(d_un[:, 0] - d_u[:, 0])) / den if den > 0 else 0.0
                d_u = d_un
                if resn < opts.tolerance:
                    conv = True
                    break
        changed = False
        if n_cab:
            u_new = d_u[:, 0].cpu().numpy().reshape(model.n_nodes, 6)
            den = float(np.linalg.norm(u_new))
            changed = bool(den > 0.0 and float(np.linalg.norm(u_new - u_host)) > CABLE_TOL * den)
            u_host = u_new
        if eng.n_unilateral or eng.n_curves or eng.n_hinge_curves:
            # reactions and member end forces of the current state decide which unilateral supports stay closed and
            # how stiff the load-dependent springs (supports, hinges) are in the next pass
            now = eng.recover(d_xyz, d_u, d_f, d_peq, facp, span, d_scale, d_axial, kg_mode, sup, detail=False)
            r_host = now["R"][:, 0].cpu().numpy()
            if eng.n_unilateral:
                new_active = unilateral_update(model, sup_active, d_u[:, 0].cpu().numpy(), r_host)
                changed = not np.array_equal(new_active, sup_active)
                sup_active = new_active
            if eng.n_curves:
                new_k = curve_stiffness(model, r_host)
                changed = changed or not curves_converged(new_k, curve_k, opts.tolerance)
                curve_k = new_k
            if eng.n_hinge_curves:
                new_k = hinge_curve_stiffness(model, now["f_local"].view(nm, 12, -1)[:, :, 0].cpu().numpy())
                if not curves_converged(new_k, hinge_k, opts.tolerance):
                    changed = True
                    eng.set_rel(rel_with_hinge_curves(model, new_k))
                hinge_k = new_k
        if len(uni) == 0:
            if changed:
                continue
            settled = True
            break
        # elongation decides the active set
        u = d_u[:, 0].cpu().numpy().reshape(model.n_nodes, 6)
        d = xyz[model.m_n2[uni]] - xyz[model.m_n1[uni]]
        lx = d / np.linalg.norm(d, axis=1)[:, None]
        elong = ((u[model.m_n2[uni], :3] - u[model.m_n1[uni], :3]) * lx).sum(1)
        tens = np.isin(model.m_type[uni], (wire.MT_TENSION, wire.MT_CABLE))
        want = np.where(tens, elong >= 0.0, elong <= 0.0)
        new_scale = scale.copy()
        new_scale[uni] = np.where(want, 1.0, INACTIVE_SCALE)
        if np.array_equal(new_scale, scale) and not changed:
            settled = True
            break
        scale = new_scale
    if not settled:
        # the bound of the outer loop was reached with the active set (tension / compression-only members, unilateral
        # supports) or a stiffness curve still changing: the stored state is the last iterate, flagged as such
        conv = False
        out.converged[col] = False
        out.warnings.append(("W_ACTIVE_SET_NOT_CONVERGED", f"result column {col}: the active set / stiffness curves did not "
                             f"settle within {_outer + 1} outer iterations; the last iterate is reported"))
    d_scale = eng._t(scale, np.float64) if (len(uni) or n_cab) else None
    cab = _cable_report(eng, model, xyz, u_host, cab_q) if n_cab else None
    rec = eng.recover(d_xyz, d_u, d_f, d_peq, facp, span, d_scale, d_axial, kg_mode, sup, detail=detail, cables=cab)
    _store(out, model, [col], d_u, rec, xyz)
    if second_order:
        out.iterations[col] = total_it
        out.converged[col] = bool(conv) and settled
        out.residual[col] = resn


def _cable_report(eng: Engine, model, xyz, u, cab_q) -> dict:
    """Results of the catenary cables at the node displacements u [Nn,6] (one column): end forces in the stiffness sign
    (-a at the start, a - q L0 at the end), the start force carried along the unstretched length at the report stations
    (q s - a, no moments) and the extrema over {ends, stations}, as device tensors [Ncable, ., 1] for Engine.recover."""
    cm = model.c_member
    ni, nj = model.m_n1[cm], model.m_n2[cm]
    _p, _g, a = cable.element_state(xyz[ni] + u[ni, :3], xyz[nj] + u[nj, :3], cab_q, model.c_l0, model.c_ea)
    R, _, _ = member_frames_host(dataclasses.replace(model, xyz=xyz), cm)
    nc = len(cm)
    fg = np.zeros((nc, 12))
    fg[:, 0:3] = -a
    fg[:, 6:9] = a - cab_q * model.c_l0[:, None]
    fl = np.zeros((nc, 12))
    fl[:, 0:3] = np.einsum("mij,mj->mi", R, fg[:, 0:3])
    fl[:, 6:9] = np.einsum("mij,mj->mi", R, fg[:, 6:9])
    tg = -cable.tension_along(a, cab_q, model.c_l0, SECTION_FRACS)                # [nc, ns, 3] global
    sec_g = np.zeros((nc, len(SECTION_FRACS), 6)); sec_g[:, :, 0:3] = tg
    sec = np.zeros_like(sec_g); sec[:, :, 0:3] = np.einsum("mij,msj->msi", R, tg)
    loc = np.concatenate([fl[:, None, 0:6], fl[:, None, 6:12], sec], axis=1)
    glo = np.concatenate([fg[:, None, 0:6], fg[:, None, 6:12], sec_g], axis=1)
    t = lambda x: eng._t(np.ascontiguousarray(x).reshape(nc, -1, 1), np.float64)          # noqa: E731
    return dict(rows=eng.d_c_member.to(torch.int64), f_local=t(fl), f_global=t(fg), sec=t(sec), lmax=t(loc.max(1)),
                lmin=t(loc.min(1)), gmax=t(glo.max(1)), gmin=t(glo.min(1)))


def _store(out: ResultArrays, model, cols, d_u, rec, xyz) -> None:
    """Device results [.., nr] -> host result arrays for the given rhs columns.  The column-major device layout is
    transposed on the device and copied straight into the destination rows (one D2H per field, no host-side
    gather) when the columns are consecutive."""
    n = len(cols)
    cols = np.asarray(cols)
    nn, nm = model.n_nodes, model.n_members
    c0 = int(cols[0]) if n else 0
    consecutive = n > 0 and bool(np.array_equal(cols, np.arange(c0, c0 + n)))

    def dev(t, shape):
        return t.view(shape + (t.shape[-1],)).movedim(-1, 0)[:n]

    def host(t, shape):
        return dev(t, shape).contiguous().cpu().numpy()

    pending = []

    def put(dst, t, shape):
        # D2H through pinned staging (torch's caching host allocator makes it cheap after the first pass), all
        # fields in flight at once, one synchronisation; pageable destinations would serialise at a few GB/s
        src = dev(t, shape).contiguous()
        if src.numel() * 8 <= STAGE_MAX_BYTES:
            stage = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            stage.copy_(src, non_blocking=True)
            pending.append((dst, stage))
        elif consecutive and dst[c0:c0 + n].flags.c_contiguous:
            torch.from_numpy(dst[c0:c0 + n]).copy_(src)
        else:
            dst[cols] = src.cpu().numpy()

    def flush():
        if pending:
            torch.cuda.current_stream(d_u.device).synchronize()
        for dst, stage in pending:
            if consecutive:
                dst[c0:c0 + n] = stage.numpy()
            else:
                dst[cols] = stage.numpy()
        pending.clear()

    if out.unity:
        out.unity.accumulate([int(c) for c in cols], out.unity_xyz, d_u, rec,
                             torch.cuda.current_stream(d_u.device).cuda_stream)
    put(out.U, d_u, (nn, 6))
    sup_pending = None
    if "sup_nodes" in rec and 2 * len(rec["sup_nodes"][0]) < nn:
        # reactions: only the rows of the supported nodes are downloaded (the rest of out.R is, and stays, zero)
        sup_host, sup_dev = rec["sup_nodes"]
        src = rec["R"].view(nn, 6, rec["R"].shape[-1]).index_select(0, sup_dev).movedim(-1, 0)[:n].contiguous()
        stage_r = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
        stage_r.copy_(src, non_blocking=True)
        sup_pending = (sup_host, stage_r)
    else:
        put(out.R, rec["R"], (nn, 6))
    if nm:
        put(out.f_local, rec["f_local"], (nm, 12))
        if "bw" in rec:
            out.bw[np.asarray(cols)[:, None], model.w_member[None, :]] = host(rec["bw"], (len(model.w_member), 2))
        put(out.f_global, rec["f_global"], (nm, 12))
        put(out.u_local, rec["u_local"], (nm, 12))
        if "sec" in rec:
            put(out.sec, rec["sec"], (nm, len(SECTION_FRACS), 6))
            put(out.lmax, rec["lmax"], (nm, 6)); put(out.lmin, rec["lmin"], (nm, 6))
            put(out.gmax, rec["gmax"], (nm, 6)); put(out.gmin, rec["gmin"], (nm, 6))
    if model.n_plates:
        npl = model.n_plates
        put(out.p_res, rec["p_res"], (npl, 8))
        put(out.p_nodal, rec["p_nodal"], (npl, 3, 6))
        cd = host(rec["p_cd"], (npl, 12))
        out.p_cdg[cols] = cd[..., :6]; out.p_cdl[cols] = cd[..., 6:]
        out.p_centroid = xyz[model.p_nodes].mean(axis=1)
    flush()
    if sup_pending is not None:
        torch.cuda.current_stream(d_u.device).synchronize()
        sup_host, stage_r = sup_pending
        out.R[np.asarray(cols)[:, None], sup_host[None, :]] = stage_r.numpy()


class FrameBatch:
    """BASELINE.json configs[4]: a batch of independent equal-sized frames (one concatenated, block-diagonal model, nodes
    numbered frame by frame), one combination, linear or P-Delta, resident on one GPU.

    Every frame is factorised by the batched dense Cholesky; the P-Delta fixed point runs for all frames at once and each
    frame stops at ITS OWN convergence (converged frames are frozen on the device), so the result equals analysing the
    frames one by one.  The host never waits inside the loop: the "any frame still active" flag of iteration i travels to
    pinned memory asynchronously and is looked at after the cheap first half of iteration i + 1 (axial forces, tangent
    matrix) has been enqueued -- at most that half is enqueued in vain, and it changes nothing because every frame is frozen
    by then; the factorisations are only enqueued once the flags say they are needed."""

    def __init__(self, model: wire.ModelArrays, nodes_per_frame: int, device=None):
        self.eng = eng = Engine(model, device)
        self.model = model
        self.plan = rhs_plan(model)
        if self.plan.n_rhs != 1:
            raise EngineError("FrameBatch expects exactly one right-hand side (one combination, no load cases)")
        if model.n_nodes % nodes_per_frame or model.n_plates or len(model.c_member):
            raise EngineError("FrameBatch: model is not a batch of equal-sized beam frames (plates and cables take engine.analyze)")
        self.nodes_per_frame = nodes_per_frame
        self.nf = model.n_nodes // nodes_per_frame
        self.n = 6 * nodes_per_frame
        if model.n_members % self.nf:
            raise EngineError("FrameBatch: members are not evenly distributed over the frames")
        self.nmf = model.n_members // self.nf
        grp = self.plan.groups[0]
        self.xyz = imperfect_coordinates(model, grp.imperfection) if grp.imperfection is not None else model.xyz
        self.d_xyz = eng._t(self.xyz, np.float64) if grp.imperfection is not None else eng.d_xyz0
        opts = model.options
        self.second_order = opts.order == "NONLINEAR"
        self.kg_mode = (1 if opts.pdelta_formulation == "CONSISTENT" else 2) if self.second_order else 0
        self.sup = 0
        for ax in pdelta_suppressed_axes(model):
            self.sup |= 1 << ax
        self.fac = self.plan.factors[[0]]
        self.facp = np.zeros((1, self.fac.shape[1])); self.facp[0] = self.fac[0]
        nf, n = self.nf, self.n
        self.a = eng._empty((nf, n, n))
        self.flag = eng._zeros(1, torch.int32)
        self.info_dev = eng._zeros(nf, torch.int32)
        self.bad_dev = torch.zeros(nf, dtype=torch.bool, device=eng.dev)
        self.h_flags_i32 = torch.zeros((max(opts.max_iterations, 1) + 2, 2), dtype=torch.int32).pin_memory()
        self.stats: dict = {}

    # asynchronous: no host synchronisation; frames whose `active` flag is 0 are skipped; failed pivots and members that
    # cross frames are collected on the device and looked at where the host synchronises anyway
    def _factor_solve(self, rhs, active=None):
        eng, lib, st = self.eng, self.eng.lib, self.eng.stream
        u = rhs.clone()
        _lib.check(lib.fers_b200_frame_batch_factor_solve(
            self.nodes_per_frame, self.nf, _lib.ptr(eng.d_rowptr), _lib.ptr(eng.d_colidx), _lib.ptr(eng.d_vals), _lib.ptr(eng.d_fixed),
            _lib.ptr(eng.d_ks), _lib.ptr(self.a), _lib.ptr(active), _lib.ptr(self.info_dev), _lib.ptr(self.flag), _lib.ptr(u), st),
            "frame_batch_factor_solve")
        return u

    def _raise_if_bad(self):
        if int(self.flag.item()) != 0:
            raise EngineError("FrameBatch: a member connects two frames; the model is not a batch of independent frames")
        bad = torch.nonzero(self.bad_dev).reshape(-1).cpu().tolist()
        if bad:
            raise EngineError(f"{len(bad)} frame(s) of the batch are not positive definite (first: frame {bad[0]}): "
                              "buckled under the applied load or insufficiently supported")

    def solve(self, detail: bool = False) -> dict:
        """Device work of one analysis: load vectors, assembly, factorisations, P-Delta loop, recovery.  Returns the
        result tensors (device) of Engine.recover plus "U"; self.stats holds iteration counts / residuals (device)."""
        eng, model, lib = self.eng, self.model, self.eng.lib
        opts = model.options
        nf, n, nm, nmf = self.nf, self.n, model.n_members, self.nmf
        d_xyz, kg_mode, sup = self.d_xyz, self.kg_mode, self.sup
        self.flag.zero_(); self.bad_dev.zero_()
        f_case, peq_case = eng.load_cases(d_xyz)
        d_f = eng.combine(f_case, self.fac, 1)
        d_peq = eng.combine(peq_case, self.fac, 1)
        eng.assemble(d_xyz)
        rhs = (d_f * (1 - eng.d_fixed.to(torch.float64))[:, None]).contiguous()      # the mask assemble() builds does not change
        d_u = self._factor_solve(rhs)
        self.bad_dev.logical_or_(self.info_dev != 0)
        d_axial = None
        iters = torch.zeros(nf, dtype=torch.int64, device=eng.dev)
        resid = torch.zeros(nf, dtype=torch.float64, device=eng.dev)
        active = torch.ones(nf, dtype=torch.bool, device=eng.dev)
        enqueued = 0
        if self.second_order and nm:
            events = []
            max_it = max(1, opts.max_iterations)
            d_axial = eng._empty((nm, 1))
            d_flags = torch.zeros((max_it, 2), dtype=torch.int32, device=eng.dev)
            h_flags = self.h_flags_i32
            h_flags.zero_()
            for it in range(max_it):
                # The host runs one iteration ahead of the flags it reads: iteration `it` is enqueued as soon as the flags of
                # iteration it - 2 are known, so the GPU never waits for the host.  An iteration enqueued after every frame
                # has frozen costs a few empty launches: every kernel of the loop skips frozen frames.
                if it >= 2:
                    events[it - 2].synchronize()
                    go_on, bad = (int(v) for v in h_flags[it - 2])
                    if bad:
                        self._raise_if_bad()
                    if not go_on:
                        break
                # axial forces of the current state (frozen frames keep theirs), tangent matrix, factorisation, substitution
                rec = eng.recover(d_xyz, d_u, d_f, d_peq, self.facp, None, None, d_axial if it else None, kg_mode, sup, detail=False)
                _lib.check(lib.fers_b200_member_axial_masked(nm, nmf, _lib.ptr(rec["f_local"]), _lib.ptr(active) if it else None,
                                                             _lib.ptr(d_axial), eng.stream), "member_axial_masked")
                eng.assemble(d_xyz, None, d_axial[:, 0], kg_mode, sup, count=False)
                d_un = self._factor_solve(rhs, active)
                # convergence test, new state of the active frames, (go on, bad pivot) of this iteration -> pinned memory
                _lib.check(lib.fers_b200_frame_batch_step(nf, n, _lib.ptr(d_un), _lib.ptr(d_u), _lib.ptr(active), _lib.ptr(iters),
                                                          _lib.ptr(resid), float(opts.tolerance), _lib.ptr(self.info_dev),
                                                          _lib.ptr(self.bad_dev), _lib.ptr(d_flags[it]), eng.stream), "frame_batch_step")
                h_flags[it].copy_(d_flags[it], non_blocking=True)
                ev = torch.cuda.Event(); ev.record(torch.cuda.current_stream(eng.dev)); events.append(ev)
                enqueued = it + 1
            # flags not looked at yet (the last one or two iterations enqueued)
            for j in range(max(0, enqueued - 2), enqueued):
                events[j].synchronize()
                if int(h_flags[j][1]):
                    self._raise_if_bad()
            # the matrix in eng.d_vals must be the tangent of the frozen state for the reactions
            eng.assemble(d_xyz, None, d_axial[:, 0], kg_mode, sup, count=False)
        span = eng.span_tables() if detail else None
        rec = eng.recover(d_xyz, d_u, d_f, d_peq, self.facp, span, None, d_axial, kg_mode, sup, detail=detail)
        rec["U"] = d_u
        self.stats = {"iters": iters, "resid": resid, "active": active, "iterations_enqueued": enqueued}
        return rec

    def results(self, rec: dict) -> ResultArrays:
        """Host result arrays of one solve() (D2H of every field) + the convergence record."""
        model = self.model
        self._raise_if_bad()
        out = ResultArrays.empty(model, self.plan)
        _store(out, model, [0], rec["U"], rec, self.xyz)
        out.method = "Linear" if not self.second_order else "P_DELTA"
        it_host = self.stats["iters"].cpu().numpy()
        out.iterations[0] = int(it_host.max()) if self.second_order else -1
        out.converged[0] = not bool(self.stats["active"].any()) if self.second_order else True
        out.residual[0] = float(self.stats["resid"].max()) if self.second_order else 0.0
        self.frame_iterations = it_host
        return out


def analyze_frame_batch(model: wire.ModelArrays, nodes_per_frame: int, device=None, detail: bool = False,
                        stats_out: dict | None = None) -> ResultArrays:
    """One batch of frames, host arrays in and out (FrameBatch: upload, solve, download)."""
    fb = FrameBatch(model, nodes_per_frame, device)
    out = fb.results(fb.solve(detail=detail))
    if stats_out is not None:
        stats_out.update(vars(fb.eng.stats))
        stats_out["frame_iterations"] = fb.frame_iterations
        stats_out["frame_residual"] = fb.stats["resid"].cpu().numpy()
        stats_out["frames"] = fb.nf
        stats_out["iterations_enqueued"] = fb.stats["iterations_enqueued"]
    return out


# ---- the drop-in boundary ---------------------------------------------------------------------
def solve_document(doc: dict, device=None) -> dict:
    """dict in -> dict out (input document with ``results`` filled)."""
    model = wire.parse_document(doc)
    res = analyze(model, device)
    out = dict(doc)
    out["results"] = results_to_wire(model, res)
    return out


def solve_lazy(doc: dict, device=None):
    """dict in -> lazy ``ResultsBundle`` view over the result arrays (results.lazy_bundle): for models whose JSON
    results would not fit anywhere.  Same keys and numbers as ``solve_document(doc)["results"]``."""
    model = wire.parse_document(doc)
    return lazy_bundle(model, analyze(model, device))


def solve_fers(calc, device=None, lazy: bool = True):
    """Input fast path of the drop-in boundary: a reference ``FERS`` object in, results out, without the document round trip
    of ``FERS.run_analysis`` (wire.from_fers: the object graph straight to arrays).  Returns the lazy ``ResultsBundle`` view
    (results.lazy_bundle) or, with lazy=False, the result document as a dict (``results`` key of calculate_from_json)."""
    model = wire.from_fers(calc)
    res = analyze(model, device)
    return lazy_bundle(model, res) if lazy else results_to_wire(model, res)


def calculate_from_json(input_json: str) -> str:
    """Drop-in for ``fers_calculations.calculate_from_json`` (fers_core/fers/fers.py:300).  The result document is
    written as text straight from the result arrays (results.results_to_json): same document as
    ``json.dumps(solve_document(doc))``, several times faster once a model has a few thousand members."""
    doc = json.loads(input_json)
    model = wire.parse_document(doc)
    res = analyze(model)
    head = json.dumps({k: v for k, v in doc.items() if k != "results"})
    return head[:-1] + (', ' if len(head) > 2 else '') + '"results": ' + results_to_json(model, res) + "}"


def calculate_from_file(file_path: str) -> str:
    """Drop-in for ``fers_calculations.calculate_from_file`` (fers_core/fers/fers.py:260)."""
    with open(file_path, "r", encoding="utf-8") as fh:
        return calculate_from_json(fh.read())
