#!/usr/bin/env python
"""bench.py -- headline benchmark of the FERS analysis hot path on B200.

Metric (BASELINE.json): solved DOF x load-combinations / second on a 3D frame, with the
SpMM/assembly kernels reported against the HBM roofline.  Workload = BASELINE.json configs[3]:
synthetic 3D space frame 60x60 bays x 30 storeys (692 106 DOF), 8 load cases, 64 combinations,
linear static.  A "step" is one pass of the hot path over one batch of 64
combinations: load-vector assembly -> element stiffness + global assembly + BC -> 64-RHS solve ->
recovery (member end forces, reactions).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference ...                      (CPU arm: the oracle on host cores)

Multi-GPU: independent batches of 64 combinations per GPU (weak scaling, no data-path
collective); one NCCL all_gather collects a per-combination result summary.
"""
from __future__ import annotations

import argparse
import ctypes as C
import glob
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SYRK_KERNEL_ID = "k_syrk<G86>"          # Schur-update kernel of the default path; ncu captures under profiles/ name the one they saw
METRIC = "solved_dof_x_load_combinations_per_second"
UNIT = "DOF*combos/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--nx", type=int, default=60)
    p.add_argument("--nz", type=int, default=60)
    p.add_argument("--ny", type=int, default=30)
    p.add_argument("--combos", type=int, default=64)
    p.add_argument("--workload", default="cfg4", choices=["cfg4", "cfg5"],
                   help="cfg4 = configs[3] space frame x 64 combinations (headline); cfg5 = configs[4] batch of "
                        "parametric frames, sway imperfection, P-Delta")
    p.add_argument("--frames", type=int, default=1024, help="cfg5: frames per GPU")
    p.add_argument("--assembly-variant", type=int, default=4,
                   help="4 = gather kernel (default, no atomics); 0-3 = atomic scatter-add variants")
    p.add_argument("--spmm-variant", type=int, default=-1, help="0 register/LDG, 1 TMA-staged, 2 TMA warp-specialised (default: library default)")
    p.add_argument("--solver", default="auto", choices=["auto", "cholesky", "pcg"],
                   help="auto = sparse multifrontal Cholesky above 8 192 DOF (default); pcg = block-Jacobi PCG")
    p.add_argument("--scaling", default="weak", choices=["weak", "strong", "tree"],
                   help="weak = every GPU solves its own batch of --combos combinations (default, the driver's scaling run); strong = ONE "
                        "model x --combos combinations split over the GPUs (replicated factorisation, result tensors all-gathered); "
                        "tree = ONE model x --combos combinations, the sparse factorisation itself split over 2^k GPUs by the "
                        "subtrees of the front tree (Schur complements over NVLink; chol.TreeParallelSolver)")
    p.add_argument("--detail", action="store_true",
                   help="cfg4: the timed step also recovers section forces along the members and their extrema "
                        "(run_linear(detail=True)); default: member end forces, displacements and reactions only")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--profile-iters", type=int, default=0,
                   help="cap PCG at this many iterations and accept the unconverged result (ncu captures only)")
    return p.parse_args()


def measured_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        d["_source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "_source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_model(args, rank: int):
    from fers_core_b200 import synth
    model = synth.grid_frame(args.nx, args.nz, args.ny, n_combos=args.combos)
    if rank > 0:   # each GPU solves its own batch of combinations
        rng = np.random.default_rng(synth.COMBO_SEED + rank)
        f = synth.COMBO_LEVELS[rng.integers(0, len(synth.COMBO_LEVELS), size=model.combo_factors.shape)]
        f[:, 0] = np.where(f[:, 0] == 0.0, 1.0, f[:, 0])
        model.combo_factors = f
    return model


def pin_model(model):
    """Back the large input arrays by pinned host memory (source of the timed H2D copies)."""
    import torch

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy()
    for name in ("xyz", "m_n1", "m_n2", "m_type", "m_roll", "m_prop", "m_rel", "bc_code", "bc_k"):
        setattr(model, name, pin(getattr(model, name)))
    for lc in model.load_cases:
        for name in ("nod_node", "nod_vec", "dl_member", "dl_data", "pl_member", "pl_data"):
            if len(getattr(lc, name)):
                setattr(lc, name, pin(getattr(lc, name)))
    return model


def workload_name(args, n_comb: int) -> str:
    return (f"configs[3] synthetic 3D space frame {args.nx}x{args.nz} bays x {args.ny} storeys, linear static, "
            f"{n_comb} combinations per GPU"
            + (", section forces along the members, their extrema and a Generic unity check on every member evaluated in the timed step"
               if getattr(args, "detail", False) else ""))


def cpu_pass(args):
    """The CPU path on THIS run's workload (same frame, same combinations): oracle/cpu_frame -- C++ / OpenMP restatement
    of the oracle with a multifrontal Cholesky on BLAS-3 (OpenBLAS), every host thread.  Returns (CpuFrame, model)."""
    from fers_core_b200 import synth
    from oracle import cpu_frame
    model = synth.grid_frame(args.nx, args.nz, args.ny, n_combos=args.combos)
    return cpu_frame.CpuFrame(model), model


CPU_KIND_NOTE = ("C++ / OpenMP restatement of the reference path (oracle/cpu_frame.cpp): element stiffness, assembly, load "
                 "vectors, multifrontal Cholesky with OpenBLAS dpotrf / dtrsm / dsyrk / dgemm on all host threads, 64-RHS "
                 "substitution, member forces and reactions; the reference's own engine is a closed Rust wheel absent from "
                 "the reference tree")


def run_reference(args, rank: int, world: int) -> None:
    """CPU arm: the reference's engine (closed Rust wheel fers_calculations==0.2.43) cannot be installed here, so the CPU
    restatement of the path is timed on the host cores -- on the SAME workload as the GPU arm (configs[3] itself).  One
    step = one full pass.  A pass takes tens of seconds, so the number of timed steps is capped by a wall budget
    (FERS_REF_BUDGET_S, default 240 s; 0 = run exactly --steps): the line reports the steps actually timed."""
    if rank != 0:
        return
    from oracle import cpu_frame
    cf, model = cpu_pass(args)
    factors = model.combo_factors
    units = model.n_dof * factors.shape[0]
    budget = float(os.environ.get("FERS_REF_BUDGET_S", "240"))
    t_begin = time.perf_counter()
    times, phases = [], {}

    def one_pass(timed: bool) -> float:
        t0 = time.perf_counter()
        cf.run_linear(factors)
        dt = time.perf_counter() - t0
        if timed:
            times.append(dt)
            for k, v in cf.timings.items():
                phases[k] = phases.get(k, 0.0) + v
        return dt
    # the first pass (symbolic plan, page faults) is a warm-up pass unless --warmup 0 asks for none; its duration sizes
    # the rest of the run against the wall budget
    t_first = one_pass(timed=args.warmup == 0)
    warm, steps = args.warmup, args.steps
    if budget > 0:
        room = max(budget - t_first, 0.0)
        afford = int(room // max(t_first, 1e-9))                     # further passes the budget allows
        steps = max(1, min(args.steps, afford + (1 if args.warmup == 0 else 0)))
        warm = 0 if args.warmup == 0 else max(1, min(args.warmup, afford - steps + 1))
    for _ in range(max(warm - 1, 0)):
        one_pass(False)
    while len(times) < steps:
        one_pass(True)
    phases = {k: v / steps for k, v in phases.items()}
    ms = 1e3 * float(np.mean(times))
    value = units / (ms / 1e3)
    cores = cpu_frame.threads()
    sample = (f"the workload itself: {args.nx}x{args.nz}x{args.ny} frame ({model.n_dof} DOF) x {factors.shape[0]} combinations, one full "
              f"pass per step on {cores} host threads; {steps} timed step(s) after {warm} warm-up pass(es) "
              f"(requested {args.steps} / {args.warmup}; wall budget {budget:g} s, FERS_REF_BUDGET_S)")
    emit(({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, factors.shape[0]), "nodes": model.n_nodes, "members": model.n_members,
                   "dof": model.n_dof, "load_cases": len(model.load_cases), "combinations_per_gpu": int(factors.shape[0]),
                   "sample": sample, "cpu_path": CPU_KIND_NOTE, "max_rel_residual": cf.relres,
                   "phase_seconds": {k: round(v, 4) for k, v in phases.items()},
                   "factor_gflops": cf.flops / 1e9 / phases["factor_s"], "wall_s": time.perf_counter() - t_begin},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "host_cores": os.cpu_count() or 1, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_b200(args, rank: int, world: int, local_rank: int) -> None:
    import torch
    import torch.distributed as dist
    from fers_core_b200 import _lib, engine
    from fers_core_b200 import dist as fdist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    if args.spmm_variant >= 0:
        _lib.check(lib.fers_b200_set_spmm_variant(args.spmm_variant), "set_spmm_variant")
    model = pin_model(build_model(args, rank))
    factors = model.combo_factors
    n_dof, n_comb = model.n_dof, factors.shape[0]
    units = n_dof * n_comb

    if args.profile_iters > 0:
        engine.PCG_MAX_ITER = args.profile_iters
        engine.PCG_ALLOW_UNCONVERGED = True
    eng = engine.Engine(model, dev, args.assembly_variant, args.solver)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    strong = args.scaling == "strong" and world > 1
    if strong:
        # ONE model x n_comb combinations over all ranks: every rank factorises K and solves its shard of the columns; the
        # result tensors of all combinations are all-gathered (NCCL over NVLink) inside the timed step
        model.combo_factors = build_model(args, 0).combo_factors
        factors = model.combo_factors
        units = n_dof * n_comb / world                      # `value` multiplies by world below

        def step_with(e):
            out = e.run_linear_sharded(factors, gather=("U", "R", "f_local"))
            return {"U": out["U"], "R": out["R"], "f_local": out["f_local"]}
    else:
        span_cache = {}
        unity_acc = {}
        if args.detail:
            # a Generic unity check on every member (N / N_Rd + |Mz| / M_Rd with constant resistances), evaluated on the
            # device over all combinations inside the timed step (row A13): the check is compiled once, its envelope reset per step
            from fers_core_b200 import hostlogic, unity
            q = lambda kind, **kw: {"Quantity": {kind: kw}}
            model.unity_checks = [{"id": "nm", "name": "N + M interaction", "applies_to": {"type": "AllMembers"},
                                   "spec": {"Generic": {"demand": "abs(N) / NRd + M / MRd", "capacity": "1", "variables": [
                                       {"name": "N", "source": q("MemberForce", component="N", aggregation={"type": "MaxAbs"})},
                                       {"name": "M", "source": q("MemberForce", component="Mz", aggregation={"type": "MaxAbs"})},
                                       {"name": "NRd", "source": {"Quantity": {"Constant": 3.5e6}}},
                                       {"name": "MRd", "source": {"Quantity": {"Constant": 4.6e5}}}]}}}]
            unity_plan = hostlogic.rhs_plan(model)

        def step_with(e):
            if args.detail:
                if id(e) not in span_cache:
                    span_cache[id(e)] = e.span_tables()
                    unity_acc[id(e)] = unity.UnityAccumulator(model, unity_plan, e.dev)
                rec_ = e.run_linear(factors, detail=True, span=span_cache[id(e)])
                acc = unity_acc[id(e)]
                acc.reset()
                acc.accumulate(list(range(factors.shape[0])), e.d_xyz0, rec_["U"], rec_, e.stream)
                rec_["unity_max_utilisation"] = float(acc._best_d[0][:, 0].nan_to_num(nan=-np.inf).max())     # D2H of one number
                return rec_
            return e.run_linear(factors)

    def one_step():
        return step_with(eng)

    # ---- device-resident steps -----------------------------------------------------------------
    for _ in range(args.warmup):
        rec = one_step()
    direct = eng.stats.method == "cholesky"
    sampler = ClockSampler(local_rank)
    if direct:
        lib.fers_b200_chol_set_profile(2)       # events around the Schur-update launches + totals only
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = lib.fers_b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        rec = one_step()
    e1.record()
    barrier()
    launches = lib.fers_b200_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = fdist.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    value = world * units / (ms_step / 1e3)
    timed_prof = eng._chol.profile() if direct else None
    lib.fers_b200_chol_set_profile(0)
    prof = (C.c_double * 6)()
    lib.fers_b200_pcg_profile(prof)
    iters = int(prof[4])
    relres_solver = eng.stats.relres
    refine_passes = eng.stats.iterations if direct else 0

    # summary gather over NVLink: per-combination max |u| from every rank
    summ = rec["U"].abs().amax(dim=0)[:n_comb].contiguous()
    all_summ = fdist.gather_summary(summ)            # [world * n_comb] over NCCL / NVLink
    assert all_summ.shape[0] == world * n_comb

    # true residual of the last step (free DOFs), on the device through the C ABI
    nr = rec["U"].shape[1]
    f_case, _ = eng.load_cases(eng.d_xyz0)
    d_f = eng.combine(f_case, factors, nr)
    ku = torch.empty_like(rec["U"])
    _lib.check(lib.fers_b200_spmm(model.n_nodes, nr, _lib.ptr(eng.d_rowptr), _lib.ptr(eng.d_colidx), _lib.ptr(eng.d_vals),
                                  _lib.ptr(eng.d_fixed), _lib.ptr(eng.d_ks), _lib.ptr(rec["U"]), _lib.ptr(ku),
                                  eng.stream), "spmm")
    free = (eng.d_fixed == 0).to(torch.float64)[:, None]
    resid = float((torch.linalg.norm((ku - d_f) * free, dim=0) / torch.linalg.norm(d_f * free, dim=0))[:n_comb].max())

    # ---- kernel roofline numbers ------------------------------------------------------------------
    peaks = measured_peaks()
    peak = float(peaks["hbm_gbs"])
    nnzb, nn = eng.nnzb, model.n_nodes
    V = n_dof * nr * 8
    spmm_bytes = nnzb * (36 * 8 + 4) + (nn + 1) * 4 + 2 * V
    kern = []

    def timed_launches(fn, reps=5, pre=None):
        # events around a burst of 8 back-to-back launches (one launch alone carries ~10 us of host launch gap between the
        # first event and the kernel: asm_probe, r2x); kernels with a preparation pass are timed one launch at a time.
        # fn(i): launch i of the burst works on copy i % 3 of the matrix values, so that no launch finds the lines of the
        # previous one in the 126 MB L2 (three copies of K = 0.67 GB at configs[3])
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        burst = 1 if pre is not None else 8
        out = []
        for _ in range(reps):
            if pre is not None:
                pre()
            a0.record()
            for _b in range(burst):
                fn(_b)
            a1.record()
            torch.cuda.synchronize(dev)
            out.append(a0.elapsed_time(a1) / burst)
        return float(np.median(out))

    def hbm_row(name, b, ms, launches_per_step):
        gbs = b / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"kernel": name, "bound": "hbm", "algorithmic_bytes": int(b), "ms": ms, "GB/s": gbs, "frac": gbs / peak,
                "share_of_step": ms * launches_per_step / ms_step if ms_step > 0 else 0.0}

    # assembly and SpMM kernels timed on their own (few launches per step): events around dedicated launches
    vals_rot = [eng.d_vals, eng.d_vals.clone(), eng.d_vals.clone()]
    if args.assembly_variant == engine.ASSEMBLY_GATHER:
        am = timed_launches(lambda i: _lib.check(lib.fers_b200_assemble_beams_gather(
            model.n_nodes, model.n_members, _lib.ptr(eng.d_xyz0), _lib.ptr(eng.d_n1), _lib.ptr(eng.d_n2),
            _lib.ptr(eng.d_m_prop), _lib.ptr(eng.d_type), _lib.ptr(eng.d_roll), None, None, _lib.ptr(eng.d_prop), eng.n_prop,
            None, 0, None, 0, 0, _lib.ptr(eng.d_m_slot), _lib.ptr(eng.d_hm_ptr), _lib.ptr(eng.d_hm_list),
            eng.n_dup_half_members, _lib.ptr(eng.d_diag_slot), _lib.ptr(eng.d_member_cache), _lib.ptr(vals_rot[i % 3]), eng.stream), "assemble_gather"))
        asm_name = "k_assemble_gather"
        # unique bytes: member records in, every block of K out once (the gather kernel has no zero pass, no atomics)
        asm_bytes = model.n_members * 64 + nnzb * 36 * 8
    else:
        am = timed_launches(lambda i: _lib.check(lib.fers_b200_assemble_beams(
            model.n_members, _lib.ptr(eng.d_xyz0), _lib.ptr(eng.d_n1), _lib.ptr(eng.d_n2), _lib.ptr(eng.d_m_prop),
            _lib.ptr(eng.d_type), _lib.ptr(eng.d_roll), None, None, _lib.ptr(eng.d_prop), eng.n_prop, None, 0, None, 0, 0,
            _lib.ptr(eng.d_m_slot), _lib.ptr(eng.d_vals), args.assembly_variant, eng.stream), "assemble"),
            pre=lambda: lib.fers_b200_zero(_lib.ptr(eng.d_vals), eng.d_vals.numel() * 8, eng.stream))
        asm_name = "k_assemble_beams (atomic scatter-add)"
        asm_bytes = model.n_members * (64 + 4 * 36 * 8)      # 4 blocks per member, RMW not counted
    if direct:
        sm = timed_launches(lambda i: _lib.check(lib.fers_b200_spmm(
            model.n_nodes, nr, _lib.ptr(eng.d_rowptr), _lib.ptr(eng.d_colidx), _lib.ptr(vals_rot[i % 3]), _lib.ptr(eng.d_fixed),
            _lib.ptr(eng.d_ks), _lib.ptr(rec["U"]), _lib.ptr(ku), eng.stream), "spmm"))
        kern.append(hbm_row("k_spmm64_tma (64 RHS; residual check / refinement)", spmm_bytes, sm, refine_passes + 1))
        x1 = rec["U"][:, 0].contiguous()
        y1 = torch.empty_like(x1)
        sv = timed_launches(lambda i: _lib.check(lib.fers_b200_spmm(
            model.n_nodes, 1, _lib.ptr(eng.d_rowptr), _lib.ptr(eng.d_colidx), _lib.ptr(vals_rot[i % 3]), _lib.ptr(eng.d_fixed),
            _lib.ptr(eng.d_ks), _lib.ptr(x1), _lib.ptr(y1), eng.stream), "spmv"))
        spmv_bytes = nnzb * (36 * 8 + 4) + (nn + 1) * 4 + 2 * n_dof * 8 + n_dof      # BSR-6 values + indices, x, y, BC mask
        kern.append(hbm_row("k_spmv (1 RHS, not on the step path)", spmv_bytes, sv, 0))
        kern.append(hbm_row(asm_name, asm_bytes, am, 1))
        if args.assembly_variant == engine.ASSEMBLY_GATHER:
            # SURVEY.md 8(d) counts the assembly's algorithmic bytes per MEMBER: 64 B read + four 6x6 blocks written (1 152 B),
            # atomics' read-modify-write not counted -- 0.403 GB at configs[3].  The gather kernel sums the diagonal blocks on
            # chip and writes every block of K once (0.245 GB, the figure `frac` above uses); both are reported.
            survey_bytes = model.n_members * (64 + 4 * 36 * 8)
            kern[-1].update({"survey_8d_algorithmic_bytes": int(survey_bytes),
                             "GB/s_survey_8d": survey_bytes / (am * 1e-3) / 1e9 if am > 0 else 0.0,
                             "frac_survey_8d": survey_bytes / (am * 1e-3) / 1e9 / peak if am > 0 else 0.0})
        # fp64 pipe peak measured live by the library (register-resident DMMA.8x8x4 / DFMA bursts)
        scratch = torch.empty(8 * 1024 * 160, dtype=torch.uint8, device=dev)
        dmma, dfma = C.c_double(), C.c_double()
        _lib.check(lib.fers_b200_fp64_peak(50.0, C.byref(dmma), C.byref(dfma), _lib.ptr(scratch), scratch.numel(),
                                           eng.stream), "fp64_peak")
        ex = eng._chol.plan.export()
        fr = ex["fronts"]
        k1 = 6.0 * fr["n1"].astype(np.float64)
        m2 = 6.0 * fr["n2"].astype(np.float64)
        n_fac = max(timed_prof["n_factor"], 1.0)
        # which fronts' Schur updates went through the int8 tensor-core kernel: the library's rule (chol_numeric.cu), level by level
        oz_S, oz_min_k = int(timed_prof["oz_setting"]) // 100000, int(timed_prof["oz_setting"]) % 100000
        on_oz = np.zeros(len(fr), bool)
        lp, lf = ex["level_ptr"], ex["level_fronts"]
        for l in range(len(lp) - 2):                      # the root level has no Schur complement
            ids = lf[lp[l]:lp[l + 1]]
            if oz_S >= 4 and 0 < len(ids) <= 64 and (k1[ids] >= oz_min_k).all() and (m2[ids] >= 1024).all():
                on_oz[ids] = True
        if abs(on_oz.sum() * n_fac - timed_prof["n_oz_fronts"]) > 0.5:
            raise RuntimeError(f"bench: {timed_prof['n_oz_fronts'] / n_fac} fronts took the int8 path, expected {on_oz.sum()}")
        syrk_fl = k1 * m2 * (m2 + 1)
        cls_flops = {"potrf": float((k1 ** 3 / 3).sum()), "panel": float((k1 * k1 * m2).sum()),
                     "syrk": float(syrk_fl[~on_oz].sum()), "syrk_int8": float(syrk_fl[on_oz].sum())}
        syrk_ms_per_fac = timed_prof["syrk_sum_ms"] / n_fac
        oz_ms_per_fac = timed_prof["oz_sum_ms"] / n_fac
        syrk_launches = max(timed_prof["n_syrk_launches"], 1.0)
        tf = cls_flops["syrk"] / (syrk_ms_per_fac * 1e-3) / 1e12 if syrk_ms_per_fac > 0 else 0.0
        # one more step with events around every kernel class (outside the timed region) for the breakdown
        lib.fers_b200_chol_set_profile(1)
        eng.run_linear(factors)
        full = eng._chol.profile()
        lib.fers_b200_chol_set_profile(0)
        breakdown = {k: full[k] for k in ("init_ms", "scatter_ms", "potrf_ms", "trsm_ms", "update_ms", "syrk_ms", "oz_ms",
                                           "factor_ms", "solve_ms", "factor_launches", "solve_launches")}
        upd_flops = cls_flops["panel"] + cls_flops["potrf"]
        pairs = oz_S * (oz_S + 1) // 2
        int8_peak = 2.0 * float(peaks.get("bf16_tflops", 1590.0))
        if oz_ms_per_fac > 0:
            tf_oz = cls_flops["syrk_int8"] / (oz_ms_per_fac * 1e-3) / 1e12
            kern.insert(0, {"kernel": f"k_oz_syrk_wide<{oz_S} digit planes, 128x128 tiles in two passes over K> + k_oz_split + k_oz_rowmax (Schur "
                                      f"complements of the upper tree levels, own columns >= {oz_min_k}: tcgen05.mma kind::i8 M=128 N=128, "
                                      f"{pairs} digit-plane products per fp64 product)",
                            "bound": "tensor", "algorithmic_flops_fp64": cls_flops["syrk_int8"],
                            "algorithmic_int8_ops": pairs * cls_flops["syrk_int8"], "ms": oz_ms_per_fac,
                            "fp64_equivalent_TFLOP/s": tf_oz, "vs_fp64_tensor_pipe_peak": tf_oz / dmma.value if dmma.value > 0 else None,
                            "int8_TOP/s": pairs * tf_oz, "frac": pairs * tf_oz / int8_peak, "fronts_per_factorisation": int(on_oz.sum()),
                            "share_of_step": oz_ms_per_fac / ms_step})
        kern.insert(1 if oz_ms_per_fac > 0 else 0,
                    {"kernel": "k_syrk<128x64 DMMA> (Schur complements of the other levels, F22 -= L21 L21^T)", "bound": "tensor",
                     "algorithmic_flops": cls_flops["syrk"], "ms": syrk_ms_per_fac, "TFLOP/s": tf,
                     "frac": tf / dmma.value if dmma.value > 0 else None,
                     "share_of_step": syrk_ms_per_fac / ms_step})
        um = full["update_ms"] + full["trsm_ms"] + full["potrf_ms"]
        # with look-ahead (the timed steps) the panel chain overlaps the trailing updates: effective panel time =
        # timed factorisation - timed Schur updates - extend-add / scatter of the profiled step
        um_eff = timed_prof["factor_sum_ms"] / n_fac - syrk_ms_per_fac - oz_ms_per_fac - full["init_ms"] - full["scatter_ms"]
        kern.insert(2 if oz_ms_per_fac > 0 else 1,
                    {"kernel": "panel factorisation (k_potrf_inv + k_trsm + k_update<128x64 DMMA>)", "bound": "tensor",
                     "algorithmic_flops": upd_flops, "ms": um, "TFLOP/s": upd_flops / (um * 1e-3) / 1e12 if um > 0 else 0.0,
                     "frac": upd_flops / (um * 1e-3) / 1e12 / dmma.value if um > 0 and dmma.value > 0 else None,
                     "share_of_step": um_eff / ms_step, "ms_effective_with_lookahead": um_eff,
                     "TFLOP/s_effective_with_lookahead": upd_flops / (um_eff * 1e-3) / 1e12 if um_eff > 0 else None,
                     "timing": "profiled extra step (events around every launch, look-ahead off: the timed steps overlap "
                               "the panel chain with the trailing updates)"})
        dom_int8 = oz_ms_per_fac > 0 and oz_ms_per_fac >= syrk_ms_per_fac
        kernel_id = f"k_oz_syrk<S={oz_S}>" if dom_int8 else SYRK_KERNEL_ID
        # DRAM traffic comes from a committed ncu --set full capture, not from this run: it is reported only when the
        # capture was taken on the same problem and the same Schur-update kernel (else null until re-captured)
        traffic, traffic_note = None, "no matching ncu --set full capture under profiles/ (captured value, never measured live)"
        for cap in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_syrk_ncu_summary.json")), reverse=True):
            summ = json.load(open(cap))
            cc = summ.get("captured_config") or {}
            if cc.get("dof") == n_dof and cc.get("nrhs") == nr and cc.get("kernel_id") == kernel_id:
                traffic = float(summ["dram_bytes_per_launch"])
                traffic_note = (f"CAPTURED, not live: dram__bytes_read.sum + dram__bytes_write.sum summed over the "
                                f"{summ['launches_per_factorisation']} launches of this kernel in one factorisation, divided by their "
                                f"number (per launch, like `achieved`); {os.path.relpath(cap, ROOT)} (commit {cc.get('commit')}) has every launch")
                break
        common = {"traffic": traffic, "traffic_note": traffic_note,
                  "factor_tflops_all_kernels": float(sum(cls_flops.values())) / (timed_prof["factor_sum_ms"] / n_fac * 1e-3) / 1e12,
                  "factor_ms": timed_prof["factor_sum_ms"] / n_fac,
                  "substitution_ms": timed_prof["solve_sum_ms"] / max(timed_prof["n_solve"], 1.0),
                  "substitutions_per_step": timed_prof["n_solve"] / n_fac,
                  "fp64_tensor_pipe_peak_tflops": dmma.value, "fp64_vector_pipe_peak_tflops": dfma.value,
                  "class_ms_profiled_step": breakdown, "hbm_peak": peak, "kernels": kern}
        if dom_int8:
            n_oz = max(timed_prof["n_oz_fronts"], 1.0)
            roofline = {"bound": "tensor", "kernel": kern[0]["kernel"], "achieved": kern[0]["int8_TOP/s"], "peak": int8_peak,
                        "unit": "TOP/s", "frac": kern[0]["frac"],
                        "peak_source": "2 x MEASURED_PEAKS.json bf16_tflops (" + peaks["_source"] + "): kind::i8 issues 8192 MAC/clk/SM, twice "
                                       "kind::f16; the file holds no int8 figure (nominal dense int8: 4500 TOP/s)",
                        "fp64_equivalent_tflops": kern[0]["fp64_equivalent_TFLOP/s"],
                        "vs_fp64_tensor_pipe_peak": kern[0]["vs_fp64_tensor_pipe_peak"],
                        "algorithmic_ops_per_launch": pairs * cls_flops["syrk_int8"] / (n_oz / n_fac),
                        "ms_per_launch": timed_prof["oz_sum_ms"] / n_oz, "launches_per_step": n_oz / n_fac,
                        "launch": "one front: k_oz_rowmax + k_oz_split + k_oz_syrk_wide, timed together",
                        "int8_mma_issue_peak_TOPs": 4576.0,
                        "frac_of_int8_mma_issue_peak": kern[0]["int8_TOP/s"] / 4576.0,
                        "int8_mma_issue_peak_source": "tools/micro/mma_rate.cu on this pool's B200 (profiles/r2o_mma_rate.txt): kind::i8 M=128 N=256, "
                                                      "operands resident in shared memory, 148 SMs, 128 cycles per MMA",

                        "timing": "CUDA events on the launch stream around the int8 Schur updates of every tree level of the "
                                  f"{int(n_fac)} factorisations inside the timed region", **common}
        else:
            roofline = {"bound": "tensor", "kernel": kern[0]["kernel"], "achieved": tf, "peak": dmma.value, "unit": "TFLOP/s",
                        "frac": kern[0]["frac"],
                        "peak_source": "fp64 tensor pipe (DMMA.8x8x4) measured live by fers_b200_fp64_peak on this device, "
                                       f">=50 ms bursts (DFMA vector pipe: {dfma.value:.2f} TFLOP/s); MEASURED_PEAKS.json holds "
                                       "no fp64 figure",
                        "algorithmic_flops_per_launch": cls_flops["syrk"] / (syrk_launches / n_fac),
                        "ms_per_launch": timed_prof["syrk_sum_ms"] / syrk_launches, "launches_per_step": syrk_launches / n_fac,
                        "timing": "CUDA events on the launch stream around every Schur-update launch of the "
                                  f"{int(n_fac)} factorisations inside the timed region", **common}
        solver_cfg = {"solver": "sparse multifrontal Cholesky (nested dissection; fp64 DMMA tile GEMMs, Schur complements of the "
                                "upper levels as int8 digit-plane products on tcgen05) + "
                                f"residual check, iterative refinement while rel. residual > {engine.CHOL_REFINE_TOL:g} "
                                f"({refine_passes} pass(es) taken), 64 RHS",
                      "fronts": int(len(fr)), "factor_gb": eng._chol.plan.info.factor_doubles * 8 / 1e9,
                      "factor_tflop": float(sum(cls_flops.values())) / 1e12, "solver_rel_residual": relres_solver}
        cache_note = "factor (16 GB) and multi-vectors (354 MB each) >> 126 MB L2; no flush needed"
    else:
        upd_bytes = 6 * V + nn * 36 * 8
        dir_bytes = 3 * V + nn * 36 * 8
        for name, b, ms in (("k_spmm<64,dot>", spmm_bytes, prof[0]), ("k_pcg_update<64>", upd_bytes, prof[1]),
                            ("k_pcg_direction<64>", dir_bytes, prof[2])):
            kern.append(hbm_row(name, b, ms, iters))
        kern.append(hbm_row(asm_name, asm_bytes, am, 1))
        dom = max(kern[:3], key=lambda k: k["share_of_step"])
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["GB/s"], "peak": peak, "unit": "GB/s",
                    "frac": dom["frac"], "traffic": None, "peak_source": peaks["_source"] + " (MEASURED_PEAKS.json hbm_gbs)",
                    "algorithmic_bytes_per_launch": dom["algorithmic_bytes"], "ms_per_launch": dom["ms"],
                    "timing": f"CUDA events on the launch stream around {int(prof[3])} sampled PCG iterations inside the timed region",
                    "kernels": kern}
        solver_cfg = {"solver": "block-Jacobi PCG, 64 RHS, tol 1e-12", "pcg_iterations": iters}
        cache_note = "working set (5 multi-vectors x 354 MB + 227 MB matrix) >> 126 MB L2; no flush needed"

    # ---- end-to-end through the public array API with HOST buffers -------------------------------
    e2e = None
    if not args.no_e2e:
        del rec, ku, d_f, f_case
        nm = model.n_members
        h_u = torch.empty((n_dof, nr), dtype=torch.float64).pin_memory()
        n_sup = len(eng.sup_nodes)
        h_r = torch.empty((n_sup, 6, nr), dtype=torch.float64).pin_memory()      # reactions: rows of the supported nodes
        barrier()
        t_e2e = []
        h2d = 0
        for i in range(2):
            del eng
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            eng = engine.Engine(model, dev, args.assembly_variant, args.solver)   # H2D of the model + symbolic pattern / plan
            out = step_with(eng)
            h_u.copy_(out["U"], non_blocking=True); h_r.copy_(eng.reactions_at_supports(out["R"])[1], non_blocking=True)
            torch.cuda.synchronize(dev)
            t_e2e.append(time.perf_counter() - t0)
            h2d = eng.h2d_bytes
            del out
        # one more pass with a synchronisation after every phase: where the end-to-end time goes (diagnostic only); it also
        # downloads the member end forces (1.9 GB at configs[3]), which the timed passes leave on the device for whoever asks
        del eng
        torch.cuda.synchronize(dev)
        marks = [time.perf_counter()]

        def mark():
            torch.cuda.synchronize(dev)
            marks.append(time.perf_counter())
        eng = engine.Engine(model, dev, args.assembly_variant, args.solver); mark()
        out = step_with(eng); mark()
        h_u.copy_(out["U"], non_blocking=True); h_r.copy_(eng.reactions_at_supports(out["R"])[1], non_blocking=True); mark()
        h_f = torch.empty(tuple(out["f_local"].shape), dtype=torch.float64).pin_memory()
        torch.cuda.synchronize(dev)
        t_f0 = time.perf_counter()
        h_f.copy_(out["f_local"], non_blocking=True)
        torch.cuda.synchronize(dev)
        t_f = time.perf_counter() - t_f0
        del out
        phases = {"engine_setup_ms (H2D of the model, pattern, gather map)": 1e3 * (marks[1] - marks[0]),
                  "run_linear_ms (incl. factorisation plan on the host)": 1e3 * (marks[2] - marks[1]),
                  "of which chol_plan_ms (host)": eng.timings.get("chol_plan_ms"),
                  "d2h_ms (U, R)": 1e3 * (marks[3] - marks[2]),
                  "optional_d2h_of_member_end_forces_ms (not in the timed passes)": 1e3 * t_f,
                  "optional_d2h_of_member_end_forces_bytes": int(h_f.numel() * 8)}
        del h_f
        te = fdist.max_over_ranks(min(t_e2e), dev)
        e2e = {"value": world * units / te, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(h_u.numel() * 8 + h_r.numel() * 8),
               "ms_per_step": 1e3 * te, "phases_of_a_synchronised_pass": phases,
               "api": "fers_core_b200.engine.Engine(model_arrays)." + ("run_linear_sharded" if strong else "run_linear")
                      + "(factors) + D2H of the displacements U of every combination and of the reactions R at the "
                        "supported nodes (Engine.reactions_at_supports: R is exactly zero elsewhere; member end forces stay on the "
                        "device until asked for: optional_d2h_* in the phases)"}

    # ---- CPU baseline: ONE full pass of the same workload on the host cores (rank 0, N=1 only) ------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_frame
        cf, cpu_model = cpu_pass(args)
        t0 = time.perf_counter()
        cf.run_linear(cpu_model.combo_factors)
        dt = time.perf_counter() - t0
        cpu = {"value": cpu_model.n_dof * args.combos / dt, "unit": UNIT, "cores": cpu_frame.threads(),
               "host_cores": os.cpu_count() or 1, "kind": "port", "seconds": dt,
               "phase_seconds": {k: round(v, 4) for k, v in cf.timings.items()},
               "factor_gflops": cf.flops / 1e9 / cf.timings["factor_s"], "max_rel_residual": cf.relres,
               "sample": f"the workload itself ({args.nx}x{args.nz}x{args.ny} frame, {cpu_model.n_dof} DOF x {args.combos} combinations), "
                         "one cold pass including the symbolic plan; " + CPU_KIND_NOTE}

    if rank == 0:
        emit(({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, n_comb) if not strong else
                       workload_name(args, n_comb).replace("per GPU", f"in total, split over {world} GPUs (strong scaling)"),
                       "nodes": model.n_nodes, "members": model.n_members, "dof": n_dof, "free_dof": eng.stats.n_free,
                       "nnz_blocks": nnzb, "load_cases": len(model.load_cases), "combinations_per_gpu": n_comb,
                       **solver_cfg, "max_rel_residual": resid, "parallelism": (f"one model, {n_comb} combinations sharded x{world}, replicated factorisation, "
                                       "all_gather of U, R, f_local" if strong else f"combination batches x{world}"),
                       "cache": cache_note},
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu,
        }))
    if world > 1:
        dist.destroy_process_group()


def run_tree(args, rank: int, world: int, local_rank: int) -> None:
    """--scaling tree: ONE model x --combos combinations on 2^k GPUs, the factorisation split by the subtrees of the front
    tree (strong scaling of the headline problem; every rank ends with all results)."""
    import torch
    import torch.distributed as dist
    from fers_core_b200 import _lib, dist as fdist, engine
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING", "false")
    dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    model = pin_model(build_model(args, 0))                 # the same model and combinations on every rank
    factors = model.combo_factors
    n_dof, n_comb = model.n_dof, factors.shape[0]

    def make():
        return engine.Engine(model, dev, args.assembly_variant, "cholesky", tree_parallel=(rank, world))

    def barrier():
        dist.barrier()
        torch.cuda.synchronize(dev)
    eng = make()
    for _ in range(args.warmup):
        rec = eng.run_linear(factors)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = lib.fers_b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        rec = eng.run_linear(factors)
    e1.record()
    barrier()
    launches = lib.fers_b200_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = fdist.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    value = n_dof * n_comb / (ms_step / 1e3)
    sol = eng._chol
    flops = float(sol.plan.info.flops)
    link = dict(sol.link_bytes)
    resid = float(eng.stats.relres)
    dmma, dfma = C.c_double(), C.c_double()
    scratch = torch.empty(8 * 1024 * 160, dtype=torch.uint8, device=dev)
    _lib.check(lib.fers_b200_fp64_peak(50.0, C.byref(dmma), C.byref(dfma), _lib.ptr(scratch), scratch.numel(), eng.stream), "fp64_peak")
    # end to end with host buffers: model upload, pattern, plan + partition, pass, D2H of U and of the reactions at the supports
    nr = rec["U"].shape[1]
    h_u = torch.empty((n_dof, nr), dtype=torch.float64).pin_memory()
    h_r = torch.empty((len(eng.sup_nodes), 6, nr), dtype=torch.float64).pin_memory()
    del rec
    t_e2e = []
    h2d = 0
    for _ in range(2):
        del eng
        barrier()
        t0 = time.perf_counter()
        eng = make()
        out = eng.run_linear(factors)
        if rank == 0:
            h_u.copy_(out["U"], non_blocking=True); h_r.copy_(eng.reactions_at_supports(out["R"])[1], non_blocking=True)
        torch.cuda.synchronize(dev)
        t_e2e.append(time.perf_counter() - t0)
        h2d = eng.h2d_bytes
        del out
    te = fdist.max_over_ranks(min(t_e2e), dev)
    if rank == 0:
        tf = flops / (ms_step * 1e-3) / 1e12
        emit({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(args, n_comb).replace("per GPU", f"in total on {world} GPUs (strong scaling, tree-parallel factorisation)"),
                       "nodes": model.n_nodes, "members": model.n_members, "dof": n_dof, "combinations": n_comb,
                       "solver": "sparse multifrontal Cholesky split over the ranks by the subtrees of the front tree (chol.TreeParallelSolver)",
                       "fronts_kept_rank0": sol.fronts_kept, "max_rel_residual": resid,
                       "parallelism": f"one model on {world} GPUs: rank s factorises the s-th subtree of depth {sol.k}, upper fronts on the leftmost rank "
                                      "below them; Schur complements / boundary right-hand sides up the tree over NVLink (NCCL send / recv), "
                                      "ancestors' solution down, one all_gather of the solution rows; every other stage replicated",
                       "link_bytes_rank0": link,
                       "cache": "factor (16 GB over the ranks) and multi-vectors (354 MB each) >> 126 MB L2; no flush needed"},
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "whole factorisation + substitution (the kernels of the N = 1 line: k_oz_syrk_wide int8, DMMA tile GEMMs)",
                         "achieved": tf, "peak": world * dmma.value, "unit": "TFLOP/s", "frac": tf / (world * dmma.value) if dmma.value > 0 else None,
                         "traffic": None,
                         "peak_source": f"{world} x the fp64 tensor-pipe peak measured live by fers_b200_fp64_peak (fp64-equivalent flops of one "
                                        "factorisation over the time of a whole pass; the int8 digit-plane path runs above the fp64 pipe's rate, "
                                        "the upper fronts of the tree run on one rank each while the others wait)",
                         "algorithmic_flops_per_step": flops},
            "e2e": {"value": n_dof * n_comb / te, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(h_u.numel() * 8 + h_r.numel() * 8), "ms_per_step": 1e3 * te,
                    "api": "fers_core_b200.engine.Engine(model_arrays, tree_parallel=(rank, world)).run_linear(factors) on every rank + D2H of "
                           "U and of the reactions at the supported nodes on rank 0"},
            "cpu_baseline": None,
        })
    dist.destroy_process_group()


CFG5_WORKLOAD = ("configs[4] batch of {n} parametric 3D frames per GPU (2x2 bays x 4 storeys, 45 nodes, 270 DOF each), sway "
                 "imperfection 1/200, second-order P-Delta (consistent K_g), tolerance 1e-8")


def _cfg5_cpu_worker(job):
    """Oracle (numpy / scipy) on frames [i0, i1) of the batch drawn with `seed`; returns seconds of pure analysis."""
    i0, i1, seed, total = job
    from fers_core_b200 import synth
    from oracle import fers_oracle
    params = synth.portal_params(total, seed)[i0:i1]
    t0 = time.perf_counter()
    for p in params:
        fers_oracle.analyze(synth.portal_frame(tolerance=1e-8, **p), detail=False)
    return time.perf_counter() - t0


def cfg5_cpu_sample(frames: int, seed: int, n_sample: int):
    """The CPU path (oracle, one frame at a time as the reference engine would) on the first n_sample frames of the batch,
    spread over every host core.  Returns (DOF*combos/s, seconds, cores, sample description)."""
    import multiprocessing as mp
    from fers_core_b200 import synth
    cores = os.cpu_count() or 1
    n_sample = max(1, min(n_sample, frames))
    chunk = (n_sample + cores - 1) // cores
    jobs = [(i, min(i + chunk, n_sample), seed, frames) for i in range(0, n_sample, chunk)]
    ctx = mp.get_context("spawn")                   # never fork a process that may hold a CUDA context
    with ctx.Pool(min(cores, len(jobs))) as pool:
        pool.map(_cfg5_cpu_worker, [(0, 1, seed, frames)] * min(cores, len(jobs)))       # imports / first-call costs
        t0 = time.perf_counter()
        pool.map(_cfg5_cpu_worker, jobs)
        dt = time.perf_counter() - t0
    sample = (f"the first {n_sample} of the {frames} frames of the batch, oracle/fers_oracle.py (numpy assembly, dense Cholesky, "
              f"P-Delta loop, recovery) one frame at a time on {min(cores, len(jobs))} host processes")
    return n_sample * 6 * synth.PORTAL_NODES / dt, dt, min(cores, len(jobs)), sample


def run_cfg5_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    from fers_core_b200 import synth
    n_sample = int(os.environ.get("FERS_CFG5_CPU_SAMPLE", "256"))
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, cores, sample = cfg5_cpu_sample(args.frames, synth.BATCH_SEED, n_sample)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    emit(({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([d for _, d in vals])), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": CFG5_WORKLOAD.format(n=args.frames), "frames_per_gpu": args.frames, "dof_per_frame": 270,
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "host_cores": os.cpu_count() or 1, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_cfg5(args, rank: int, world: int, local_rank: int) -> None:
    """BASELINE.json configs[4]: batch of parametric 3D frames (45 nodes, 270 DOF each), sway imperfection 1/200,
    second-order P-Delta (consistent K_g, tolerance 1e-8); frames sharded over GPUs, no data-path collective."""
    # CPU baseline first (rank 0, N = 1): before this process touches CUDA
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from fers_core_b200 import synth as _synth
        v, dt, cores, sample = cfg5_cpu_sample(args.frames, _synth.BATCH_SEED, int(os.environ.get("FERS_CFG5_CPU_SAMPLE", "256")))
        cpu = {"value": v, "unit": UNIT, "cores": cores, "host_cores": os.cpu_count() or 1, "kind": "port", "seconds": dt,
               "sample": sample}
    import torch
    import torch.distributed as dist
    from fers_core_b200 import _lib, engine, synth
    from fers_core_b200 import dist as fdist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    big, _ = synth.portal_batch(args.frames, seed=synth.BATCH_SEED + rank)
    big = pin_model(big)
    units = big.n_dof

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident steps: the batch lives on the GPU, one step = one whole P-Delta analysis of every frame -------
    fb = engine.FrameBatch(big, synth.PORTAL_NODES, dev)
    for _ in range(max(args.warmup, 1)):
        rec = fb.solve()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = lib.fers_b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        rec = fb.solve()
    e1.record()
    barrier()
    launches = lib.fers_b200_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = fdist.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    res = fb.results(rec)
    it_frames = fb.frame_iterations
    enq = fb.stats["iterations_enqueued"]
    summ = fdist.gather_summary(rec["U"].view(args.frames, -1).abs().amax(dim=1))       # per-frame max |u| over NVLink

    # ---- kernels on their own (events around dedicated launches) -------------------------------------------------------
    n, nf = fb.n, fb.nf
    st = fb.eng.stream

    def timed(fn, pre=None, reps=5):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        out = []
        for _ in range(reps):
            if pre is not None:
                pre()
            a0.record(); fn(); a1.record()
            torch.cuda.synchronize(dev)
            out.append(a0.elapsed_time(a1))
        return float(np.median(out))

    def expand():
        _lib.check(lib.fers_b200_dense_expand_batched(
            synth.PORTAL_NODES, nf, _lib.ptr(fb.eng.d_rowptr), _lib.ptr(fb.eng.d_colidx), _lib.ptr(fb.eng.d_vals),
            _lib.ptr(fb.eng.d_fixed), _lib.ptr(fb.eng.d_ks), _lib.ptr(fb.a), _lib.ptr(fb.flag), st), "dense_expand_batched")
    chol_ms = timed(lambda: _lib.check(lib.fers_b200_dense_cholesky(n, nf, _lib.ptr(fb.a), _lib.ptr(fb.info_dev), None, st),
                                       "dense_cholesky"), pre=expand)
    expand_ms = timed(expand)
    u_tmp = torch.ones((big.n_dof, 1), dtype=torch.float64, device=dev)
    solve_ms = timed(lambda: _lib.check(lib.fers_b200_dense_solve(n, 1, nf, _lib.ptr(fb.a), _lib.ptr(u_tmp), st), "dense_solve"))
    scratch = torch.empty(8 * 1024 * 160, dtype=torch.uint8, device=dev)
    dmma, dfma = C.c_double(), C.c_double()
    _lib.check(lib.fers_b200_fp64_peak(50.0, C.byref(dmma), C.byref(dfma), _lib.ptr(scratch), scratch.numel(), st), "fp64_peak")
    peaks = measured_peaks()
    chol_flops = nf * (n ** 3 / 3.0 + n * n / 2.0)
    chol_bytes = 2.0 * nf * n * n * 8
    n_fact = enq + 1                                            # factorisations per step (linear start + one per iteration)
    tf = chol_flops / (chol_ms * 1e-3) / 1e12
    kern = [{"kernel": "k_chol_small (one CTA factorises one 270 x 270 frame matrix in shared-memory panels, DFMA)", "bound": "fp64 vector pipe",
             "algorithmic_flops": chol_flops, "algorithmic_bytes": chol_bytes, "ms": chol_ms, "TFLOP/s": tf, "frac": tf / dfma.value,
             "GB/s": chol_bytes / (chol_ms * 1e-3) / 1e9, "frac_of_hbm": chol_bytes / (chol_ms * 1e-3) / 1e9 / float(peaks["hbm_gbs"]),
             "launches_per_step": n_fact, "share_of_step": n_fact * chol_ms / ms_step},
            {"kernel": "k_dense_expand_batched (block-CSR -> dense frame matrices)", "bound": "hbm", "algorithmic_bytes": chol_bytes / 2,
             "ms": expand_ms, "GB/s": chol_bytes / 2 / (expand_ms * 1e-3) / 1e9,
             "frac": chol_bytes / 2 / (expand_ms * 1e-3) / 1e9 / float(peaks["hbm_gbs"]), "launches_per_step": n_fact,
             "share_of_step": n_fact * expand_ms / ms_step},
            {"kernel": "k_solve_small1 (forward + backward substitution, one CTA per frame)", "bound": "hbm", "algorithmic_bytes": chol_bytes / 2,
             "ms": solve_ms, "GB/s": chol_bytes / 2 / (solve_ms * 1e-3) / 1e9,
             "frac": chol_bytes / 2 / (solve_ms * 1e-3) / 1e9 / float(peaks["hbm_gbs"]), "launches_per_step": n_fact,
             "share_of_step": n_fact * solve_ms / ms_step}]
    roofline = {"bound": "tensor", "kernel": kern[0]["kernel"], "achieved": tf, "peak": dfma.value, "unit": "TFLOP/s", "frac": tf / dfma.value,
                "traffic": None, "peak_source": "fp64 vector pipe (DFMA) measured live by fers_b200_fp64_peak on this device (the kernel is "
                "a chain of 270 dependent pivot steps per matrix, latency-bound: neither the fp64 pipe nor HBM is near its limit; "
                f"fp64 tensor pipe {dmma.value:.1f} TFLOP/s); MEASURED_PEAKS.json holds no fp64 figure",
                "algorithmic_flops_per_launch": chol_flops, "ms_per_launch": chol_ms, "launches_per_step": n_fact,
                "timing": "CUDA events around dedicated launches on freshly expanded matrices, median of 5", "kernels": kern}

    # ---- end to end: host arrays in, host arrays out, every step (engine construction, H2D, analysis, D2H) ---------------
    e2e = None
    if not args.no_e2e:
        del fb, rec
        barrier()
        ts = []
        h2d = d2h = 0
        for _ in range(max(args.steps, 2)):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            fb2 = engine.FrameBatch(big, synth.PORTAL_NODES, dev)
            out = fb2.results(fb2.solve())
            torch.cuda.synchronize(dev)
            ts.append(time.perf_counter() - t0)
            h2d = fb2.eng.h2d_bytes
            d2h = sum(getattr(out, f).nbytes for f in ("U", "R", "f_local", "f_global", "u_local"))
            del fb2
        te = fdist.max_over_ranks(float(np.median(ts)), dev)
        e2e = {"value": world * units / te, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * te,
               "api": "fers_core_b200.engine.analyze_frame_batch == FrameBatch(model, nodes_per_frame).solve() -> .results() "
                      "(model upload, pattern, analysis, D2H of U, R, member end forces and displacements)"}
    if rank == 0:
        emit(({
            "metric": METRIC, "value": world * units / (ms_step / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": CFG5_WORKLOAD.format(n=args.frames), "frames_per_gpu": args.frames, "dof_per_frame": 270,
                       "pdelta_iterations_max": int(it_frames.max()), "pdelta_iterations_mean": float(it_frames.mean()),
                       "pdelta_iterations_enqueued": int(enq), "converged": bool(res.converged[0]),
                       "frames_total": int(summ.shape[0]), "parallelism": f"frame batches x{world}",
                       "cache": "597 MB of frame matrices per factorisation >> 126 MB L2; no flush needed"},
            "clocks": clocks, "gpu_launches": int(launches), "frames_x_iterations_per_s": world * float(it_frames.sum()) / (ms_step / 1e3),
            "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu,
        }))
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def emit(obj: dict) -> None:
    """The one JSON line of this run, on the process's ORIGINAL stdout."""
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    # stdout carries exactly one JSON line: everything else that writes to fd 1 (NCCL's version banner, library
    # chatter) is sent to stderr by pointing fd 1 at fd 2 and keeping a private copy of the real stdout
    global _JSON_FD
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.workload == "cfg5":
            run_cfg5_reference(args, rank, world)
        else:
            run_reference(args, rank, world)
        return
    if args.workload == "cfg5":
        run_cfg5(args, rank, world, local_rank)
        return
    if args.scaling == "tree" and world > 1:
        run_tree(args, rank, world, local_rank)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
