"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): bit-exact block-CSR pattern, DOF numbering and BC masks; fp64
displacements, reactions and member forces within 1e-9 relative (per result column, relative to
the column's largest magnitude)."""
import ctypes as C

import numpy as np
import pytest

from fers_core_b200 import synth, wire
from oracle import fers_oracle
from tests import helpers

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def engine_mod():
    from fers_core_b200 import engine
    return engine


@pytest.mark.parametrize("name", helpers.golden_names())
def test_golden_documents_match_oracle(engine_mod, name):
    case = helpers.golden(name)
    model = wire.parse_document(case["doc"])
    want = fers_oracle.analyze(model)
    got = engine_mod.analyze(wire.parse_document(case["doc"]))
    rtol = RTOL if "rigid" not in name and "hinge" not in name else 1e-7   # 1e6-stiff rigid links: cond ~1e13
    helpers.assert_results_close(got, want, rtol, label=name)
    assert np.array_equal(got.iterations, want.iterations), (got.iterations, want.iterations)


@pytest.mark.parametrize("name", helpers.golden_names())
def test_golden_documents_meet_reference_expectations(engine_mod, name):
    """Same closed-form / known-answer expectations as the oracle, through calculate_from_json."""
    import json
    case = helpers.golden(name)
    out = json.loads(engine_mod.calculate_from_json(json.dumps(case["doc"])))
    helpers.check_expectations(out["results"], case["expect"])


@pytest.mark.parametrize("name", ["space_frame_3d", "plate_strip_surface_load", "truss_spring_support",
                                  "portal_2d_combinations", "nafems_LE6_8x8"])
def test_pattern_numbering_and_masks_bit_exact(engine_mod, name):
    model = wire.parse_document(helpers.golden(name)["doc"])
    eng = engine_mod.Engine(model)
    eng.assemble(eng.d_xyz0)
    rowptr, colidx = fers_oracle.bsr_pattern(model)
    assert np.array_equal(eng.d_rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(eng.d_colidx.cpu().numpy(), colidx)
    asm = fers_oracle.Assembly(model)
    _, free, dead = fers_oracle.constrain(asm.K_elastic(), model)
    fixed = np.ones(model.n_dof, np.uint8); fixed[free] = 0
    free_index = np.where(fixed == 0, np.cumsum(1 - fixed.astype(np.int64)) - 1, -1).astype(np.int32)
    assert np.array_equal(eng.d_fixed.cpu().numpy(), fixed)
    assert np.array_equal(eng.d_free_index.cpu().numpy(), free_index)
    assert eng.stats.n_free == len(free) and eng.stats.n_dead == int(dead.sum())


def _bsr_to_scipy(eng, n_nodes):
    import scipy.sparse as sp
    vals = eng.d_vals.cpu().numpy().reshape(-1, 6, 6)
    return sp.bsr_matrix((vals, eng.d_colidx.cpu().numpy(), eng.d_rowptr.cpu().numpy()), shape=(6 * n_nodes, 6 * n_nodes))


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4])
def test_assembled_values_match_oracle(engine_mod, variant):
    """Every assembly variant (aggregated/plain atomics x TMA-staged/global section table; 4 = gather)."""
    model = synth.grid_frame(5, 4, 3, timoshenko=True)
    eng = engine_mod.Engine(model, assembly_variant=variant)
    eng.assemble(eng.d_xyz0)
    K = _bsr_to_scipy(eng, model.n_nodes).toarray()
    want = fers_oracle.Assembly(model).K_elastic().toarray()
    assert np.abs(K - want).max() <= 1e-13 * np.abs(want).max()


@pytest.mark.parametrize("dims,dup", [((5, 4, 3), False), ((7, 6, 5), True), ((1, 1, 1), False)])
def test_plain_gather_kernel_matches_half_member_kernel(engine_mod, dims, dup):
    """The element-lane kernel for plain models (27 values per half-member, one lane per block entry) writes the K of
    the per-half-member kernel: same closed forms per entry (on a reciprocal length instead of divisions), same summation
    order on the diagonal -> equal to a few ulp, duplicate members between a node pair left to k_gather_dups, trusses
    and rolled members included."""
    import torch
    from fers_core_b200 import _lib
    model = synth.grid_frame(*dims, timoshenko=True)
    rng = np.random.default_rng(5)
    model.m_roll = np.where(rng.random(model.n_members) < 0.3, rng.normal(size=model.n_members), 0.0)
    model.m_type = np.where(rng.random(model.n_members) < 0.2, 1, model.m_type).astype(np.int32)
    if dup:                                     # a second member on some node pairs
        k = model.n_members // 7
        for name in ("m_n1", "m_n2", "m_prop", "m_type", "m_roll", "m_rel"):
            a = getattr(model, name)
            setattr(model, name, np.concatenate([a, a[:k]]))
        model.m_ids = np.arange(1, len(model.m_n1) + 1, dtype=np.int64)
    lib = _lib.load()
    out = []
    try:
        for kind in (1, 0):
            _lib.check(lib.fers_b200_assemble_set_gather_kernel(kind), "set_gather_kernel")
            eng = engine_mod.Engine(model)
            eng.d_vals.fill_(float("nan"))
            eng.assemble(eng.d_xyz0)
            out.append(eng.d_vals.cpu().numpy().copy())
    finally:
        lib.fers_b200_assemble_set_gather_kernel(0)
    assert np.isfinite(out[1]).all()
    assert np.abs(out[0] - out[1]).max() <= 2e-14 * np.abs(out[0]).max()
    K = _bsr_to_scipy(eng, model.n_nodes).toarray()
    want = fers_oracle.Assembly(model).K_elastic().toarray()
    assert np.abs(K - want).max() <= 1e-13 * np.abs(want).max()


def test_gather_assembly_general_paths_match_scatter(engine_mod):
    """Gather kernel against the atomic scatter-add kernel on a model with duplicate members between
    the same node pair, end releases and geometric stiffness (the dense-local-matrix path); the gather
    result is bit-reproducible."""
    import torch
    model = synth.grid_frame(7, 5, 4, timoshenko=True)
    nm = model.n_members
    dup = np.arange(0, nm, 9)                                  # every 9th member doubled (same nodes, other section)
    for name in ("m_n1", "m_n2", "m_type", "m_roll", "m_prop", "m_rel"):
        a = getattr(model, name)
        setattr(model, name, np.concatenate([a, a[dup]]))
    model.m_prop[nm:] = (model.m_prop[nm:] + 1) % max(model.prop.shape[0], 1)
    rng = np.random.default_rng(5)
    model.rel = np.full((3, wire.REL_STRIDE), -1.0)
    model.rel[:, wire.REL_EIYZ] = 0.0
    model.rel[:, 12:] = 0.0
    model.rel[2, 14:20] = (0.2, -0.1, 0.05, 0.0, 0.15, -0.1)   # end offsets on top of a double release
    model.rel[0, wire.REL_EIYZ] = 3.0e5                        # product of inertia on top of a release
    model.rel[1, 12:14] = (0.031, -0.012)      # shear-centre offset on top of a release spring
    model.rel[0, 5] = 0.0; model.rel[1, 11] = 2.5e5; model.rel[2, [4, 10]] = 0.0
    model.m_rel = np.where(rng.random(model.n_members) < 0.2, rng.integers(0, 3, model.n_members), -1).astype(np.int32)
    vals = {}
    for variant in (1, 4):
        eng = engine_mod.Engine(model, assembly_variant=variant)
        axial = torch.from_numpy(rng.normal(size=model.n_members) * 1e5).to(eng.dev) if variant == 1 else axial
        out = []
        for kg_mode, sup in ((0, 0), (1, 0), (2, 2)):
            eng.assemble(eng.d_xyz0, d_axial=axial if kg_mode else None, kg_mode=kg_mode, kg_suppress=sup)
            out.append(eng.d_vals.cpu().numpy().copy())
        vals[variant] = out
        if variant == 4:
            eng.assemble(eng.d_xyz0, d_axial=axial, kg_mode=2, kg_suppress=2)
            assert np.array_equal(eng.d_vals.cpu().numpy(), out[2])
            # same kernel with the per-member cache (frames and coefficients formed once per member)
            from fers_core_b200 import _lib
            P = _lib.ptr
            v2 = torch.zeros_like(eng.d_vals)
            cache = torch.empty(model.n_members * 24, dtype=torch.float64, device=eng.dev)
            _lib.check(eng.lib.fers_b200_assemble_beams_gather(
                model.n_nodes, model.n_members, P(eng.d_xyz0), P(eng.d_n1), P(eng.d_n2), P(eng.d_m_prop), P(eng.d_type),
                P(eng.d_roll), None, P(eng.d_rel_idx), P(eng.d_prop), eng.n_prop, P(eng.d_rel), eng.d_rel.shape[0], P(axial), 2, 2,
                P(eng.d_m_slot), P(eng.d_hm_ptr), P(eng.d_hm_list), eng.n_dup_half_members, P(eng.d_diag_slot), P(cache), P(v2),
                eng.stream), "gather with member cache")
            assert np.abs(v2.cpu().numpy() - out[2]).max() <= 1e-14 * np.abs(out[2]).max()
    for a, b in zip(vals[1], vals[4]):
        assert np.abs(a - b).max() <= 1e-13 * np.abs(a).max()


@pytest.mark.parametrize("nrhs,variant", [(1, 0), (8, 0), (16, 0), (32, 0), (64, 0), (64, 1), (64, 2)])
def test_spmm_matches_scipy(engine_mod, nrhs, variant):
    import torch
    from fers_core_b200 import _lib
    model = synth.grid_frame(6, 5, 4)
    eng = engine_mod.Engine(model)
    _lib.check(eng.lib.fers_b200_set_spmm_variant(variant), "set_spmm_variant")
    eng.assemble(eng.d_xyz0)
    K = _bsr_to_scipy(eng, model.n_nodes).tocsr()
    rng = np.random.default_rng(3)
    x = rng.normal(size=(model.n_dof, nrhs))
    fixed = eng.d_fixed.cpu().numpy().astype(bool)
    d_x = torch.from_numpy(x).to(eng.dev)
    d_y = torch.empty_like(d_x)
    for masked in (False, True):
        _lib.check(eng.lib.fers_b200_spmm(model.n_nodes, nrhs, _lib.ptr(eng.d_rowptr), _lib.ptr(eng.d_colidx),
                                          _lib.ptr(eng.d_vals), _lib.ptr(eng.d_fixed) if masked else None,
                                          _lib.ptr(eng.d_kspring) if masked else None, _lib.ptr(d_x), _lib.ptr(d_y),
                                          eng.stream), "spmm")
        want = K @ x
        if masked:
            want[fixed] = 0.0
        got = d_y.cpu().numpy()
        assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()
    eng.lib.fers_b200_set_spmm_variant(0)


@pytest.mark.parametrize("dims,force_pcg,variant", [((4, 4, 4), False, 0), ((4, 4, 4), True, 0), ((9, 8, 6), False, 0),
                                                    ((12, 12, 10), True, 0), ((12, 12, 10), True, 1), ((12, 12, 10), True, 2)])
def test_synthetic_frame_matches_oracle(engine_mod, monkeypatch, dims, force_pcg, variant):
    """configs[3]-shaped frames at oracle-sized scale, dense and PCG solve paths (both SpMM kernels)."""
    from fers_core_b200 import _lib
    _lib.load().fers_b200_set_spmm_variant(variant)
    n_combos = 60 if variant >= 1 else 11      # 8 cases + 60 combinations -> one 64-wide and one 8-wide pass
    model = synth.grid_frame(*dims, n_combos=n_combos, timoshenko=(dims[0] == 9), solve_loadcases=True)
    want = fers_oracle.analyze(model)
    stats = {}
    got = engine_mod.analyze(model, stats_out=stats, solver="pcg" if force_pcg else "auto")
    _lib.load().fers_b200_set_spmm_variant(0)
    assert stats["method"] == ("pcg" if force_pcg else "dense")
    helpers.assert_results_close(got, want, RTOL, label=str(dims))


def test_midsize_frame_properties(engine_mod):
    """Size-independent properties on a frame too large for the oracle's comfort: global
    equilibrium, true residual, linearity of combinations."""
    import torch
    from fers_core_b200.hostlogic import rhs_plan
    model = synth.grid_frame(24, 24, 12, n_combos=8, solve_loadcases=True)
    got = engine_mod.analyze(model, detail=False)
    plan = rhs_plan(model)
    nlc = len(model.load_cases)
    for col in range(nlc, plan.n_rhs):
        want = np.einsum("l,lnd->nd", plan.factors[col], got.U[:nlc])
        assert np.abs(got.U[col] - want).max() <= 1e-8 * np.abs(want).max()
    asm = fers_oracle.Assembly(model)
    F, _ = fers_oracle.load_vectors(asm)
    for li in range(nlc):
        tot = F[li].reshape(-1, 6)[:, :3].sum(0) + got.R[li][:, :3].sum(0)
        assert np.abs(tot).max() <= 1e-7 * np.abs(F[li]).sum()
    K = asm.K_elastic()
    _, free, _ = fers_oracle.constrain(K, model)
    for li in range(nlc):
        r = (K @ got.U[li].reshape(-1) - F[li])[free]
        assert np.linalg.norm(r) <= 1e-9 * np.linalg.norm(F[li][free])


def test_full_size_frame_properties(engine_mod):
    """BASELINE.json configs[3] at FULL size (60x60x30 storeys, 692 106 DOF, 64 combinations) through the
    array API and the sparse direct solve: size-independent properties only -- true residual of every
    column, global force equilibrium (reactions against applied loads), linearity (a combination equals
    the factor-weighted sum of the load-case solutions), zero displacement on restrained DOFs, and two
    bit-identical stiffness values / displacements when assembly, factorisation and substitution are repeated."""
    import torch
    from fers_core_b200 import _lib
    model = synth.grid_frame(60, 60, 30, n_combos=64)
    eng = engine_mod.Engine(model)
    factors = model.combo_factors
    nlc = factors.shape[1]
    rec = eng.run_linear(factors)
    assert eng.stats.method == "cholesky" and eng.stats.relres <= 1e-11
    U, R = rec["U"], rec["R"]
    assert bool(torch.isfinite(U).all())
    # true residual with the assembled K (SpMM through the C ABI)
    f_case, _ = eng.load_cases(eng.d_xyz0)
    d_f = eng.combine(f_case, factors, 64)
    ku = torch.empty_like(U)
    _lib.check(eng.lib.fers_b200_spmm(model.n_nodes, 64, _lib.ptr(eng.d_rowptr), _lib.ptr(eng.d_colidx), _lib.ptr(eng.d_vals),
                                      _lib.ptr(eng.d_fixed), _lib.ptr(eng.d_ks), _lib.ptr(U), _lib.ptr(ku), eng.stream), "spmm")
    free = (eng.d_fixed == 0).to(torch.float64)[:, None]
    relres = torch.linalg.norm((ku - d_f) * free, dim=0) / torch.linalg.norm(d_f * free, dim=0)
    assert float(relres.max()) <= 1e-11
    assert float((U * (1 - free)).abs().max()) == 0.0
    # equilibrium: sum of reactions + sum of applied loads = 0 for the three force components
    f3 = d_f.view(model.n_nodes, 6, 64)[:, :3].sum(0)
    r3 = R.view(model.n_nodes, 6, 64)[:, :3].sum(0)
    scale = d_f.view(model.n_nodes, 6, 64)[:, :3].abs().sum(0).max(dim=0).values
    assert float(((f3 + r3).abs().max(dim=0).values / scale).max()) <= 1e-9
    # linearity: unit load-case columns (padded to 8) reproduce every combination
    del rec, ku
    cases = eng.run_linear(np.eye(nlc))["U"][:, :nlc]
    want = cases @ torch.from_numpy(factors.T.copy()).to(cases.device)
    err = (U - want).abs().max(dim=0).values / want.abs().max(dim=0).values
    assert float(err.max()) <= 1e-9
    # reproducibility: assembly (gather kernel), factorisation (look-ahead on two streams) and substitution use
    # no atomics -- the same load vectors give bit-identical K and bit-identical displacements.  (The load-vector
    # kernels accumulate with atomics, so a whole pass is reproducible only to round-off.)
    del cases, want
    eng.assemble(eng.d_xyz0)
    vals1 = eng.d_vals.clone()
    u1 = eng.solve(d_f)
    eng.assemble(eng.d_xyz0)                     # forces a new numeric factorisation
    assert torch.equal(eng.d_vals, vals1)
    u2 = eng.solve(d_f)
    assert torch.equal(u1, u2)


@pytest.mark.parametrize("nrhs,n_cases,n_rows", [(64, 8, 1000), (8, 3, 77), (16, 1, 5), (3, 4, 33), (32, 8, 4099)])
def test_combine_matches_numpy(engine_mod, nrhs, n_cases, n_rows):
    """Factor matrix x per-case vectors (LoadCombination, pydantic_models.py:491-505) through the C ABI."""
    import torch
    from fers_core_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(nrhs + n_rows)
    per_case = rng.normal(size=(n_cases, n_rows))
    fac = rng.normal(size=(nrhs, n_cases))
    dev = torch.device("cuda", 0)
    d_in, d_fac = torch.from_numpy(per_case).to(dev), torch.from_numpy(fac).to(dev)
    d_out = torch.empty((n_rows, nrhs), dtype=torch.float64, device=dev)
    _lib.check(lib.fers_b200_combine(n_rows, n_cases, nrhs, _lib.ptr(d_in), _lib.ptr(d_fac), _lib.ptr(d_out),
                                     torch.cuda.current_stream(dev).cuda_stream), "combine")
    want = per_case.T @ fac.T
    assert np.abs(d_out.cpu().numpy() - want).max() <= 1e-14 * np.abs(want).max()


@pytest.mark.parametrize("name", ["space_frame_3d", "portal_pdelta_consistent", "hinge_spring"])
def test_warping_sections_match_oracle(engine_mod, name):
    """Golden frames with a warping constant on every section (7th nodal DOF active on all members, non-collinear
    joints, releases, P-Delta): GPU against the oracle, including warping displacements and bimoments."""
    import copy
    doc = copy.deepcopy(helpers.golden(name)["doc"])
    for sec in doc["model"]["sections"]:
        sec["i_w"] = 40.0 * float(sec["i_y"]) * 1e-2          # same order as a rolled I-section: Iw ~ Iy h^2 / 4
    doc["analysis"]["options"]["include_warping"] = True
    model = wire.parse_document(doc)
    assert len(model.w_member) and model.n_nodes > model.n_real
    want = fers_oracle.analyze(model)
    got = engine_mod.analyze(wire.parse_document(doc))
    if "hinge" not in name:          # the hinge case carries no torsion
        assert np.abs(want.bw).max() > 0 and np.abs(want.U[:, model.n_real:, 0]).max() > 0
    helpers.assert_results_close(got, want, 1e-7 if "hinge" in name else RTOL, label=name + "+warping")


@pytest.mark.parametrize("name", ["plate_strip_surface_load", "nafems_LE1_24x4"])
def test_plate_modifiers_and_plane_strain_match_oracle(engine_mod, name):
    """PlateElement.stiffness_modifiers and plane_state = PlaneStrain on the GPU path against the oracle."""
    import copy
    doc = copy.deepcopy(helpers.golden(name)["doc"])
    for i, pl in enumerate(doc["model"]["plate_elements"]):
        if i % 2 == 0:
            pl["stiffness_modifiers"] = {"membrane": 0.6, "bending": 0.35, "shear": 1.4}
        if i % 3 == 0:
            pl["plane_state"] = "PlaneStrain"
    model = wire.parse_document(doc)
    assert (model.p_flags & 4).any() and (model.p_data[:, 3] != 1.0).any()
    want = fers_oracle.analyze(model)
    base = fers_oracle.analyze(wire.parse_document(helpers.golden(name)["doc"]))
    assert np.abs(want.U - base.U).max() > 1e-3 * np.abs(base.U).max()        # the options do change the answer
    got = engine_mod.analyze(wire.parse_document(doc))
    helpers.assert_results_close(got, want, RTOL, label=name + "+plate options")


@pytest.mark.parametrize("batch,n,nrhs", [(5, 77, 3), (9, 270, 1), (3, 33, 1), (2, 512, 1), (2, 600, 2), (40, 64, 1), (4, 271, 1), (2, 1000, 1), (3, 17, 1)])
def test_dense_cholesky_batched(engine_mod, batch, n, nrhs):
    """Batched dense factor/solve used for batches of small frames (configs[4]): the single-launch small-matrix
    kernel (n <= 512), the CTA-per-matrix blocked substitution (one right-hand side) and the tiled general path."""
    import torch
    from fers_core_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    A = rng.normal(size=(batch, n, n)); A = A @ A.transpose(0, 2, 1) + n * np.eye(n)
    B = rng.normal(size=(batch, n, nrhs))
    dev = torch.device("cuda", 0)
    dA = torch.from_numpy(A.copy()).to(dev); dB = torch.from_numpy(B.copy()).to(dev)
    info_dev = torch.zeros(batch, dtype=torch.int32, device=dev)
    info = (C.c_int * batch)()
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.fers_b200_dense_cholesky(n, batch, _lib.ptr(dA), _lib.ptr(info_dev), info, st), "chol")
    assert list(info) == [0] * batch
    _lib.check(lib.fers_b200_dense_solve(n, nrhs, batch, _lib.ptr(dA), _lib.ptr(dB), st), "solve")
    X = dB.cpu().numpy()
    want = np.linalg.solve(A, B)
    assert np.abs(X - want).max() <= 1e-11 * np.abs(want).max()


def test_dense_cholesky_reports_the_failing_minor(engine_mod):
    import torch
    from fers_core_b200 import _lib
    lib = _lib.load()
    n, batch = 50, 3
    A = np.stack([np.eye(n) * 4.0] * batch)
    A[1, 17, 17] = -1.0                                  # second matrix: leading minor 18 is not positive definite
    dev = torch.device("cuda", 0)
    dA = torch.from_numpy(A.copy()).to(dev)
    info_dev = torch.zeros(batch, dtype=torch.int32, device=dev)
    info = (C.c_int * batch)()
    _lib.check(lib.fers_b200_dense_cholesky(n, batch, _lib.ptr(dA), _lib.ptr(info_dev), info,
                                            torch.cuda.current_stream(dev).cuda_stream), "chol")
    assert list(info) == [0, 18, 0]


def test_frame_batch_pdelta_matches_per_frame_oracle(engine_mod):
    """configs[4] at oracle scale: a batch of parametric frames with sway imperfection and P-Delta,
    solved at once on the GPU (batched dense Cholesky), against the oracle run frame by frame."""
    n_frames = 12
    big, frames = synth.portal_batch(n_frames, tolerance=1e-10)
    stats = {}
    got = engine_mod.analyze_frame_batch(big, synth.PORTAL_NODES, detail=True, stats_out=stats)
    nn, nm = frames[0].n_nodes, frames[0].n_members
    for f, fm in enumerate(frames):
        want = fers_oracle.analyze(fm)
        assert stats["frame_iterations"][f] == want.iterations[0], (f, stats["frame_iterations"][f], want.iterations)
        for name, lo, hi in (("U", f * nn, (f + 1) * nn), ("R", f * nn, (f + 1) * nn),
                             ("f_local", f * nm, (f + 1) * nm), ("f_global", f * nm, (f + 1) * nm),
                             ("sec", f * nm, (f + 1) * nm), ("lmax", f * nm, (f + 1) * nm)):
            a, b = getattr(got, name)[0, lo:hi], getattr(want, name)[0]
            scale = np.abs(getattr(want, "f_local" if name not in ("U", "R") else name)[0]).max()
            assert np.abs(a - b).max() <= 1e-9 * scale, (f, name, np.abs(a - b).max() / scale)
