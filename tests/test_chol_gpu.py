"""GPU parity of the sparse multifrontal Cholesky (chol_plan.cu / chol_numeric.cu) through the C ABI."""
import numpy as np
import pytest

from tests import chol_emulator
from tests.test_chol_plan import _pattern, _random_spd_blocks

pytestmark = pytest.mark.gpu


def _solver_for(rowptr, colidx, xyz, leaf, max_nrhs=64, node_restrained=None):
    import torch
    from fers_core_b200 import chol
    plan = chol.CholPlan(rowptr, colidx, xyz, leaf_nodes=leaf, node_restrained=node_restrained)
    return plan, chol.CholSolver(plan, torch.device("cuda", 0), max_nrhs=max_nrhs)


@pytest.mark.parametrize("drop", [False, True])
@pytest.mark.parametrize("shape,leaf,nrhs", [((3, 2, 2), 4, 8), ((4, 4, 3), 6, 64), ((6, 5, 2), 8, 2),
                                             ((9, 8, 5), 48, 16), ((1, 1, 40), 3, 64)])
def test_chol_random_spd_matches_dense(shape, leaf, nrhs, drop):
    import torch
    from fers_core_b200 import synth
    m = synth.grid_frame(*shape, n_combos=1)
    rowptr, colidx = _pattern(m.n_nodes, m.m_n1, m.m_n2)
    rng = np.random.default_rng(11)
    vals = _random_spd_blocks(rowptr, colidx, rng)
    fixed = (m.bc_code.reshape(-1) != 0).astype(np.uint8)
    fixed[rng.integers(0, len(fixed), 7)] = 1
    kspring = np.where(rng.random(len(fixed)) < 0.05, rng.uniform(1.0, 50.0, len(fixed)), 0.0)
    b = rng.standard_normal((6 * m.n_nodes, nrhs))
    diag_slot = np.array([rowptr[i] + np.searchsorted(colidx[rowptr[i]:rowptr[i + 1]], i) for i in range(m.n_nodes)], np.int32)
    restrained = (m.bc_code != 0).all(axis=1) if drop else None      # base nodes left out of the elimination
    plan, sol = _solver_for(rowptr, colidx, m.xyz, leaf, node_restrained=restrained)
    dev = sol.dev
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dt)).to(dev)
    d_fixed = t(fixed, np.uint8)
    info = sol.factorize(t(vals, np.float64), t(diag_slot, np.int32), d_fixed, t(kspring, np.float64))
    assert info == 0
    d_x = torch.empty((6 * m.n_nodes, nrhs), dtype=torch.float64, device=dev)
    sol.solve(t(b, np.float64), d_x, d_fixed)
    x = d_x.cpu().numpy()
    a = chol_emulator.expand_masked(rowptr, colidx, vals, fixed, np.where(fixed != 0, 0.0, kspring))
    want = np.linalg.solve(a, np.where(fixed[:, None] != 0, 0.0, b))
    assert np.abs(x - want).max() <= 1e-10 * np.abs(want).max()
    # accumulate mode: x += K^-1 b
    sol.solve(t(b, np.float64), d_x, d_fixed, accumulate=True)
    assert np.abs(d_x.cpu().numpy() - 2 * want).max() <= 2e-10 * np.abs(want).max()


def test_chol_reports_indefinite_matrix():
    import torch
    from fers_core_b200 import synth
    m = synth.grid_frame(3, 3, 2, n_combos=1)
    rowptr, colidx = _pattern(m.n_nodes, m.m_n1, m.m_n2)
    rng = np.random.default_rng(3)
    vals = _random_spd_blocks(rowptr, colidx, rng)
    diag_slot = np.array([rowptr[i] + np.searchsorted(colidx[rowptr[i]:rowptr[i + 1]], i) for i in range(m.n_nodes)], np.int32)
    vals[diag_slot[5]] -= (np.eye(6) * 1e4).reshape(-1)          # strongly negative pivot
    fixed = np.zeros(6 * m.n_nodes, np.uint8)
    plan, sol = _solver_for(rowptr, colidx, m.xyz, 4)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dt)).to(sol.dev)
    assert sol.factorize(t(vals, np.float64), t(diag_slot, np.int32), t(fixed, np.uint8)) > 0


@pytest.mark.parametrize("shape", [(8, 7, 5), (14, 12, 9), (20, 20, 12)])
def test_chol_frame_stiffness_matches_oracle(shape):
    """Real frame stiffness (element kernels + assembly + BC on the GPU) solved by the sparse
    Cholesky against the oracle's SuperLU solve of the same model, 1e-9 relative."""
    import torch
    from fers_core_b200 import engine, synth
    from oracle import fers_oracle
    model = synth.grid_frame(*shape, n_combos=64 if shape[0] == 20 else 8)     # 20x20x12: 34 398 DOF, 64 columns
    stats = {}
    got = engine.analyze(model, device="cuda:0", detail=False, stats_out=stats, solver="cholesky")
    assert stats["method"] == "cholesky"
    want = fers_oracle.analyze(model)
    for name in ("U", "R", "f_local"):
        a, b = getattr(got, name), getattr(want, name)
        assert np.abs(a - b).max() <= 1e-9 * np.abs(b).max(), name


@pytest.mark.parametrize("shape,leaf,nrhs", [((9, 8, 5), 48, 16), ((1, 1, 40), 3, 64), ((14, 14, 6), 512, 64)])
def test_wavefront_substitution_is_deterministic_and_matches_per_step_kernels(shape, leaf, nrhs):
    """One-launch-per-level wavefront substitution (default) against the per-step kernels (variant bit 3): same
    answer to rounding, and bit-identical from run to run although its CTAs synchronise through flags."""
    import torch
    from fers_core_b200 import _lib, synth
    lib = _lib.load()
    m = synth.grid_frame(*shape, n_combos=1)
    rowptr, colidx = _pattern(m.n_nodes, m.m_n1, m.m_n2)
    rng = np.random.default_rng(5)
    vals = _random_spd_blocks(rowptr, colidx, rng)
    fixed = (m.bc_code.reshape(-1) != 0).astype(np.uint8)
    b = rng.standard_normal((6 * m.n_nodes, nrhs))
    diag_slot = np.array([rowptr[i] + np.searchsorted(colidx[rowptr[i]:rowptr[i + 1]], i) for i in range(m.n_nodes)], np.int32)
    plan, sol = _solver_for(rowptr, colidx, m.xyz, leaf)
    dev = sol.dev
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dt)).to(dev)
    d_fixed, d_b = t(fixed, np.uint8), t(b, np.float64)
    assert sol.factorize(t(vals, np.float64), t(diag_slot, np.int32), d_fixed, None) == 0
    xs = []
    try:
        for variant in (0, 0, 8):
            _lib.check(lib.fers_b200_chol_set_variant(variant), "set_variant")
            d_x = torch.empty((6 * m.n_nodes, nrhs), dtype=torch.float64, device=dev)
            sol.solve(d_b, d_x, d_fixed)
            xs.append(d_x.cpu().numpy())
    finally:
        lib.fers_b200_chol_set_variant(0)
    assert np.array_equal(xs[0], xs[1])
    assert np.abs(xs[0] - xs[2]).max() <= 1e-12 * np.abs(xs[2]).max()


@pytest.mark.parametrize("name", ["portal_2d_combinations", "plate_strip_surface_load", "nafems_LE5_4el", "nafems_LE6_8x8"])
def test_host_pattern_equals_the_device_pattern(name):
    """The plan thread starts from a host-built pattern (engine._plan_worker): it must be the device's, integer for
    integer, including the extra node pairs of warping members and the plate pairs."""
    from fers_core_b200 import chol, engine, wire
    from tests import helpers
    m = wire.parse_document(helpers.golden(name)["doc"])
    eng = engine.Engine(m, "cuda:0", solver="cholesky", plan_ahead=True)
    p1, p2 = eng._pattern_pairs()
    rowptr, colidx = chol.host_pattern(m.n_nodes, p1, p2, m.p_nodes if m.n_plates else None)
    assert np.array_equal(rowptr, eng.d_rowptr.cpu().numpy()) and np.array_equal(colidx, eng.d_colidx.cpu().numpy())
    # the solve picks up the plan the host thread built from that pattern, and agrees with the inline plan bit for bit
    f_case, _ = eng.load_cases(eng.d_xyz0)
    fac = np.ones((1, max(len(m.load_cases), 1)))
    eng.assemble(eng.d_xyz0)
    u_thread = eng.solve(eng.combine(f_case, fac, 8)).cpu().numpy()
    assert "chol_plan_thread_ms" in eng.timings
    ref = engine.Engine(m, "cuda:0", solver="cholesky")
    f_case, _ = ref.load_cases(ref.d_xyz0)
    ref.assemble(ref.d_xyz0)
    u_inline = ref.solve(ref.combine(f_case, fac, 8)).cpu().numpy()
    assert "chol_plan_thread_ms" not in ref.timings
    if m.n_plates:      # plate stiffness and plate pressures are summed with atomics: equal to round-off, not bit for bit
        assert np.abs(u_thread - u_inline).max() <= 1e-11 * np.abs(u_inline).max(), np.abs(u_thread - u_inline).max() / np.abs(u_inline).max()
    else:
        assert np.array_equal(u_thread, u_inline), np.abs(u_thread - u_inline).max() / np.abs(u_inline).max()
