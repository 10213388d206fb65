"""Host-side checks of the sparse-Cholesky plan (ordering + symbolic): no GPU needed."""
import numpy as np
import pytest
import scipy.sparse as sp

from fers_core_b200 import chol, synth
from tests import chol_emulator


def _pattern(n_nodes, n1, n2):
    a = sp.coo_matrix((np.ones(len(n1)), (n1, n2)), shape=(n_nodes, n_nodes))
    a = (a + a.T + sp.identity(n_nodes)).tocsr()
    a.sort_indices()
    return a.indptr.astype(np.int32), a.indices.astype(np.int32)


def _random_spd_blocks(rowptr, colidx, rng):
    """SPD block matrix on the pattern: sum over edges of random PSD 12x12 element matrices + diagonal."""
    nn = len(rowptr) - 1
    nnzb = len(colidx)
    vals = np.zeros((nnzb, 36))
    slot = {}
    for r in range(nn):
        for k in range(rowptr[r], rowptr[r + 1]):
            slot[(r, int(colidx[k]))] = k
    for r in range(nn):
        vals[slot[(r, r)]] += (np.eye(6) * rng.uniform(0.5, 2.0)).reshape(-1)
        for k in range(rowptr[r], rowptr[r + 1]):
            c = int(colidx[k])
            if c <= r:
                continue
            g = rng.standard_normal((12, 12))
            ke = g @ g.T
            vals[slot[(r, r)]] += ke[:6, :6].reshape(-1)
            vals[slot[(r, c)]] += ke[:6, 6:].reshape(-1)
            vals[slot[(c, r)]] += ke[6:, :6].reshape(-1)
            vals[slot[(c, c)]] += ke[6:, 6:].reshape(-1)
    return vals


@pytest.mark.parametrize("shape,leaf", [((3, 2, 2), 4), ((4, 4, 3), 6), ((6, 5, 2), 8), ((1, 1, 6), 2)])
def test_plan_emulated_solve_matches_dense(shape, leaf):
    m = synth.grid_frame(*shape, n_combos=1)
    rowptr, colidx = _pattern(m.n_nodes, m.m_n1, m.m_n2)
    plan = chol.CholPlan(rowptr, colidx, m.xyz, leaf_nodes=leaf)
    perm = plan.perm()
    assert sorted(perm.tolist()) == list(range(m.n_nodes))
    ex = plan.export()
    fr = ex["fronts"]
    assert int(fr["n1"].sum()) == m.n_nodes
    # parent of a front of level l sits in level l + 1
    level_of = np.zeros(len(fr), np.int64)
    for l in range(plan.info.n_levels):
        level_of[ex["level_fronts"][ex["level_ptr"][l]:ex["level_ptr"][l + 1]]] = l
    has_parent = fr["parent"] >= 0
    assert np.all(level_of[fr["parent"][has_parent]] == level_of[has_parent] + 1)
    rng = np.random.default_rng(5)
    vals = _random_spd_blocks(rowptr, colidx, rng)
    fixed = (m.bc_code.reshape(-1) != 0).astype(np.uint8)
    fixed[rng.integers(0, len(fixed), 5)] = 1
    b = rng.standard_normal((6 * m.n_nodes, 3))
    x = chol_emulator.factor_solve(ex, rowptr, colidx, vals, fixed, b)
    a = chol_emulator.expand_masked(rowptr, colidx, vals, fixed)
    want = np.linalg.solve(a, np.where(fixed[:, None] != 0, 0.0, b))
    assert np.abs(x - want).max() <= 1e-10 * np.abs(want).max()


def test_plan_handles_disconnected_and_coincident_nodes():
    # two disconnected chains + a cluster of coincident nodes
    n = 40
    n1 = np.concatenate([np.arange(0, 14), np.arange(15, 29), np.arange(30, 39)])
    n2 = n1 + 1
    xyz = np.zeros((n, 3))
    xyz[:15, 0] = np.arange(15)
    xyz[15:30, 0] = np.arange(15); xyz[15:30, 1] = 50.0
    rowptr, colidx = _pattern(n, n1, n2)
    plan = chol.CholPlan(rowptr, colidx, xyz, leaf_nodes=3)
    ex = plan.export()
    rng = np.random.default_rng(7)
    vals = _random_spd_blocks(rowptr, colidx, rng)
    fixed = np.zeros(6 * n, np.uint8)
    b = rng.standard_normal((6 * n, 2))
    x = chol_emulator.factor_solve(ex, rowptr, colidx, vals, fixed, b)
    want = np.linalg.solve(chol_emulator.expand_masked(rowptr, colidx, vals, fixed), b)
    assert np.abs(x - want).max() <= 1e-10 * np.abs(want).max()


def test_plan_statistics_scale():
    m = synth.grid_frame(12, 12, 8, n_combos=1)
    rowptr, colidx = _pattern(m.n_nodes, m.m_n1, m.m_n2)
    plan = chol.CholPlan(rowptr, colidx, m.xyz)
    lv = plan.levels()
    assert lv[:, 0].sum() == plan.info.n_fronts
    assert plan.info.nnz_l > len(colidx) * 18
    assert plan.info.root_nodes <= 13 * 9


def test_plan_leaves_fully_restrained_nodes_out():
    """Nodes whose six DOFs are restrained by supports are not eliminated: smaller fronts, same solution."""
    m = synth.grid_frame(5, 4, 3, n_combos=1)
    rowptr, colidx = _pattern(m.n_nodes, m.m_n1, m.m_n2)
    restrained = (m.bc_code != 0).all(axis=1)
    assert restrained.sum() == 6 * 5
    full = chol.CholPlan(rowptr, colidx, m.xyz, leaf_nodes=6)
    plan = chol.CholPlan(rowptr, colidx, m.xyz, leaf_nodes=6, node_restrained=restrained)
    perm = plan.perm()
    n_act = m.n_nodes - int(restrained.sum())
    assert sorted(perm[:n_act].tolist()) == np.nonzero(~restrained)[0].tolist() and np.all(perm[n_act:] == -1)
    assert plan.info.flops < full.info.flops and plan.info.nnz_l < full.info.nnz_l
    ex = plan.export()
    assert int(ex["fronts"]["n1"].sum()) == n_act
    assert not np.isin(ex["nodes"], np.nonzero(restrained)[0]).any()
    rng = np.random.default_rng(9)
    vals = _random_spd_blocks(rowptr, colidx, rng)
    fixed = (m.bc_code.reshape(-1) != 0).astype(np.uint8)
    fixed[rng.integers(0, len(fixed), 5)] = 1
    b = rng.standard_normal((6 * m.n_nodes, 3))
    x = chol_emulator.factor_solve(ex, rowptr, colidx, vals, fixed, b)
    want = np.linalg.solve(chol_emulator.expand_masked(rowptr, colidx, vals, fixed), np.where(fixed[:, None] != 0, 0.0, b))
    assert np.abs(x - want).max() <= 1e-10 * np.abs(want).max()
