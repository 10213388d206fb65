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


@pytest.mark.parametrize("name", ["portal_2d_combinations", "plate_strip_surface_load", "hinge_spring"])
def test_host_pattern_equals_the_oracle_pattern(name):
    """fers_b200_host_pattern (the copy of the device pattern the plan thread starts from) against the oracle's
    block pattern: members, plates and the extra pairs of warping members."""
    from fers_core_b200 import chol, wire
    from oracle import fers_oracle
    from tests import helpers
    m = wire.parse_document(helpers.golden(name)["doc"])
    p1, p2 = m.m_n1, m.m_n2
    if len(m.w_member):
        s_, e_ = m.m_n1[m.w_member], m.m_n2[m.w_member]
        a1, a2 = m.w_aux[:, 0], m.w_aux[:, 1]
        p1 = np.concatenate([p1, np.stack([s_, s_, e_, e_, a1], axis=1).reshape(-1)])
        p2 = np.concatenate([p2, np.stack([a1, a2, a1, a2, a2], axis=1).reshape(-1)])
    rowptr, colidx = chol.host_pattern(m.n_nodes, p1, p2, m.p_nodes if m.n_plates else None)
    want_rp, want_ci = fers_oracle.bsr_pattern(m)
    assert np.array_equal(rowptr, want_rp) and np.array_equal(colidx, want_ci)


def test_host_pattern_grid_frame_and_errors():
    from fers_core_b200 import chol, synth, _lib
    m = synth.grid_frame(5, 4, 3, n_combos=1)
    rowptr, colidx = chol.host_pattern(m.n_nodes, m.m_n1, m.m_n2)
    rp, ci = _pattern(m.n_nodes, m.m_n1, m.m_n2)
    assert np.array_equal(rowptr, rp) and np.array_equal(colidx, ci)
    dup = chol.host_pattern(m.n_nodes, np.concatenate([m.m_n1, m.m_n1]), np.concatenate([m.m_n2, m.m_n2]))
    assert np.array_equal(dup[0], rp) and np.array_equal(dup[1], ci)          # parallel members do not add blocks
    with pytest.raises(_lib.FersB200Error):
        chol.host_pattern(3, np.array([0, 7]), np.array([1, 2]))


@pytest.mark.parametrize("world", [2, 4])
def test_tree_partition_splits_the_fronts_between_the_ranks(world):
    """fers_b200_chol_plan_partition: rank s keeps the subtree under the s-th depth-k front, an upper front goes to the
    leftmost rank below it; the kept fronts are disjoint and cover the tree, every subtree is a contiguous range of the
    elimination order that ends with its top front, and every offset of the whole plan survives (what one rank computes
    can be copied to the same place on another)."""
    import ctypes as C
    from fers_core_b200 import _lib, chol, synth
    model = synth.grid_frame(9, 8, 6)
    rowptr, colidx = chol.host_pattern(model.n_nodes, model.m_n1, model.m_n2)
    k = world.bit_length() - 1
    n_top, n_out = (2 << k) - 1, 4 + 4 * world
    kept, infos, fronts_full = [], [], None
    for rank in range(world):
        plan = chol.CholPlan(rowptr, colidx, model.xyz)
        full = plan.export()
        if fronts_full is None:
            fronts_full = full["fronts"].copy()
            nf = plan.info.n_fronts
            lv_full = [set(full["level_fronts"][full["level_ptr"][l]:full["level_ptr"][l + 1]].tolist()) for l in range(plan.info.n_levels)]
        out = (C.c_longlong * n_out)()
        _lib.check(plan.lib.fers_b200_chol_plan_partition(plan.handle, rank, world, out, n_out), "partition")
        part = plan.export()
        assert np.array_equal(part["fronts"], fronts_full)                      # offsets untouched
        lp = part["level_ptr"]
        mine = set()
        for l in range(plan.info.n_levels):
            lv = set(part["level_fronts"][lp[l]:lp[l + 1]].tolist())
            assert lv <= lv_full[l]
            mine |= lv
        assert len(mine) == out[2] and set(np.unique(part["kscatter"]["front"][: out[3]]).tolist()) <= mine
        kept.append(mine)
        infos.append([int(v) for v in out])
        plan.close()
    assert all(i[:2] == infos[0][:2] and i[4:] == infos[0][4:] for i in infos) and infos[0][0] == k
    top = infos[0][4:4 + n_top]
    lo = infos[0][4 + n_top:4 + n_top + world]
    assert set().union(*kept) == set(range(nf)) and sum(len(s) for s in kept) == nf
    for d in range(k + 1):
        for j in range(1 << d):
            f = top[(1 << d) - 1 + j]
            assert f in kept[j << (k - d)]
            if d < k:                                                            # children in the order of the child list
                assert fronts_full[top[(2 << d) - 1 + 2 * j]]["parent"] == f and fronts_full[top[(2 << d) - 1 + 2 * j + 1]]["parent"] == f
    upper = set(top[: world - 1])
    for r in range(world):
        t = top[world - 1 + r]
        hi = int(fronts_full[t]["pad0"] + fronts_full[t]["n1"])
        pos = sorted(p for f in kept[r] if f not in upper
                     for p in range(int(fronts_full[f]["pad0"]), int(fronts_full[f]["pad0"] + fronts_full[f]["n1"])))
        assert pos == list(range(lo[r], hi))
    with pytest.raises(Exception):
        plan = chol.CholPlan(rowptr, colidx, model.xyz)
        _lib.check(plan.lib.fers_b200_chol_plan_partition(plan.handle, 0, 3, (C.c_longlong * 64)(), 64), "partition")
