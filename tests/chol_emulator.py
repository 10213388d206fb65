"""numpy emulation of the multifrontal numeric phase driven by the index arrays of a CholPlan
(test infrastructure: proves the host-side symbolic analysis -- front node lists, inverse maps,
K scatter lists, level order -- independently of the CUDA kernels)."""
import numpy as np


def expand_masked(rowptr, colidx, vals, fixed, kspring=None):
    """Dense K with constrained DOFs replaced by identity rows/columns (what the factorisation sees)."""
    nn = len(rowptr) - 1
    a = np.zeros((6 * nn, 6 * nn))
    for r in range(nn):
        for k in range(rowptr[r], rowptr[r + 1]):
            c = colidx[k]
            a[6 * r:6 * r + 6, 6 * c:6 * c + 6] = vals[k].reshape(6, 6)
    if kspring is not None:
        a[np.diag_indices_from(a)] += kspring
    fx = fixed.astype(bool)
    a[fx, :] = 0.0
    a[:, fx] = 0.0
    a[fx, fx] = 1.0
    return a


class Emulator:
    """The numeric phase front by front: factor(fronts) -> forward(fronts) -> backward(fronts) on a global right-hand-side
    array, so that a caller can run any subset of the tree (the whole of it: factor_solve; a rank's share of it: the
    tree-parallel protocol test)."""

    def __init__(self, plan_export, rowptr, colidx, vals, fixed, kspring=None):
        self.ex, self.rowptr, self.colidx, self.vals, self.kspring = plan_export, rowptr, colidx, vals, kspring
        self.fx = fixed.astype(bool)
        nf = len(plan_export["fronts"])
        self.panels, self.schur, self.dofs = [None] * nf, [None] * nf, [None] * nf
        ex = plan_export
        self.level_of = {}
        for l in range(len(ex["level_ptr"]) - 1):
            for q in range(ex["level_ptr"][l], ex["level_ptr"][l + 1]):
                self.level_of[int(ex["level_fronts"][q])] = l

    def fronts_in_order(self):
        ex = self.ex
        return [int(ex["level_fronts"][q]) for q in range(ex["level_ptr"][-1])]

    def factor(self, front_ids):
        ex, fx, kspring, rowptr, colidx, vals = self.ex, self.fx, self.kspring, self.rowptr, self.colidx, self.vals
        fronts, nodes, child, inv_off, inv = ex["fronts"], ex["nodes"], ex["child"], ex["inv_off"], ex["inv"]
        schur = self.schur
        for i in front_ids:
            l = self.level_of[i]
            f = fronts[i]
            n1, n2 = int(f["n1"]), int(f["n2"])
            nt = n1 + n2
            F = np.zeros((6 * nt, 6 * nt))
            for ce in range(f["child_begin"], f["child_end"]):
                c = child[ce]
                mp = inv[inv_off[ce]:inv_off[ce] + nt]
                sel = np.nonzero(mp >= 0)[0]
                src = mp[sel]
                assert len(sel) == fronts[c]["n2"] and np.all(np.diff(src) > 0), "inverse map is not a monotone bijection"
                rows = (6 * sel[:, None] + np.arange(6)).reshape(-1)
                srows = (6 * src[:, None] + np.arange(6)).reshape(-1)
                assert schur[c] is not None, f"front {i} needs the Schur complement of its child {c}"
                F[np.ix_(rows, rows)] += schur[c][np.ix_(srows, srows)]
                schur[c] = None
            ks = ex["kscatter"][ex["kscatter_ptr"][l]:ex["kscatter_ptr"][l + 1]]
            for e in ks[ks["front"] == i]:
                nd = nodes[f["nodes_off"]:f["nodes_off"] + nt]
                gi, gj = nd[e["bi"]], nd[e["bj"]]
                assert e["bj"] < n1 and e["bi"] >= e["bj"]
                assert colidx[e["k"]] == gj and rowptr[gi] <= e["k"] < rowptr[gi + 1]
                blk = vals[e["k"]].reshape(6, 6).copy()
                fr = fx[6 * gi:6 * gi + 6]; fc = fx[6 * gj:6 * gj + 6]
                blk[fr, :] = 0.0; blk[:, fc] = 0.0
                if gi == gj:
                    if kspring is not None:
                        blk[np.diag_indices(6)] += np.where(fr, 0.0, kspring[6 * gi:6 * gi + 6])
                    blk[fr, fr] = 1.0
                F[6 * e["bi"]:6 * e["bi"] + 6, 6 * e["bj"]:6 * e["bj"] + 6] += blk
                if e["bi"] != e["bj"]:
                    F[6 * e["bj"]:6 * e["bj"] + 6, 6 * e["bi"]:6 * e["bi"] + 6] += blk.T
            k1 = 6 * n1
            L11 = np.linalg.cholesky(F[:k1, :k1])
            L21 = np.linalg.solve(L11, F[:k1, k1:]).T if n2 else np.zeros((0, k1))
            schur[i] = F[k1:, k1:] - L21 @ L21.T
            self.panels[i] = (L11, L21)
            nd = nodes[f["nodes_off"]:f["nodes_off"] + nt]
            self.dofs[i] = (6 * nd[:, None] + np.arange(6)).reshape(-1)

    def start(self, b):
        return np.where(self.fx[:, None], 0.0, b).astype(np.float64)

    def forward(self, x, front_ids):
        for i in front_ids:
            L11, L21 = self.panels[i]
            k1 = L11.shape[0]
            d = self.dofs[i]
            y = np.linalg.solve(L11, x[d[:k1]])
            x[d[:k1]] = y
            if len(d) > k1:
                x[d[k1:]] -= L21 @ y

    def backward(self, x, front_ids):
        for i in reversed(list(front_ids)):
            L11, L21 = self.panels[i]
            k1 = L11.shape[0]
            d = self.dofs[i]
            y = x[d[:k1]]
            if len(d) > k1:
                y = y - L21.T @ x[d[k1:]]
            x[d[:k1]] = np.linalg.solve(L11.T, y)


def factor_solve(plan_export, rowptr, colidx, vals, fixed, b, kspring=None):
    """Multifrontal LL^T + substitution following the plan; returns x (same shape as b [n_dof, nrhs])."""
    em = Emulator(plan_export, rowptr, colidx, vals, fixed, kspring)
    order = em.fronts_in_order()
    em.factor(order)
    x = em.start(b)
    em.forward(x, order)
    em.backward(x, order)
    return x
