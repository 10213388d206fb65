"""Host-only statistics of the sparse-Cholesky plan (ordering + symbolic) for a synthetic grid frame."""
import ctypes as C
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, ".")
from fers_core_b200 import synth  # noqa: E402

lib = C.CDLL("fers_core_b200/libfers_b200.so")
nx, nz, ny = (int(a) for a in sys.argv[1:4])
leaf = int(sys.argv[4]) if len(sys.argv) > 4 else 48
m = synth.grid_frame(nx, nz, ny, n_combos=1)
nn = m.n_nodes
a = sp.coo_matrix((np.ones(len(m.m_n1)), (m.m_n1, m.m_n2)), shape=(nn, nn))
a = (a + a.T + sp.identity(nn)).tocsr()
a.sort_indices()
rp = a.indptr.astype(np.int32); ci = a.indices.astype(np.int32)
xyz = np.ascontiguousarray(m.xyz, np.float64)
plan = C.c_void_p()
t = time.perf_counter()
rc = lib.fers_b200_chol_plan_create(nn, rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p),
                                    xyz.ctypes.data_as(C.c_void_p), leaf, C.byref(plan))
dt = time.perf_counter() - t
assert rc == 0, rc
out = (C.c_longlong * 12)()
lib.fers_b200_chol_plan_info(plan, out, 12)
names = ["fronts", "levels", "max_front_nodes", "factor_doubles", "schur0", "schur1", "rhs_rows", "index_bytes",
         "Mflop", "nnzL", "kscatter", "root_sep_nodes"]
print(f"nodes {nn} dof {6*nn} plan {dt*1e3:.0f} ms")
for k, v in zip(names, out):
    print(f"  {k:16s} {v}")
print(f"  factor GB {out[3]*8/1e9:.2f}  schur GB {(out[4]+out[5])*8/1e9:.2f}  rhs GB(64) {out[6]*64*8/1e9:.2f}  Tflop {out[8]/1e6:.2f}")
nl = out[1]
lv = (C.c_longlong * (4 * nl))()
lib.fers_b200_chol_plan_levels(plan, lv, nl)
for l in range(nl):
    print(f"  level {l:2d}: fronts {lv[4*l]:6d} max_n1 {lv[4*l+1]:5d} max_n {lv[4*l+2]:5d} Gflop {lv[4*l+3]/1e3:10.1f}")
