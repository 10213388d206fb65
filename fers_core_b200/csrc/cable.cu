// cable.cu -- device side of the elastic catenary cables (MemberType.Cable, fers_core/types/pydantic_models.py:592-598;
// `pretension` 1525-1528, `unstretched_length` 1542-1545; intrinsic self weight 1337-1340).
//
// The catenary equations of the handful of cables of a model are solved on the host (fers_core_b200/cable.py); what they
// produce per Newton step of the outer iteration is, per cable, a symmetric 3x3 tangent G between the translations of its end
// nodes and a correction of the right-hand side.  The kernels here put them where the solver reads them:
//   K[(a,a)] += G, K[(a,b)] -= G, K[(b,a)] -= G, K[(b,b)] += G   on the translational 3x3 corner of the 6x6 node blocks
//   f[6 a + i, col] += c_i,  f[6 b + i, col] += c_{3+i}
// Several cables may meet at a node: atomic adds (a few hundred values; the order of the additions is the only
// non-reproducible part, as in the plate assembly).
#include "common.cuh"

namespace fers {

__global__ void k_assemble_cables(int n_c, const int* __restrict__ c_member, const double* __restrict__ g,
                                  const int* __restrict__ m_slot, double* __restrict__ vals) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = t / 9, e = t - 9 * c;
    if (c >= n_c) return;
    const int4 sl = reinterpret_cast<const int4*>(m_slot)[c_member[c]];
    const int at = (e / 3) * 6 + (e % 3);
    const double v = g[(size_t)c * 9 + e];
    atomicAdd(vals + (size_t)sl.x * 36 + at, v);
    atomicAdd(vals + (size_t)sl.y * 36 + at, -v);
    atomicAdd(vals + (size_t)sl.z * 36 + at, -v);
    atomicAdd(vals + (size_t)sl.w * 36 + at, v);
}

__global__ void k_cable_rhs(int n_c, const int* __restrict__ c_member, const int* __restrict__ n1, const int* __restrict__ n2,
                            const double* __restrict__ corr, double* __restrict__ f, int nr, int col) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = t / 6, e = t - 6 * c;
    if (c >= n_c) return;
    const int m = c_member[c];
    const int node = e < 3 ? n1[m] : n2[m];
    atomicAdd(f + ((size_t)6 * node + (e % 3)) * nr + col, corr[(size_t)c * 6 + e]);
}

}  // namespace fers

using namespace fers;

extern "C" {

int fers_b200_assemble_cables(int n_cables, const int* c_member, const double* g, const int* m_slot, double* vals,
                              void* stream) {
    if (n_cables <= 0) return 0;
    FERS_REQUIRE(c_member && g && m_slot && vals, "assemble_cables: bad arguments");
    k_assemble_cables<<<(n_cables * 9 + 127) / 128, 128, 0, as_stream(stream)>>>(n_cables, c_member, g, m_slot, vals);
    FERS_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int fers_b200_cable_rhs(int n_cables, const int* c_member, const int* m_n1, const int* m_n2, const double* corr, double* f,
                        int nrhs, int col, void* stream) {
    if (n_cables <= 0) return 0;
    FERS_REQUIRE(c_member && m_n1 && m_n2 && corr && f && nrhs > 0 && col >= 0 && col < nrhs, "cable_rhs: bad arguments");
    k_cable_rhs<<<(n_cables * 6 + 127) / 128, 128, 0, as_stream(stream)>>>(n_cables, c_member, m_n1, m_n2, corr, f, nrhs, col);
    FERS_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

}  // extern "C"
