// warping.cu -- restrained-warping (Vlasov) torsion: the 7th nodal DOF of the reference's beam element
// (fers_core/types/pydantic_models.py:1366-1369 include_warping, 1384-1387 "14 DOFs: 0-6 start, 7-13 end",
// Section.i_w 1052-1114, NodalSupport.warping_condition 2104-2115; formulation SURVEY.md A.11, pinned by
// NAFEMS LE5 = 107 643 284.3 Pa).
//
// The warping DOF psi = theta_x' of a real node lives in DOF 0 of an auxiliary node (fers_core_b200/wire.py), so
// the 6x6 block machinery (pattern, assembly, solvers, reactions) is unchanged.  Torsion of a warping member:
//   GJ/L [[1,-1],[-1,1]] on (theta_x1, theta_x2)                                 -- St-Venant, linear
//   EIw/L^3 [[12,6L,-12,6L],[6L,4L^2,-6L,2L^2],[-12,-6L,12,-6L],[6L,2L^2,-6L,4L^2]]
//                                     on (theta_x1, psi1, theta_x2, psi2)        -- cubic Hermite
// The theta-theta part 12 EIw / L^3 is folded into the member's torsion coefficient (beam_coef); the kernels
// here add the psi couplings to K and to the recovered end torques, and form the bimoments (nodal force at psi).
#include "common.cuh"

namespace fers {

struct WarpMember {
    double lx[3];
    double c6, c4, c2;     // 6 EIw / L^2, 4 EIw / L, 2 EIw / L (scaled by the member's stiffness multiplier)
    int m, s, e;
};

__device__ __forceinline__ WarpMember warp_member(int w, const int* __restrict__ w_member, const double* __restrict__ xyz,
                                                  const int* __restrict__ n1, const int* __restrict__ n2,
                                                  const int* __restrict__ m_prop, const double* __restrict__ m_scale,
                                                  const double* __restrict__ prop) {
    WarpMember W;
    W.m = w_member[w];
    W.s = n1[W.m]; W.e = n2[W.m];
    double d[3], l2 = 0.0;
    for (int i = 0; i < 3; ++i) { d[i] = xyz[3 * W.e + i] - xyz[3 * W.s + i]; l2 += d[i] * d[i]; }
    const double L = sqrt(l2);
    for (int i = 0; i < 3; ++i) W.lx[i] = d[i] / L;
    const double eiw = prop[(size_t)m_prop[W.m] * NPROP + P_EIW] * (m_scale ? m_scale[W.m] : 1.0);
    W.c6 = 6.0 * eiw / (L * L); W.c4 = 4.0 * eiw / L; W.c2 = 2.0 * eiw / L;
    return W;
}

// w_slot[w][5][4]: block slots (a,a),(a,b),(b,a),(b,b) of the node pairs (s,s'), (s,e'), (e,s'), (e,e'), (s',e')
__global__ void k_assemble_warping(int n_w, const int* __restrict__ w_member, const double* __restrict__ xyz,
                                   const int* __restrict__ n1, const int* __restrict__ n2, const int* __restrict__ m_prop,
                                   const double* __restrict__ m_scale, const double* __restrict__ prop,
                                   const int* __restrict__ w_slot, double* __restrict__ vals) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_w) return;
    const WarpMember W = warp_member(w, w_member, xyz, n1, n2, m_prop, m_scale, prop);
    const int* sl = w_slot + (size_t)w * 20;
    for (int pair = 0; pair < 4; ++pair) {
        const double sign = pair < 2 ? 1.0 : -1.0;           // theta_x1 rows: +6L, theta_x2 rows: -6L
        double* b_ab = vals + (size_t)sl[4 * pair + 1] * 36;   // (real node, aux node): rotation rows x psi column
        double* b_ba = vals + (size_t)sl[4 * pair + 2] * 36;   // transpose block
        for (int i = 0; i < 3; ++i) {
            const double v = sign * W.c6 * W.lx[i];
            atomicAdd(b_ab + (3 + i) * 6, v);
            atomicAdd(b_ba + 3 + i, v);
        }
    }
    atomicAdd(vals + (size_t)sl[16] * 36, W.c4);               // (s', s')
    atomicAdd(vals + (size_t)sl[17] * 36, W.c2);               // (s', e')
    atomicAdd(vals + (size_t)sl[18] * 36, W.c2);               // (e', s')
    atomicAdd(vals + (size_t)sl[19] * 36, W.c4);               // (e', e')
}

// one warp per warping member, lanes over the right-hand sides
__global__ void __launch_bounds__(256)
k_recover_warping(int n_w, int nr, const int* __restrict__ w_member, const int* __restrict__ w_aux,
                  const double* __restrict__ xyz, const int* __restrict__ n1, const int* __restrict__ n2,
                  const int* __restrict__ m_prop, const double* __restrict__ m_scale, const double* __restrict__ prop,
                  const double* __restrict__ u, double* __restrict__ f_local, double* __restrict__ f_global,
                  double* __restrict__ bw) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_w) return;
    const WarpMember W = warp_member(w, w_member, xyz, n1, n2, m_prop, m_scale, prop);
    const size_t a1 = 6 * (size_t)w_aux[2 * w], a2 = 6 * (size_t)w_aux[2 * w + 1];
    for (int r = lane; r < nr; r += 32) {
        double th1 = 0.0, th2 = 0.0;
        for (int i = 0; i < 3; ++i) {
            th1 += W.lx[i] * u[((size_t)6 * W.s + 3 + i) * nr + r];
            th2 += W.lx[i] * u[((size_t)6 * W.e + 3 + i) * nr + r];
        }
        const double psi1 = u[a1 * nr + r], psi2 = u[a2 * nr + r];
        const double d = W.c6 * (psi1 + psi2);
        const size_t fm = (size_t)W.m * 12;
        f_local[(fm + 3) * nr + r] += d;
        f_local[(fm + 9) * nr + r] -= d;
        if (f_global != nullptr) {
            for (int i = 0; i < 3; ++i) {
                f_global[(fm + 3 + i) * nr + r] += d * W.lx[i];
                f_global[(fm + 9 + i) * nr + r] -= d * W.lx[i];
            }
        }
        bw[((size_t)2 * w) * nr + r] = W.c6 * (th1 - th2) + W.c4 * psi1 + W.c2 * psi2;
        bw[((size_t)2 * w + 1) * nr + r] = W.c6 * (th1 - th2) + W.c2 * psi1 + W.c4 * psi2;
    }
}

}  // namespace fers

using namespace fers;

extern "C" {

int fers_b200_assemble_warping(int n_w, const int* w_member, const double* xyz, const int* m_n1, const int* m_n2,
                               const int* m_prop, const double* m_scale, const double* prop, const int* w_slot, double* vals,
                               void* stream) {
    if (n_w <= 0) return 0;
    FERS_REQUIRE(w_member && xyz && m_n1 && m_n2 && m_prop && prop && w_slot && vals, "assemble_warping: bad arguments");
    k_assemble_warping<<<(n_w + 127) / 128, 128, 0, as_stream(stream)>>>(n_w, w_member, xyz, m_n1, m_n2, m_prop, m_scale, prop,
                                                                        w_slot, vals);
    FERS_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int fers_b200_recover_warping(int n_w, int nrhs, const int* w_member, const int* w_aux, const double* xyz, const int* m_n1,
                              const int* m_n2, const int* m_prop, const double* m_scale, const double* prop, const double* u,
                              double* f_local, double* f_global, double* bw, void* stream) {
    if (n_w <= 0) return 0;
    FERS_REQUIRE(w_member && w_aux && xyz && m_n1 && m_n2 && m_prop && prop && u && f_local && bw, "recover_warping: bad arguments");
    const long long threads = (long long)n_w * 32;
    k_recover_warping<<<(unsigned)((threads + 255) / 256), 256, 0, as_stream(stream)>>>(
        n_w, nrhs, w_member, w_aux, xyz, m_n1, m_n2, m_prop, m_scale, prop, u, f_local, f_global, bw);
    FERS_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

}  // extern "C"
