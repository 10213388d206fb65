// recover.cu -- result recovery kernels (MemberResult contract
// fers_core/types/pydantic_models.py:1594-1635, conventions SURVEY.md A.3 / A.6).
//
// Layouts (right-hand-side index fastest, so the lanes of a warp write contiguous bytes):
//   u                      [n_dof][nrhs]
//   f_local/f_global/u_local [n_members][12][nrhs]
//   sec                    [n_members][18][nrhs]   (3 stations x 6 components)
//   lmax/lmin/gmax/gmin    [n_members][6][nrhs]
// One warp per member; lanes stride over the right-hand sides.  All kernels are streaming
// (gather 12 displacement rows, write >= 36 rows per member) and HBM-bound.
#include "common.cuh"

namespace fers {

constexpr int REC_THREADS = 128;
constexpr int REC_WARPS = REC_THREADS / 32;

struct RecArgs {
    int n_members, nrhs;
    const double* xyz;
    const int* n1;
    const int* n2;
    const int* m_prop;
    const int* m_type;
    const double* m_roll;
    const double* m_scale;
    const int* m_rel;
    const double* prop;
    const double* rel;
    const double* m_axial_rhs;
    int kg_mode, kg_suppress;
    const double* u;
    const double* peq;
    double* f_local;
    double* f_global;
    double* u_local;
};

// G lanes per member (a power of two): 32 in general; with few right-hand sides and no member releases a warp
// takes 32 / G members so that the lanes are not idle (P-Delta loops recover one or two columns per pass)
template <int G, bool REL>
__global__ void __launch_bounds__(REC_THREADS) k_recover_members(RecArgs A) {
    static_assert(!REL || G == 32, "released members need the whole warp");
    __shared__ double s_k[REL ? REC_WARPS : 1][144];
    const int warp = threadIdx.x >> 5;
    const int lane = (threadIdx.x & 31) % G;
    const int m = (blockIdx.x * REC_WARPS + warp) * (32 / G) + (threadIdx.x & 31) / G;
    if (m >= A.n_members) return;
    const int a = A.n1[m], b = A.n2[m];
    double pa[3] = {A.xyz[3 * a], A.xyz[3 * a + 1], A.xyz[3 * a + 2]};
    double pb[3] = {A.xyz[3 * b], A.xyz[3 * b + 1], A.xyz[3 * b + 2]};
    const int ri = REL ? A.m_rel[m] : -1;
    const double* rec = (ri >= 0) ? A.rel + REL_STRIDE * (size_t)ri : nullptr;
    if (rec) rec_offset_points(rec, pa, pb);
    const Frame f = member_frame(pa, pb, A.m_roll[m]);
    const int mtype = A.m_type[m];
    const double scale = A.m_scale ? A.m_scale[m] : 1.0;
    const BeamCoef c = beam_coef(A.prop + (size_t)A.m_prop[m] * NPROP, f.L, mtype, scale);
    const int kw = REL ? warp : 0;
    // rigid arms of the end offsets (local axes): K_g acts on the displacements of the elastic axis, u_axis = E u_node
    double arm[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const bool armed = REL && rec && rec_has_offsets(rec);
    if (armed) arm_local(rec, f, arm);
    if (ri >= 0) {
        if (lane == 0) {
            beam_local_k(c, s_k[kw]);
            beam_modify(s_k[kw], rec, f, nullptr, 0);
        }
        __syncwarp();
    }
    const int nr = A.nrhs;
    for (int col = lane; col < nr; col += G) {
        double ug[12], ul[12], fl[12], fg[12];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            ug[q] = A.u[((size_t)a * 6 + q) * nr + col];
            ug[6 + q] = A.u[((size_t)b * 6 + q) * nr + col];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) rot_to_local(f, ug + 3 * q, ul + 3 * q);
        if (ri < 0) {
            beam_k_times_u(c, ul, fl);
        } else {
            const double* k = s_k[kw];
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < 12; ++j) s = fma(k[i * 12 + j], ul[j], s);
                fl[i] = s;
            }
        }
        if (A.m_axial_rhs != nullptr && A.kg_mode != 0) {
            const double N = A.m_axial_rhs[(size_t)m * nr + col];
            if (N != 0.0) {
                if (A.kg_suppress == 0 && !armed) {
                    beam_kg_times_u(N, f.L, A.kg_mode == 1, mtype, ul, fl);
                } else if (A.kg_suppress == 0) {
                    double ua[12], g[12];
#pragma unroll
                    for (int q = 0; q < 12; ++q) { ua[q] = ul[q]; g[q] = 0.0; }
                    arm_fwd(arm, ua);
                    beam_kg_times_u(N, f.L, A.kg_mode == 1, mtype, ua, g);
                    arm_bwd(arm, g);
#pragma unroll
                    for (int q = 0; q < 12; ++q) fl[q] += g[q];
                } else {
                    // K_g with suppressed global translations: S (T^T k_g T) S applied to u
                    double us[12], uls[12], g[12], gg[12], gl[12];
#pragma unroll
                    for (int q = 0; q < 12; ++q) us[q] = ug[q];
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax)
                        if (A.kg_suppress & (1 << ax)) { us[ax] = 0.0; us[6 + ax] = 0.0; }
#pragma unroll
                    for (int q = 0; q < 4; ++q) rot_to_local(f, us + 3 * q, uls + 3 * q);
#pragma unroll
                    for (int q = 0; q < 12; ++q) g[q] = 0.0;
                    if (armed) arm_fwd(arm, uls);
                    beam_kg_times_u(N, f.L, A.kg_mode == 1, mtype, uls, g);
                    if (armed) arm_bwd(arm, g);
#pragma unroll
                    for (int q = 0; q < 4; ++q) rot_to_global(f, g + 3 * q, gg + 3 * q);
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax)
                        if (A.kg_suppress & (1 << ax)) { gg[ax] = 0.0; gg[6 + ax] = 0.0; }
#pragma unroll
                    for (int q = 0; q < 4; ++q) rot_to_local(f, gg + 3 * q, gl + 3 * q);
#pragma unroll
                    for (int q = 0; q < 12; ++q) fl[q] += gl[q];
                }
            }
        }
        if (A.peq != nullptr) {
#pragma unroll
            for (int q = 0; q < 12; ++q) fl[q] -= A.peq[((size_t)m * 12 + q) * nr + col];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) rot_to_global(f, fl + 3 * q, fg + 3 * q);
#pragma unroll
        for (int q = 0; q < 12; ++q) {
            const size_t o = ((size_t)m * 12 + q) * nr + col;
            A.f_local[o] = fl[q];
            A.f_global[o] = fg[q];
            A.u_local[o] = ul[q];
        }
    }
}

struct SecArgs {
    int n_members, nrhs, n_cases;
    const double* xyz;
    const int* n1;
    const int* n2;
    const double* m_roll;
    const int* sp_ptr;
    const double* sp_rows;
    const double* factors;
    const double* f_local;
    const double* f_global;
    const int* m_rel;         // member modifier records (end offsets) or nullptr
    const double* rel;
    double* sec;
    double* lmax;
    double* lmin;
    double* gmax;
    double* gmin;
};

__global__ void __launch_bounds__(REC_THREADS) k_recover_sections(SecArgs A) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = blockIdx.x * REC_WARPS + warp;
    if (m >= A.n_members) return;
    const int a = A.n1[m], b = A.n2[m];
    double pa[3] = {A.xyz[3 * a], A.xyz[3 * a + 1], A.xyz[3 * a + 2]};
    double pb[3] = {A.xyz[3 * b], A.xyz[3 * b + 1], A.xyz[3 * b + 2]};
    const double* rec = (A.m_rel && A.m_rel[m] >= 0) ? A.rel + REL_STRIDE * (size_t)A.m_rel[m] : nullptr;
    if (rec) rec_offset_points(rec, pa, pb);
    const Frame f = member_frame(pa, pb, A.m_roll[m]);
    // start of the elastic axis of an offset member: m_axis = m_node - o x f
    double arm[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const bool armed = rec && rec_has_offsets(rec);
    if (armed) arm_local(rec, f, arm);
    const int nr = A.nrhs;
    const double xs[3] = {0.25, 0.5, 0.75};
    for (int col0 = 0; col0 < nr; col0 += 32) {
        const int col = col0 + lane;
        const bool on = col < nr;
        double span[18];
#pragma unroll
        for (int q = 0; q < 18; ++q) span[q] = 0.0;
        // span-load resultants: computed once per (case, member) by lane 0, broadcast, factored per column
        for (int l = 0; l < A.n_cases; ++l) {
            const int* ptr = A.sp_ptr + (size_t)l * (A.n_members + 1);
            const int q0 = ptr[m], q1 = ptr[m + 1];
            if (q0 == q1) continue;                      // warp-uniform
            double res[18];
#pragma unroll
            for (int q = 0; q < 18; ++q) res[q] = 0.0;
            if (lane == 0) {
                for (int q = q0; q < q1; ++q) {
                    const double* row = A.sp_rows + 9 * (size_t)q;
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        if (row[0] == 0.0) span_dist_resultant(f, row + 1, xs[s], 1.0, res + 6 * s);
                        else span_point_resultant(f, row + 1, xs[s], 1.0, res + 6 * s);
                    }
                }
            }
            const double fac = on ? A.factors[(size_t)col * A.n_cases + l] : 0.0;
#pragma unroll
            for (int q = 0; q < 18; ++q) span[q] = fma(fac, __shfl_sync(FULL, res[q], 0), span[q]);
        }
        if (!on) continue;
        double fl[12], fg[12];
#pragma unroll
        for (int q = 0; q < 12; ++q) {
            const size_t o = ((size_t)m * 12 + q) * nr + col;
            fl[q] = A.f_local[o];
            fg[q] = A.f_global[o];
        }
        double mxl[6], mnl[6], mxg[6], mng[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            mxl[q] = fmax(fl[q], fl[6 + q]); mnl[q] = fmin(fl[q], fl[6 + q]);
            mxg[q] = fmax(fg[q], fg[6 + q]); mng[q] = fmin(fg[q], fg[6 + q]);
        }
#pragma unroll
        if (armed) {
            const double fx = fl[0], fy = fl[1], fz = fl[2];
            fl[3] -= arm[1] * fz - arm[2] * fy;
            fl[4] -= arm[2] * fx - arm[0] * fz;
            fl[5] -= arm[0] * fy - arm[1] * fx;
        }
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const double d = xs[s] * f.L;
            double sl[6], sg[6];
            sl[0] = fl[0] + span[6 * s + 0];
            sl[1] = fl[1] + span[6 * s + 1];
            sl[2] = fl[2] + span[6 * s + 2];
            sl[3] = fl[3] + span[6 * s + 3];
            sl[4] = fl[4] + d * fl[2] + span[6 * s + 4];
            sl[5] = fl[5] - d * fl[1] + span[6 * s + 5];
            rot_to_global(f, sl, sg);
            rot_to_global(f, sl + 3, sg + 3);
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                A.sec[((size_t)m * 18 + 6 * s + q) * nr + col] = sl[q];
                mxl[q] = fmax(mxl[q], sl[q]); mnl[q] = fmin(mnl[q], sl[q]);
                mxg[q] = fmax(mxg[q], sg[q]); mng[q] = fmin(mng[q], sg[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const size_t o = ((size_t)m * 6 + q) * nr + col;
            A.lmax[o] = mxl[q]; A.lmin[o] = mnl[q]; A.gmax[o] = mxg[q]; A.gmin[o] = mng[q];
        }
    }
}

__global__ void k_reaction_finish(size_t total, int nrhs, const uint8_t* __restrict__ supported,
                                  const double* __restrict__ f, double* __restrict__ r) {
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t node = t / ((size_t)6 * nrhs);
        r[t] = supported[node] ? (r[t] - f[t]) : 0.0;
    }
}

__global__ void k_member_axial(size_t total, int nrhs, const double* __restrict__ f_local, double* __restrict__ axial) {
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t m = t / nrhs, col = t - m * nrhs;
        axial[t] = f_local[(m * 12 + 6) * nrhs + col];
    }
}

// one right-hand side, members numbered frame by frame: frozen frames of a P-Delta batch keep their axial forces
__global__ void k_member_axial_masked(int n_members, int members_per_frame, const double* __restrict__ f_local,
                                      const uint8_t* __restrict__ active, double* __restrict__ axial) {
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n_members; m += gridDim.x * blockDim.x)
        if (active == nullptr || active[m / members_per_frame]) axial[m] = f_local[(size_t)m * 12 + 6];
}

// plates: one warp per element; lane 0 builds B matrices and the global stiffness in shared memory
__global__ void __launch_bounds__(REC_THREADS) k_recover_plates(int n_plates, int nrhs, const double* __restrict__ xyz,
                                                                const int* __restrict__ pn, const double* __restrict__ pdata,
                                                                const int* __restrict__ pflags, const double* __restrict__ u,
                                                                double* __restrict__ p_res, double* __restrict__ p_nodal,
                                                                double* __restrict__ p_cd) {
    __shared__ double s_k[REC_WARPS][324];
    __shared__ double s_b[REC_WARPS][8][18];
    __shared__ double s_r[REC_WARPS][9];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * REC_WARPS + warp;
    if (e >= n_plates) return;
    const int nd[3] = {pn[3 * e], pn[3 * e + 1], pn[3 * e + 2]};
    const double* pd = pdata + NPDATA * (size_t)e;
    const double t = pd[PD_T];
    const int flags = pflags[e];
    if (lane == 0) {
        double p0[3] = {xyz[3 * nd[0]], xyz[3 * nd[0] + 1], xyz[3 * nd[0] + 2]};
        double p1[3] = {xyz[3 * nd[1]], xyz[3 * nd[1] + 1], xyz[3 * nd[1] + 2]};
        double p2[3] = {xyz[3 * nd[2]], xyz[3 * nd[2] + 1], xyz[3 * nd[2] + 2]};
        PlateGeom g = plate_geom(p0, p1, p2, pd[PD_THETA]);
        double kl[324];
        plate_local_k(g, pd, flags, kl);
        double blk[36];
        for (int bi = 0; bi < 3; ++bi)
            for (int bj = 0; bj < 3; ++bj) {
                plate_rotate_block(kl, g, bi, bj, blk);
                for (int r = 0; r < 6; ++r)
                    for (int q = 0; q < 6; ++q) s_k[warp][(6 * bi + r) * 18 + 6 * bj + q] = blk[r * 6 + q];
            }
        double Bm[3][18], Bb[3][18], Bs[2][18];
        plate_B_theory(g, flags, Bm, Bb, Bs);
        for (int j = 0; j < 18; ++j) {
            for (int i = 0; i < 3; ++i) { s_b[warp][i][j] = Bm[i][j]; s_b[warp][3 + i][j] = Bb[i][j]; }
            s_b[warp][6][j] = Bs[0][j]; s_b[warp][7][j] = Bs[1][j];
        }
        for (int i = 0; i < 9; ++i) s_r[warp][i] = g.r[i / 3][i % 3];
    }
    __syncwarp();
    double Dm[3][3], D[3][3], Gs[2];
    plate_constitutive(pd, flags, Dm, D, Gs);
    const double cm = (flags & 1) ? pd[PD_MODM] * t : 0.0;
    const double cb = (flags & 2) ? -pd[PD_MODB] * t * t * t / 12.0 : 0.0;
    const double cs = (flags & 2) ? pd[PD_MODS] * (5.0 / 6.0) * t : 0.0;
    for (int col = lane; col < nrhs; col += 32) {
        double ug[18], ul[18];
        for (int n = 0; n < 3; ++n)
            for (int q = 0; q < 6; ++q) ug[6 * n + q] = u[((size_t)nd[n] * 6 + q) * nrhs + col];
        for (int b = 0; b < 6; ++b)
            for (int i = 0; i < 3; ++i)
                ul[3 * b + i] = s_r[warp][3 * i] * ug[3 * b] + s_r[warp][3 * i + 1] * ug[3 * b + 1] + s_r[warp][3 * i + 2] * ug[3 * b + 2];
        double st[8];
        {
            double um[18];                                  // strains are those of the mid-surface (plate offset)
            for (int j = 0; j < 18; ++j) um[j] = ul[j];
            plate_offset_fwd(pd[PD_OFFSET], um);
            for (int i = 0; i < 8; ++i) {
                double s = 0.0;
                for (int j = 0; j < 18; ++j) s = fma(s_b[warp][i][j], um[j], s);
                st[i] = s;
            }
        }
        double out[8];
        for (int i = 0; i < 3; ++i) {
            out[i] = cm * (Dm[i][0] * st[0] + Dm[i][1] * st[1] + Dm[i][2] * st[2]);
            out[3 + i] = cb * (D[i][0] * st[3] + D[i][1] * st[4] + D[i][2] * st[5]);
        }
        out[6] = cs * Gs[0] * st[6];
        out[7] = cs * Gs[1] * st[7];
        for (int i = 0; i < 8; ++i) p_res[((size_t)e * 8 + i) * nrhs + col] = out[i];
        for (int i = 0; i < 18; ++i) {
            double s = 0.0;
            for (int j = 0; j < 18; ++j) s = fma(s_k[warp][i * 18 + j], ug[j], s);
            p_nodal[((size_t)e * 18 + i) * nrhs + col] = s;
        }
        for (int q = 0; q < 6; ++q) {
            p_cd[((size_t)e * 12 + q) * nrhs + col] = (ug[q] + ug[6 + q] + ug[12 + q]) / 3.0;
            p_cd[((size_t)e * 12 + 6 + q) * nrhs + col] = (ul[q] + ul[6 + q] + ul[12 + q]) / 3.0;
        }
    }
}

}  // namespace fers

using namespace fers;

extern "C" {

int fers_b200_recover_members(int n_members, int nrhs, const double* xyz, const int* m_n1, const int* m_n2,
                              const int* m_prop, const int* m_type, const double* m_roll, const double* m_scale,
                              const int* m_rel, const double* prop, const double* rel, const double* m_axial_rhs,
                              int kg_mode, int kg_suppress, const double* u, const double* peq, double* f_local,
                              double* f_global, double* u_local, void* stream) {
    if (n_members <= 0 || nrhs <= 0) return 0;
    RecArgs A;
    A.n_members = n_members; A.nrhs = nrhs; A.xyz = xyz; A.n1 = m_n1; A.n2 = m_n2; A.m_prop = m_prop; A.m_type = m_type;
    A.m_roll = m_roll; A.m_scale = m_scale; A.m_rel = m_rel; A.prop = prop; A.rel = rel; A.m_axial_rhs = m_axial_rhs;
    A.kg_mode = kg_mode; A.kg_suppress = kg_suppress; A.u = u; A.peq = peq; A.f_local = f_local; A.f_global = f_global;
    A.u_local = u_local;
    int g = 32;
    if (m_rel == nullptr) {
        g = 1;
        while (g < nrhs && g < 32) g *= 2;
    }
    const int per_cta = REC_WARPS * (32 / g);
    const int grid = (n_members + per_cta - 1) / per_cta;
    cudaStream_t st = as_stream(stream);
    switch (g) {
        case 1: k_recover_members<1, false><<<grid, REC_THREADS, 0, st>>>(A); break;
        case 2: k_recover_members<2, false><<<grid, REC_THREADS, 0, st>>>(A); break;
        case 4: k_recover_members<4, false><<<grid, REC_THREADS, 0, st>>>(A); break;
        case 8: k_recover_members<8, false><<<grid, REC_THREADS, 0, st>>>(A); break;
        case 16: k_recover_members<16, false><<<grid, REC_THREADS, 0, st>>>(A); break;
        default:
            if (m_rel == nullptr) k_recover_members<32, false><<<grid, REC_THREADS, 0, st>>>(A);
            else k_recover_members<32, true><<<grid, REC_THREADS, 0, st>>>(A);
            break;
    }
    FERS_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int fers_b200_recover_sections(int n_members, int nrhs, int n_cases, const double* xyz, const int* m_n1,
                               const int* m_n2, const double* m_roll, const int* sp_ptr, const double* sp_rows,
                               const double* factors, const double* f_local, const double* f_global,
                               const int* m_rel, const double* rel, double* sec,
                               double* lmax, double* lmin, double* gmax, double* gmin, void* stream) {
    if (n_members <= 0 || nrhs <= 0) return 0;
    SecArgs A;
    A.n_members = n_members; A.nrhs = nrhs; A.n_cases = n_cases; A.xyz = xyz; A.n1 = m_n1; A.n2 = m_n2;
    A.m_roll = m_roll; A.sp_ptr = sp_ptr; A.sp_rows = sp_rows; A.factors = factors; A.f_local = f_local;
    A.f_global = f_global; A.m_rel = m_rel; A.rel = rel; A.sec = sec; A.lmax = lmax; A.lmin = lmin; A.gmax = gmax; A.gmin = gmin;
    int grid = (n_members + REC_WARPS - 1) / REC_WARPS;
    k_recover_sections<<<grid, REC_THREADS, 0, as_stream(stream)>>>(A);
    FERS_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int fers_b200_reactions(int n_nodes, int nrhs, const int* rowptr, const int* colidx, const double* vals,
                        const uint8_t* supported, const double* u, const double* f, double* r, void* stream) {
    if (n_nodes <= 0 || nrhs <= 0) return 0;
    int rc = fers_b200_spmm(n_nodes, nrhs, rowptr, colidx, vals, nullptr, nullptr, u, r, stream);
    if (rc) return rc;
    const size_t total = (size_t)n_nodes * 6 * nrhs;
    const int T = 256;
    size_t g = (total + T - 1) / T;
    size_t cap = (size_t)sm_count() * 16;
    k_reaction_finish<<<(int)(g < cap ? g : cap), T, 0, as_stream(stream)>>>(total, nrhs, supported, f, r);
    FERS_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int fers_b200_recover_plates(int n_plates, int nrhs, const double* xyz, const int* p_nodes, const double* p_data,
                             const int* p_flags, const double* u, double* p_res, double* p_nodal, double* p_cd,
                             void* stream) {
    if (n_plates <= 0 || nrhs <= 0) return 0;
    int grid = (n_plates + REC_WARPS - 1) / REC_WARPS;
    k_recover_plates<<<grid, REC_THREADS, 0, as_stream(stream)>>>(n_plates, nrhs, xyz, p_nodes, p_data, p_flags, u, p_res,
                                                                  p_nodal, p_cd);
    FERS_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int fers_b200_member_axial(int n_members, int nrhs, const double* f_local, double* axial, void* stream) {
    const size_t total = (size_t)n_members * nrhs;
    if (total == 0) return 0;
    const int T = 256;
    size_t g = (total + T - 1) / T;
    size_t cap = (size_t)sm_count() * 16;
    k_member_axial<<<(int)(g < cap ? g : cap), T, 0, as_stream(stream)>>>(total, nrhs, f_local, axial);
    FERS_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int fers_b200_member_axial_masked(int n_members, int members_per_frame, const double* f_local, const uint8_t* active,
                                  double* axial, void* stream) {
    FERS_REQUIRE(n_members > 0 && members_per_frame > 0 && f_local && axial, "member_axial_masked: bad arguments");
    const int T = 256;
    const int g = (n_members + T - 1) / T, cap = sm_count() * 16;
    k_member_axial_masked<<<g < cap ? g : cap, T, 0, as_stream(stream)>>>(n_members, members_per_frame, f_local, active, axial);
    FERS_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

}  // extern "C"
