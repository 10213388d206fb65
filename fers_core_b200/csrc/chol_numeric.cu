// chol_numeric.cu -- numeric phase of the sparse multifrontal Cholesky on the GPU (SURVEY.md 8a row
// A10: "a batch first-order solve that shares one factorization across many load cases",
// fers_core/types/pydantic_models.py:1216-1219).
//
// Fronts of one tree level are independent and processed as a batch (blockIdx.z = front).  Per level:
//   k_front_init    gather the children's Schur complements into the panel [F11;F21] and into F22
//   k_scatter_k     add the (equilibrated, BC-masked) stiffness blocks of the own nodes
//   per 64-column step s:  k_potrf_inv (diagonal tile: Cholesky + inverse)
//                          k_trsm      (rows below: X = B * inv(D)^T, tile GEMM)
//                          k_update    (trailing columns of the panel: C -= X X^T, tile GEMM)
//   k_syrk          F22 -= L21 L21^T  (tile GEMM, K = all own columns)  -> Schur complement for the parent
// The substitution runs the same level schedule with front-local right-hand-side blocks (forward:
// children -> parent, backward: parent -> children); all its products are tile GEMMs with the
// right-hand-side index as the contiguous dimension.  No atomics anywhere: results are deterministic.
#include <stdlib.h>

#include <mutex>
#include <type_traits>
#include <vector>

#include "chol.cuh"
#include "chol_gemm.cuh"
#include "common.cuh"

namespace fers {

void chol_plan_serialise(const CholPlanHost& P, unsigned char* out);

constexpr int LEVEL_SMALL_FRONTS = 8;

struct LevelCtx {
    const FrontDev* fronts;
    const int* level_fronts;   // fronts of this level
    const int* nodes;
    const int* child;
    const long long* inv_off;
    const int* inv;
    double* factor;
    double* schur_cur;         // arena of this level
    const double* schur_child; // arena of the children's level
    // levels with few fronts (the top of the tree, where every CTA start is on a dependent chain or repeats
    // thousands of times): the front records travel in the kernel parameters (constant bank) instead of two
    // dependent global loads at the start of every CTA
    int n_small;
    FrontDev small_fronts[LEVEL_SMALL_FRONTS];
};

__device__ __forceinline__ FrontDev level_front(const LevelCtx& L, int z) {
    if (L.n_small) return L.small_fronts[z];
    return L.fronts[L.level_fronts[z]];
}

// ---- front initialisation: children's Schur complements -> panel + F22 (lower triangle) -----------
// One thread owns two consecutive rows (the even/odd DOF pair of one node: 16-byte loads and stores
// along the contiguous dimension of the column-major fronts) and INIT_COLS columns (whole nodes: two by default,
// so the child index of a column node is looked up once per six columns).
constexpr int INIT_THREADS = 192;
constexpr int INIT_ROWS = 2 * INIT_THREADS;
// Columns per thread and resident CTAs per SM (FERS_INIT_COLS).  The kernel waits on its loads (long-scoreboard stalls; with
// 24 columns = 142 registers = 2 CTAs only 12 % of the warp slots were occupied, profiles/r2P_front_init_ncu.json), so
// occupancy is what it needs: measured per pass of configs[3] (profiles/r2Q_init_cols.txt) 24 columns 21 ms, 18 columns
// 19.0, 12 columns (80 registers, 4 CTAs) 15.1, 6 columns (62 registers, 5 CTAs) 14.6.
static int g_init_cols = getenv("FERS_INIT_COLS") ? atoi(getenv("FERS_INIT_COLS")) : 12;

template <int INIT_COLS, int MINB>
__global__ void __launch_bounds__(INIT_THREADS, MINB)
k_front_init(const __grid_constant__ LevelCtx L) {
    const FrontDev f = level_front(L, blockIdx.z);
    const int nt = f.n1 + f.n2, m = 6 * nt, k1 = 6 * f.n1, m2 = 6 * f.n2;
    const int r0 = blockIdx.x * INIT_ROWS, c0 = blockIdx.y * INIT_COLS;
    if (r0 >= m || c0 >= m || c0 > r0 + INIT_ROWS - 1) return;
    const int r = r0 + 2 * threadIdx.x;                       // rows r, r + 1 (m is even)
    if (r >= m || c0 > r + 1) return;
    const int bi = r / 6, a = r - 6 * bi;
    const int bj0 = c0 / 6;
    double2 v[INIT_COLS];
#pragma unroll
    for (int q = 0; q < INIT_COLS; ++q) v[q] = make_double2(0.0, 0.0);
    for (int ce = f.child_begin; ce < f.child_end; ++ce) {
        const int* mp = L.inv + L.inv_off[ce];
        const int ci = mp[bi];
        if (ci < 0) continue;
        const FrontDev fc = L.fronts[L.child[ce]];
        const size_t ldc = 6 * (size_t)fc.n2;
        const double* s = L.schur_child + fc.schur_off + 6 * ci + a;
#pragma unroll
        for (int qn = 0; qn < INIT_COLS / 6; ++qn) {
            const int bj = bj0 + qn;
            if (bj > bi) break;                                // column node after the row node: upper triangle
            const int cj = mp[bj];
            if (cj < 0) continue;
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                if (6 * bj + b > r + 1) continue;
                const double2 t = *reinterpret_cast<const double2*>(s + (size_t)(6 * cj + b) * ldc);
                v[6 * qn + b].x += t.x; v[6 * qn + b].y += t.y;
            }
        }
    }
    double* panel = L.factor + f.panel_off;
    double* schur = L.schur_cur + f.schur_off;
#pragma unroll
    for (int q = 0; q < INIT_COLS; ++q) {
        const int c = c0 + q;
        if (c > r + 1 || c >= m) continue;
        double2 t = v[q];
        if (c > r) t.x = 0.0;                                  // (r, r + 1) lies above the diagonal: never read
        if (c < k1) *reinterpret_cast<double2*>(panel + (size_t)c * m + r) = t;
        else *reinterpret_cast<double2*>(schur + (size_t)(c - k1) * m2 + (r - k1)) = t;
    }
}

// ---- equilibration: sc[d] = 1 / sqrt(K_dd) (1 for constrained DOFs and non-positive diagonals) ----
__global__ void k_chol_scale(int n_nodes, const double* __restrict__ vals, const int* __restrict__ diag_slot,
                             const uint8_t* __restrict__ fixed, const double* __restrict__ kspring, double* __restrict__ sc) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= 6 * n_nodes) return;
    const int n = d / 6, a = d - 6 * n;
    double k = vals[(size_t)diag_slot[n] * 36 + a * 7];
    if (kspring != nullptr) k += kspring[d];
    sc[d] = (fixed[d] || !(k > 0.0)) ? 1.0 : rsqrt(k);
}

// ---- stiffness blocks -> panel (one warp per block) ----------------------------------------------
__global__ void __launch_bounds__(256)
k_scatter_k(LevelCtx L, const KScatter* __restrict__ ks, long long n_entries, const double* __restrict__ vals,
            const uint8_t* __restrict__ fixed, const double* __restrict__ kspring, const double* __restrict__ sc) {
    const long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= n_entries) return;
    const int lane = threadIdx.x & 31;
    const KScatter e = ks[w];
    const FrontDev f = L.fronts[e.front];
    const int m = 6 * (f.n1 + f.n2);
    const int gi = L.nodes[f.nodes_off + e.bi], gj = L.nodes[f.nodes_off + e.bj];
    double* panel = L.factor + f.panel_off;
    for (int t = lane; t < 36; t += 32) {
        const int a = t / 6, b = t - 6 * a;
        if (e.bi == e.bj && a < b) continue;      // upper triangle is never read
        const int dr = 6 * gi + a, dc = 6 * gj + b;
        double v;
        if (fixed[dr] || fixed[dc]) {
            v = (dr == dc) ? 1.0 : 0.0;
        } else {
            v = vals[(size_t)e.k * 36 + t];
            if (dr == dc && kspring != nullptr) v += kspring[dr];
            v *= sc[dr] * sc[dc];
        }
        panel[(size_t)(6 * e.bj + b) * m + 6 * e.bi + a] += v;
    }
}

// ---- diagonal tile: Cholesky + explicit inverse of the factor --------------------------------------
constexpr int NB = CHOL_NB;
// 64-column steps per outer block of the two-level blocked panel update (FERS_CHOL_OUTER / fers_b200_chol_set_outer)
static int g_chol_outer = getenv("FERS_CHOL_OUTER") ? atoi(getenv("FERS_CHOL_OUTER")) : 4;
// outer blocks per group of the third blocking level (FERS_CHOL_GROUP / fers_b200_chol_set_outer); 1 = two levels only
static int g_chol_group = getenv("FERS_CHOL_GROUP") ? atoi(getenv("FERS_CHOL_GROUP")) : 6;
constexpr int TP = NB + 1;
constexpr size_t POTRF_SMEM = 2 * NB * TP * sizeof(double);

// Blocked in 8-column panels: the 8x8 diagonal block is factorised redundantly in the registers of every
// row thread (no barrier inside a panel), each row below solves its 8 entries against it, then all 256
// threads apply the rank-8 update to the trailing columns.  The inverse is built block row by block
// row, one thread per column with 8 independent accumulators (a thread only re-reads its own column).
__global__ void __launch_bounds__(256)
k_potrf_inv(const __grid_constant__ LevelCtx L, int s, int* __restrict__ info) {
    const FrontDev f = level_front(L, blockIdx.x);
    const int k1 = 6 * f.n1, m = 6 * (f.n1 + f.n2);
    const int c0 = s * NB;
    if (c0 >= k1) return;
    const int nb = min(NB, k1 - c0);
    extern __shared__ __align__(16) double potrf_smem[];
    double* T = potrf_smem;
    double* X = potrf_smem + NB * TP;
    __shared__ int s_bad;
    __shared__ double s_rdiag[NB];
    double* panel = L.factor + f.panel_off;
    const int tid = threadIdx.x;
    if (tid == 0) s_bad = 0;
    for (int t = tid; t < NB * NB; t += 256) {
        const int i = t % NB, j = t / NB;
        double v = (i == j) ? 1.0 : 0.0;
        if (i < nb && j < nb && i >= j) v = panel[(size_t)(c0 + j) * m + c0 + i];
        T[i * TP + j] = v;
    }
    __syncthreads();
    const int r = tid & 63, tq = tid >> 6;
    for (int jb = 0; jb < NB / 8; ++jb) {
        const int j0 = jb * 8;
        const bool row_thread = (tq == 0 && r >= j0);
        double d[8][8];
        if (row_thread) {
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b <= a; ++b) d[a][b] = T[(j0 + a) * TP + j0 + b];
        }
        __syncthreads();       // every copy of the diagonal block is loaded before its rows are overwritten
        if (row_thread) {
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                double p = d[b][b];
                if (!(p > 0.0)) {
                    if (r == j0 && s_bad == 0) s_bad = j0 + b + 1;
                    p = 1.0;
                }
                const double ri = rsqrt(p);
                d[b][b] = ri;                      // reciprocal of the diagonal entry of L
#pragma unroll
                for (int a = b + 1; a < 8; ++a) d[a][b] *= ri;
#pragma unroll
                for (int c = b + 1; c < 8; ++c)
#pragma unroll
                    for (int a = c; a < 8; ++a) d[a][c] = fma(-d[a][b], d[c][b], d[a][c]);
            }
            if (r < j0 + 8) {
                // a diagonal row: write row (r - j0) of the factored block
#pragma unroll
                for (int a = 0; a < 8; ++a)
                    if (a == r - j0) {
#pragma unroll
                        for (int b = 0; b < a; ++b) T[r * TP + j0 + b] = d[a][b];
                        T[r * TP + j0 + a] = 1.0 / d[a][a];
                    }
            } else {
                double x[8];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    double v = T[r * TP + j0 + b];
#pragma unroll
                    for (int c = 0; c < b; ++c) v = fma(-x[c], d[b][c], v);
                    x[b] = v * d[b][b];
                }
#pragma unroll
                for (int b = 0; b < 8; ++b) T[r * TP + j0 + b] = x[b];
            }
        }
        __syncthreads();
        if (r >= j0 + 8) {
            double lrow[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) lrow[j] = T[r * TP + j0 + j];
            for (int k = j0 + 8 + tq; k <= r; k += 4) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc = fma(lrow[j], T[k * TP + j0 + j], acc);
                T[r * TP + k] -= acc;
            }
        }
        __syncthreads();
    }
    if (tid < NB) s_rdiag[tid] = 1.0 / T[tid * TP + tid];
    __syncthreads();
    if (tid < NB) {
        const int c = tid, cb = c >> 3;
        for (int I = 0; I < NB / 8; ++I) {
            const int i0 = 8 * I;
            if (I < cb) {
#pragma unroll
                for (int a = 0; a < 8; ++a) X[(i0 + a) * TP + c] = 0.0;
                continue;
            }
            double t[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) t[a] = (i0 + a == c) ? 1.0 : 0.0;
            for (int k = 8 * cb; k < i0; ++k) {
                const double xk = X[k * TP + c];
#pragma unroll
                for (int a = 0; a < 8; ++a) t[a] = fma(-T[(i0 + a) * TP + k], xk, t[a]);
            }
            double x[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                double v = t[a];
#pragma unroll
                for (int b = 0; b < a; ++b) v = fma(-T[(i0 + a) * TP + i0 + b], x[b], v);
                x[a] = v * s_rdiag[i0 + a];
            }
#pragma unroll
            for (int a = 0; a < 8; ++a) X[(i0 + a) * TP + c] = x[a];
        }
    }
    __syncthreads();
    double* dinv = L.factor + f.dinv_off + (size_t)s * NB * NB;
    for (int t = tid; t < NB * NB; t += 256) {
        const int i = t % NB, j = t / NB;
        dinv[t] = X[i * TP + j];                                  // column-major NB x NB
        if (i < nb && j < nb && i >= j) panel[(size_t)(c0 + j) * m + c0 + i] = T[i * TP + j];
    }
    if (tid == 0 && s_bad) atomicMax(info, 6 * f.pad0 + c0 + s_bad);   // 1-based elimination-order DOF
}

// ---- tile GEMM kernels of the factorisation --------------------------------------------------------
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#ifdef FERS_CHOL_DFMA     // vector-FMA tile core (kept for A/B measurements)
using G88 = GemmTile<8, 8, false>;
using G86 = GemmTile<8, 8, false>;
using G84 = GemmTile<8, 4, false>;
using G44 = GemmTile<4, 4, false>;
using G48 = GemmTile<4, 8, false>;
using G44N = GemmTile<4, 4, true>;
using G48N = GemmTile<4, 8, true>;
#else                     // fp64 tensor-pipe tile core (default)
using G88 = GemmMma<128, 128, 2, 4, false>;
using G86 = GemmMma<128, 64, 4, 2, false, 16, 3, 2>;    // two CTAs per SM: one tile's epilogue overlaps the other's products
using G84 = GemmMma<128, 64, 2, 4, false>;
using G44 = GemmMma<64, 64, 2, 4, false, 16>;
using G48 = GemmMma<64, 128, 2, 4, false, 16>;
using G44N = GemmMma<64, 64, 2, 4, true, 16>;
using G48N = GemmMma<64, 128, 2, 4, true, 16>;
#endif

// rows below the diagonal tile of step s, inside the current outer block only (rows < min(rend, k1)):
// X = B * inv(D)^T (in place).  The rows below the block are handled once per block by k_row_panel.
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_trsm(const __grid_constant__ LevelCtx L, int s, int rend) {
    extern __shared__ __align__(16) double smem[];
    const FrontDev f = level_front(L, blockIdx.z);
    const int k1 = 6 * f.n1, m = 6 * (f.n1 + f.n2);
    const int c0 = s * NB;
    if (c0 >= k1) return;
    const int nb = min(NB, k1 - c0);
    const int row_end = min(min(rend, k1), m);
    const int r0 = c0 + nb + blockIdx.x * G84::BM;
    if (r0 >= row_end) return;
    double* A = L.factor + f.panel_off + (size_t)c0 * m + r0;
    const double* B = L.factor + f.dinv_off + (size_t)s * NB * NB;
    G84::Acc acc;
    G84::mainloop(acc, smem, A, m, row_end - r0, B, NB, nb, nb);
    G84::store<true>(acc, A, (size_t)m, row_end - r0, nb);
}

// Rows below an outer block [bs, be) (be clipped to the front's own columns): X = B * L_bb^-T by forward
// substitution over the block's 64-column sub-blocks, entirely inside the CTA that owns the 64-row tile:
//   X_j = (B_j - [X_0 .. X_{j-1}] * L_bb[j, 0..j)^T) * inv(D_j)^T
// -- one product with K = 64 j and one with K = 64 per sub-block instead of a trsm launch plus an inner-update
// launch per 64-column step over all the rows of the front.
using GRP = G44;     // tile of the row-panel kernel: 64 rows x 64 columns (128 x 64 measured no faster)
__global__ void __launch_bounds__(GEMM_THREADS, 2)
k_row_panel(const __grid_constant__ LevelCtx L, int bs, int be) {
    extern __shared__ __align__(16) double smem[];
    const FrontDev f = level_front(L, blockIdx.z);
    const int k1 = 6 * f.n1, m = 6 * (f.n1 + f.n2);
    if (bs >= k1) return;
    be = min(be, k1);
    const int r0 = be + blockIdx.x * GRP::BM;
    if (r0 >= m) return;
    const int mrem = m - r0;
    double* panel = L.factor + f.panel_off;
    // the CTA's 64-row strip of the block and the inverted diagonal tiles come from HBM and are touched sub-block by
    // sub-block: ask for all of them now (64 rows = 4 lines per column)
    for (int q = threadIdx.x; q < 4 * (be - bs); q += GEMM_THREADS) {
        const int col = q >> 2, seg = q & 3;
        if (16 * seg < mrem) prefetch_l2(panel + (size_t)(bs + col) * m + r0 + 16 * seg);
    }
    for (int q = threadIdx.x; q < (be - bs) * NB / 16; q += GEMM_THREADS)
        prefetch_l2(L.factor + f.dinv_off + (size_t)(bs / NB) * NB * NB + 16 * (size_t)q);
    for (int cj = bs; cj < be; cj += NB) {
        const int nb = min(NB, be - cj);
        double* C = panel + (size_t)cj * m + r0;
        GRP::Acc acc;
        if (cj > bs) {
            GRP::mainloop(acc, smem, panel + (size_t)bs * m + r0, m, mrem, panel + (size_t)bs * m + cj, m, nb, cj - bs);
            GRP::store<false>(acc, C, (size_t)m, mrem, nb);
            __syncthreads();                    // the tile just written is the operand of the next product
        }
        GRP::mainloop(acc, smem, C, m, mrem, L.factor + f.dinv_off + (size_t)(cj / NB) * NB * NB, NB, nb, nb);
        GRP::store<true>(acc, C, (size_t)m, mrem, nb);
        __syncthreads();
    }
}

// panel update: columns [cbeg, min(cend, k1)) of the panel, rows >= column, get C -= X X^T with
// X = panel columns [kc0, kc0 + kw).  Used twice (two-level blocking): after every 64-column step for
// the columns of the current 256-column block (kw = 64), and once per block for all later columns
// (kw = 256), so that most of the trailing matrix is read and written once per 256 eliminated columns.
template <class G>
__global__ void __launch_bounds__(GEMM_THREADS, G::MIN_CTAS)
k_update(const __grid_constant__ LevelCtx L, int kc0, int kw, int cbeg, int cend, int rend) {
    extern __shared__ __align__(16) double smem[];
    const FrontDev f = level_front(L, blockIdx.z);
    const int k1 = 6 * f.n1, m = 6 * (f.n1 + f.n2);
    if (cend > k1) cend = k1;
    if (cbeg >= cend) return;
    const int row_end = rend >= (1 << 30) ? m : min(min(rend, k1), m);   // inner updates stay inside the outer block
    const int r = cbeg + blockIdx.x * G::BM, c = cbeg + blockIdx.y * G::BN;
    if (r >= row_end || c >= cend || c >= r + G::BM) return;  // outside the panel or entirely above the diagonal
    const double* X = L.factor + f.panel_off + (size_t)kc0 * m;
    double* C = L.factor + f.panel_off + (size_t)c * m + r;
    typename G::Acc acc;
    // K is short here (64 .. 256 columns) and the C tile comes from HBM: ask for it now so that the epilogue's loads
    // find it in L2.  (Seeding the accumulator with -C would hide the latency completely, but it replaces the blocked
    // summation C - (sum of products) by a running one: relative residual of configs[3] 5e-13 -> 3e-12, measured.)
    {
        const int tr = threadIdx.x & 7, tc = threadIdx.x >> 3;               // 8 lines per column, 32 columns per pass
#pragma unroll
        for (int q = 0; q < G::BN / 32; ++q) {
            const int cc = tc + 32 * q;
            if (cc < cend - c && 16 * tr < row_end - r) prefetch_l2(C + (size_t)cc * m + 16 * tr);
        }
    }
    G::mainloop(acc, smem, X + r, m, row_end - r, X + c, m, cend - c, kw);
    G::template store<false>(acc, C, (size_t)m, row_end - r, cend - c);
}

// Schur complement: F22 -= L21 L21^T on the lower tiles
template <class G>
__global__ void __launch_bounds__(GEMM_THREADS, G::MIN_CTAS)
k_syrk(const __grid_constant__ LevelCtx L) {
    extern __shared__ __align__(16) double smem[];
    const FrontDev f = level_front(L, blockIdx.z);
    const int k1 = 6 * f.n1, m2 = 6 * f.n2, m = k1 + m2;
    const int r = blockIdx.x * G::BM, c = blockIdx.y * G::BN;
    if (r >= m2 || c >= m2 || c >= r + G::BM) return;
    const double* L21 = L.factor + f.panel_off + k1;
    typename G::Acc acc;
    G::mainloop(acc, smem, L21 + r, m, m2 - r, L21 + c, m, m2 - c, k1);
    G::template store<false>(acc, L.schur_cur + f.schur_off + (size_t)c * m2 + r, (size_t)m2, m2 - r, m2 - c);
}

// ---- substitution ------------------------------------------------------------------------------------
struct SolveCtx {
    const FrontDev* fronts;
    const int* level_fronts;
    const int* nodes;
    const int* child;
    const long long* inv_off;
    const int* inv;
    const int* pos;            // node -> elimination position
    const double* factor;
    double* W;                 // front-local right-hand sides [rhs_rows][nrhs]
    double* Y;                 // solution in elimination order [n_dof][nrhs]
    int nrhs;
};

// W_f = [ sc * b(own) ; 0 ] + sum over children of their updated boundary rows
__global__ void __launch_bounds__(256)
k_fwd_gather(SolveCtx S, const double* __restrict__ b, const uint8_t* __restrict__ fixed, const double* __restrict__ sc) {
    const FrontDev f = S.fronts[S.level_fronts[blockIdx.y]];
    const int m = 6 * (f.n1 + f.n2);
    const int nr = S.nrhs;
    const long long total = (long long)m * nr;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const int row = (int)(idx / nr), r = (int)(idx - (long long)row * nr);
        const int q = row / 6, a = row - 6 * q;
        double v = 0.0;
        if (q < f.n1) {
            const int g = 6 * S.nodes[f.nodes_off + q] + a;
            v = fixed[g] ? 0.0 : b[(size_t)g * nr + r] * sc[g];
        }
        for (int ce = f.child_begin; ce < f.child_end; ++ce) {
            const int ci = S.inv[S.inv_off[ce] + q];
            if (ci < 0) continue;
            const FrontDev fc = S.fronts[S.child[ce]];
            v += S.W[(size_t)(fc.rhs_off + 6 * (fc.n1 + ci) + a) * nr + r];
        }
        S.W[(size_t)(f.rhs_off + row) * nr + r] = v;
    }
}

// boundary rows of W_f <- solved values of the ancestors
__global__ void __launch_bounds__(256)
k_bwd_gather(SolveCtx S) {
    const FrontDev f = S.fronts[S.level_fronts[blockIdx.y]];
    const int k1 = 6 * f.n1, m2 = 6 * f.n2;
    const int nr = S.nrhs;
    const long long total = (long long)m2 * nr;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const int row = (int)(idx / nr), r = (int)(idx - (long long)row * nr);
        const int q = row / 6, a = row - 6 * q;
        const int node = S.nodes[f.nodes_off + f.n1 + q];
        S.W[(size_t)(f.rhs_off + k1 + row) * nr + r] = S.Y[(size_t)(6 * S.pos[node] + a) * nr + r];
    }
}

// x[node dof] = sc * Y[elimination dof]; constrained DOFs -> 0
__global__ void __launch_bounds__(256)
k_chol_unpermute(int n_nodes, int nr, const int* __restrict__ pos, const uint8_t* __restrict__ fixed,
                 const double* __restrict__ sc, const double* __restrict__ Y, double* __restrict__ x, int accumulate) {
    const long long total = (long long)n_nodes * 6 * nr;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const int d = (int)(idx / nr), r = (int)(idx - (long long)d * nr);
        const int n = d / 6, a = d - 6 * n;
        const int pn = pos[n];
        const double v = (fixed[d] || pn < 0) ? 0.0 : sc[d] * Y[(size_t)(6 * pn + a) * nr + r];
        if (accumulate) x[idx] += v; else x[idx] = v;
    }
}

// forward, diagonal tile of step s:  y_s = inv(D_s) w_s  -> Y
__global__ void __launch_bounds__(GEMM_THREADS, 2)
k_fwd_diag(SolveCtx S, int s) {
    extern __shared__ __align__(16) double smem[];
    const FrontDev f = S.fronts[S.level_fronts[blockIdx.x]];
    const int k1 = 6 * f.n1;
    const int c0 = s * NB;
    if (c0 >= k1) return;
    const int nb = min(NB, k1 - c0), nr = S.nrhs;
    const double* A = S.W + (size_t)(f.rhs_off + c0) * nr;
    const double* B = S.factor + f.dinv_off + (size_t)s * NB * NB;
    G44::Acc acc;
    G44::mainloop(acc, smem, A, nr, nr, B, NB, nb, nb);
    G44::store<true>(acc, S.Y + (size_t)(6 * f.pad0 + c0) * nr, (size_t)nr, nr, nb);
}

// forward, own rows below step s:  w[c0+nb .. k1) -= L[.., s] y_s.  The boundary rows are updated once
// per front by k_fwd_bnd (one product over all own columns) instead of once per step.
__global__ void __launch_bounds__(GEMM_THREADS, 2)
k_fwd_update(SolveCtx S, int s) {
    extern __shared__ __align__(16) double smem[];
    const FrontDev f = S.fronts[S.level_fronts[blockIdx.y]];
    const int k1 = 6 * f.n1, m = 6 * (f.n1 + f.n2);
    const int c0 = s * NB;
    if (c0 >= k1) return;
    const int nb = min(NB, k1 - c0), nr = S.nrhs;
    const int r0 = c0 + nb + blockIdx.x * G48::BN;
    if (r0 >= k1) return;
    const double* A = S.Y + (size_t)(6 * f.pad0 + c0) * nr;
    const double* B = S.factor + f.panel_off + (size_t)c0 * m + r0;
    G48::Acc acc;
    G48::mainloop(acc, smem, A, nr, nr, B, m, k1 - r0, nb);
    G48::store<false>(acc, S.W + (size_t)(f.rhs_off + r0) * nr, (size_t)nr, nr, k1 - r0);
}

// forward, boundary rows:  w_bnd -= L21 y_own   (K = all own columns)
__global__ void __launch_bounds__(GEMM_THREADS, 2)
k_fwd_bnd(SolveCtx S) {
    extern __shared__ __align__(16) double smem[];
    const FrontDev f = S.fronts[S.level_fronts[blockIdx.y]];
    const int k1 = 6 * f.n1, m = 6 * (f.n1 + f.n2);
    const int r0 = k1 + blockIdx.x * G48::BN;
    if (r0 >= m) return;
    const int nr = S.nrhs;
    const double* A = S.Y + (size_t)(6 * f.pad0) * nr;
    const double* B = S.factor + f.panel_off + r0;
    G48::Acc acc;
    G48::mainloop(acc, smem, A, nr, nr, B, m, m - r0, k1);
    G48::store<false>(acc, S.W + (size_t)(f.rhs_off + r0) * nr, (size_t)nr, nr, m - r0);
}

// backward, boundary coupling:  y_own -= L21^T x_bnd
__global__ void __launch_bounds__(GEMM_THREADS, 2)
k_bwd_bnd(SolveCtx S) {
    extern __shared__ __align__(16) double smem[];
    const FrontDev f = S.fronts[S.level_fronts[blockIdx.y]];
    const int k1 = 6 * f.n1, m2 = 6 * f.n2, m = k1 + m2;
    const int c = blockIdx.x * G44N::BN;
    if (c >= k1 || m2 == 0) return;
    const int nr = S.nrhs;
    const double* A = S.W + (size_t)(f.rhs_off + k1) * nr;
    const double* B = S.factor + f.panel_off + (size_t)c * m + k1;
    G44N::Acc acc;
    G44N::mainloop(acc, smem, A, nr, nr, B, m, k1 - c, m2);
    G44N::store<false>(acc, S.Y + (size_t)(6 * f.pad0 + c) * nr, (size_t)nr, nr, k1 - c);
}

// backward, diagonal tile of step s:  x_s = inv(D_s)^T y_s   (in place in Y)
__global__ void __launch_bounds__(GEMM_THREADS, 2)
k_bwd_diag(SolveCtx S, int s) {
    extern __shared__ __align__(16) double smem[];
    const FrontDev f = S.fronts[S.level_fronts[blockIdx.x]];
    const int k1 = 6 * f.n1;
    const int c0 = s * NB;
    if (c0 >= k1) return;
    const int nb = min(NB, k1 - c0), nr = S.nrhs;
    double* A = S.Y + (size_t)(6 * f.pad0 + c0) * nr;
    const double* B = S.factor + f.dinv_off + (size_t)s * NB * NB;
    G44N::Acc acc;
    G44N::mainloop(acc, smem, A, nr, nr, B, NB, nb, nb);
    G44N::store<true>(acc, A, (size_t)nr, nr, nb);
}

// backward, columns before step s:  y_before -= L[s, before]^T x_s
__global__ void __launch_bounds__(GEMM_THREADS, 2)
k_bwd_update(SolveCtx S, int s) {
    extern __shared__ __align__(16) double smem[];
    const FrontDev f = S.fronts[S.level_fronts[blockIdx.y]];
    const int k1 = 6 * f.n1, m = 6 * (f.n1 + f.n2);
    const int c0 = s * NB;
    if (c0 >= k1 || s == 0) return;
    const int nb = min(NB, k1 - c0), nr = S.nrhs;
    const int c = blockIdx.x * G48N::BN;
    if (c >= c0) return;
    const double* A = S.Y + (size_t)(6 * f.pad0 + c0) * nr;
    const double* B = S.factor + f.panel_off + (size_t)c * m + c0;
    G48N::Acc acc;
    G48N::mainloop(acc, smem, A, nr, nr, B, m, c0 - c, nb);
    G48N::store<false>(acc, S.Y + (size_t)(6 * f.pad0 + c) * nr, (size_t)nr, nr, c0 - c);
}

// ---- wavefront substitution through the own columns of every front of a level: ONE launch per level and
// direction instead of two per 64-column step.  A CTA owns one 64-row block of one front; it draws a ticket
// (dispatch order), so the blocks it waits for always hold smaller tickets and are running or done.  Forward
// (left-looking): w_i -= L[i, 0:i) y[0:i) accumulated in registers block column by block column as the
// per-front counter `done` announces them, then y_i = inv(D_i) w_i and done = i + 1.  Backward mirrors it from
// the last block up.  The K order of every accumulation is fixed (block by block), so results do not depend on
// timing.  Cross-CTA data goes through L2 only (st + __threadfence / cp.async.cg).
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// the 64 x 64 tile at `t` (column stride ld doubles) -> L2, one 128-byte line per thread (GEMM_THREADS == 256)
__device__ __forceinline__ void prefetch_tile(const double* t, size_t ld) {
    prefetch_l2(t + (size_t)(threadIdx.x >> 2) * ld + 16 * (threadIdx.x & 3));
}

// ctr[0] = ticket, ctr[1 + front-of-level] = finished blocks of that front
__global__ void __launch_bounds__(GEMM_THREADS, 2)
k_fwd_wave(SolveCtx S, int nf, int* __restrict__ ctr) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_v;
    if (threadIdx.x == 0) s_v = atomicAdd(ctr, 1);
    __syncthreads();
    const int t = s_v;
    const int i = t / nf, fi = t - i * nf;
    const FrontDev f = S.fronts[S.level_fronts[fi]];
    const int k1 = 6 * f.n1, m = 6 * (f.n1 + f.n2);
    const int c0 = i * NB;
    if (c0 >= k1) return;
    const int nb = min(NB, k1 - c0), nr = S.nrhs;
    int* done = ctr + 1 + fi;
    const double* panel = S.factor + f.panel_off;
    double* Yf = S.Y + (size_t)(6 * f.pad0) * nr;
    double* Wi = S.W + (size_t)(f.rhs_off + c0) * nr;
    // the tiles of the last step of the chain (inverse of the diagonal tile, newest block column) come from HBM:
    // fetch them into L2 while waiting
    prefetch_tile(S.factor + f.dinv_off + (size_t)i * NB * NB, NB);
    if (i > 0) prefetch_tile(panel + (size_t)(i - 1) * NB * m + c0, m);
    if (i > 1) prefetch_tile(panel + (size_t)(i - 2) * NB * m + c0, m);
    G44::Acc acc;
    G44::clear(acc);
    int have = 0;
    while (have < i) {
        __syncthreads();
        if (threadIdx.x == 0) {
            int d;
            while ((d = ld_acquire_gpu(done)) <= have) __nanosleep(40);
            s_v = d;
        }
        __syncthreads();
        const int upto = min(s_v, i);
        G44::mainloop<false>(acc, smem, Yf + (size_t)have * NB * nr, nr, nr, panel + (size_t)have * NB * m + c0, m, nb, (upto - have) * NB);
        have = upto;
    }
    if (i > 0) {
        G44::store<false>(acc, Wi, (size_t)nr, nr, nb);
        __threadfence();
        __syncthreads();
    }
    G44::mainloop(acc, smem, Wi, nr, nr, S.factor + f.dinv_off + (size_t)i * NB * NB, NB, nb, nb);
    G44::store<true>(acc, Yf + (size_t)c0 * nr, (size_t)nr, nr, nb);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(done, i + 1);
}

__global__ void __launch_bounds__(GEMM_THREADS, 2)
k_bwd_wave(SolveCtx S, int nf, int* __restrict__ ctr) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_v;
    if (threadIdx.x == 0) s_v = atomicAdd(ctr, 1);
    __syncthreads();
    const int t = s_v;
    const int ib = t / nf, fi = t - ib * nf;            // ib counts blocks from the last one
    const FrontDev f = S.fronts[S.level_fronts[fi]];
    const int k1 = 6 * f.n1, m = 6 * (f.n1 + f.n2);
    const int steps = (k1 + NB - 1) / NB;
    if (ib >= steps) return;
    const int i = steps - 1 - ib, c0 = i * NB;
    const int nb = min(NB, k1 - c0), nr = S.nrhs;
    int* done = ctr + 1 + fi;
    const double* panel = S.factor + f.panel_off;
    double* Yf = S.Y + (size_t)(6 * f.pad0) * nr;
    prefetch_tile(S.factor + f.dinv_off + (size_t)i * NB * NB, NB);
    if (ib > 0) prefetch_tile(panel + (size_t)c0 * m + c0 + NB, m);
    if (ib > 1) prefetch_tile(panel + (size_t)c0 * m + c0 + 2 * NB, m);
    G44N::Acc acc;
    G44N::clear(acc);
    for (int jb = 0; jb < ib; ++jb) {                   // source blocks steps-1, steps-2, .. i+1, always in this order
        if (jb >= s_v || jb == 0) {
            __syncthreads();
            if (threadIdx.x == 0) {
                int d;
                while ((d = ld_acquire_gpu(done)) <= jb) __nanosleep(40);
                s_v = d;
            }
            __syncthreads();
        }
        const int j = steps - 1 - jb, r0 = j * NB;
        G44N::mainloop<false>(acc, smem, Yf + (size_t)r0 * nr, nr, nr, panel + (size_t)c0 * m + r0, m, nb, min(NB, k1 - r0));
    }
    if (ib > 0) {
        G44N::store<false>(acc, Yf + (size_t)c0 * nr, (size_t)nr, nr, nb);
        __threadfence();
        __syncthreads();
    }
    G44N::mainloop(acc, smem, Yf + (size_t)c0 * nr, nr, nr, S.factor + f.dinv_off + (size_t)i * NB * NB, NB, nb, nb);
    G44N::store<true>(acc, Yf + (size_t)c0 * nr, (size_t)nr, nr, nb);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(done, ib + 1);
}

#ifndef FERS_CHOL_DFMA
// Second version of the wavefront kernels (default): the dependent step stays on chip.  The accumulator starts as
// -w_i (loaded before any waiting), so after the left-looking products it holds -(w_i - L y); that tile goes to
// shared memory as the operand of the product with the inverted diagonal tile, which has been resident since the
// kernel started.  No round trip of w_i through L2, one fence per step instead of two.
constexpr size_t wave_max(size_t a, size_t b) { return a > b ? a : b; }
constexpr size_t WAVE_RING = wave_max(wave_max(G44::SMEM, G44N::SMEM), G44::SQ_BYTES);   // cp.async ring, later the operand tile
constexpr size_t WAVE_SMEM = WAVE_RING + G44::SQ_BYTES;

template <bool BACKWARD>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
k_wave2(SolveCtx S, int nf, int* __restrict__ ctr) {
    using G = typename std::conditional<BACKWARD, G44N, G44>::type;
    extern __shared__ __align__(16) double smem[];
    double* dt = smem + WAVE_RING / sizeof(double);                        // inverted diagonal tile
    __shared__ int s_v;
    if (threadIdx.x == 0) s_v = atomicAdd(ctr, 1);
    __syncthreads();
    const int t = s_v;
    const int ib = t / nf, fi = t - ib * nf;            // ib = blocks finished before this one
    const FrontDev f = S.fronts[S.level_fronts[fi]];
    const int k1 = 6 * f.n1, m = 6 * (f.n1 + f.n2);
    const int steps = (k1 + NB - 1) / NB;
    if (ib >= steps) return;
    const int i = BACKWARD ? steps - 1 - ib : ib, c0 = i * NB;
    const int nb = min(NB, k1 - c0), nr = S.nrhs;
    int* done = ctr + 1 + fi;
    const double* panel = S.factor + f.panel_off;
    double* Yf = S.Y + (size_t)(6 * f.pad0) * nr;
    G::load_square(dt, S.factor + f.dinv_off + (size_t)i * NB * NB, NB);
    typename G::Acc acc;
    // forward: the right-hand side rows of the block; backward: the forward solution after the boundary coupling
    G::load_neg(acc, BACKWARD ? Yf + (size_t)c0 * nr : S.W + (size_t)(f.rhs_off + c0) * nr, (size_t)nr, nr, nb);
    if (ib > 0) prefetch_tile(BACKWARD ? panel + (size_t)c0 * m + c0 + NB : panel + (size_t)(i - 1) * NB * m + c0, m);
    int have = 0;
    while (have < ib) {
        __syncthreads();
        if (threadIdx.x == 0) {
            int d;
            while ((d = ld_acquire_gpu(done)) <= have) __nanosleep(40);
            s_v = d;
        }
        __syncthreads();
        const int upto = min(s_v, ib);
        if (!BACKWARD) {
            G::template mainloop<false>(acc, smem, Yf + (size_t)have * NB * nr, nr, nr, panel + (size_t)have * NB * m + c0, m, nb,
                                        (upto - have) * NB);
        } else {
            for (int jb = have; jb < upto; ++jb) {        // fixed order, one source block at a time
                const int r0 = (steps - 1 - jb) * NB;
                G::template mainloop<false>(acc, smem, Yf + (size_t)r0 * nr, nr, nr, panel + (size_t)c0 * m + r0, m, nb, min(NB, k1 - r0));
            }
        }
        have = upto;
    }
    cp_async_wait<0>();
    __syncthreads();                                     // ring idle, inverse tile landed
    G::acc_to_operand(acc, smem);
    __syncthreads();
    typename G::Acc out;
    G::template square_product<BACKWARD>(out, smem, dt);
#pragma unroll
    for (int a = 0; a < G::NT; ++a)
#pragma unroll
        for (int b = 0; b < G::MT; ++b) { out.v[a][b][0] = -out.v[a][b][0]; out.v[a][b][1] = -out.v[a][b][1]; }
    G::template store<true>(out, Yf + (size_t)c0 * nr, (size_t)nr, nr, nb);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(done, ib + 1);
}
#endif

// r = mask (f - Kx); norms2[0][c] += r_c . r_c, norms2[1][c] += (mask f_c) . (mask f_c)  (diagnostics)
__global__ void __launch_bounds__(256)
k_chol_residual(long long n_dof, int nr, const double* __restrict__ f, const double* __restrict__ kx,
                const uint8_t* __restrict__ fixed, double* __restrict__ r, double* __restrict__ norms2) {
    __shared__ double s_r[64], s_f[64];
    if (threadIdx.x < 64) { s_r[threadIdx.x] = 0.0; s_f[threadIdx.x] = 0.0; }
    __syncthreads();
    // thread -> fixed column (256 is a multiple of every supported nr), rows strided
    const int c = threadIdx.x % nr;
    const int rows_per_pass = 256 / nr;
    double ar = 0.0, af = 0.0;
    for (long long d = (long long)blockIdx.x * rows_per_pass + threadIdx.x / nr; d < n_dof; d += (long long)gridDim.x * rows_per_pass) {
        const size_t o = (size_t)d * nr + c;
        double fv = f[o], rv = 0.0;
        if (fixed[d]) fv = 0.0; else rv = fv - kx[o];
        r[o] = rv;
        ar = fma(rv, rv, ar); af = fma(fv, fv, af);
    }
    atomicAdd(&s_r[c], ar); atomicAdd(&s_f[c], af);
    __syncthreads();
    if (threadIdx.x < nr) { atomicAdd(&norms2[threadIdx.x], s_r[threadIdx.x]); atomicAdd(&norms2[nr + threadIdx.x], s_f[threadIdx.x]); }
}

// ---- host driver ---------------------------------------------------------------------------------------
struct CholProfile {
    // ms per kernel class of the last profiled factorisation / substitution
    double init = 0, scatter = 0, potrf = 0, trsm = 0, update = 0, syrk = 0, factor_total = 0;
    double oz = 0;           // Schur updates that went through the int8 tensor-core kernel (not part of `syrk`)
    double solve_total = 0;
    long long launches_factor = 0, launches_solve = 0;
    // accumulated since the last fers_b200_chol_set_profile call
    double syrk_sum = 0, factor_sum = 0, solve_sum = 0, oz_sum = 0;
    long long n_factor = 0, n_solve = 0, n_syrk_launches = 0, n_oz_fronts = 0;
};
static CholProfile g_chol_profile;
// tile-shape selection (A/B measurements): bit 0 = panel updates, bit 1 = Schur update;
// set = 128x128 tile, one CTA per SM; clear = 128x64 tile, two CTAs per SM.  FERS_CHOL_VARIANT overrides.
static int g_chol_variant = getenv("FERS_CHOL_VARIANT") ? atoi(getenv("FERS_CHOL_VARIANT")) : 0;
// internal streams and events of the look-ahead, one set per device (created on first use on that device)
struct LookaheadState {
    cudaStream_t side = nullptr;   // low-priority stream of the overlapped (look-ahead) panel updates
    cudaStream_t main = nullptr;   // high-priority stream of the factorisation when look-ahead is on
    cudaEvent_t ev_main = nullptr, ev_side = nullptr, ev_fork = nullptr;
};
static LookaheadState g_lookahead[MAX_DEVICES];
// per-launch profiling (fers_b200_chol_set_profile) keeps process-wide state: one profiled call at a time
static std::mutex g_profile_mutex;
static int g_chol_profile_on = 0;      // 0 off, 1 every kernel class, 2 Schur-update spans + totals only
// Schur updates of the upper tree levels on the int8 tensor cores (ozaki_syrk.cu): digit planes per operand (0 = off,
// 6..8; 8 = fp64 grade) and the smallest own-column count of a level that takes that path.  FERS_CHOL_OZAKI /
// FERS_CHOL_OZAKI_MIN_K / fers_b200_chol_set_ozaki.
static int g_chol_ozaki = getenv("FERS_CHOL_OZAKI") ? atoi(getenv("FERS_CHOL_OZAKI")) : 7;
static int g_chol_ozaki_min_k = getenv("FERS_CHOL_OZAKI_MIN_K") ? atoi(getenv("FERS_CHOL_OZAKI_MIN_K")) : 768;
// the group updates of the panel factorisation (K = one group of 256-column blocks) of levels whose fronts have at least
// this many own columns take the same kernel, one CTA per tile on the low-priority stream of the look-ahead (0 = off).
// Measured at configs[3] with the 128 x 128 two-pass kernel: 379.6 -> 364.0 ms per pass with groups of 6 blocks
// (1536 columns); 370.6 with 1024, 366.1 with 2048, 365.1 with 3072; a loss with K = 512.  FERS_CHOL_OZAKI_PANEL_K
static int g_chol_ozaki_panel_k = getenv("FERS_CHOL_OZAKI_PANEL_K") ? atoi(getenv("FERS_CHOL_OZAKI_PANEL_K")) : 2048;
// ... on levels of at most this many fronts (FERS_CHOL_OZAKI_PANEL_NF)
static int g_chol_ozaki_panel_nf = getenv("FERS_CHOL_OZAKI_PANEL_NF") ? atoi(getenv("FERS_CHOL_OZAKI_PANEL_NF")) : 8;
// ... and only updates whose K (block or group width) is at least this long: a 256-column update spends as long in the
// tile prologue / epilogue as in its 8 MMA steps (measured: factorisation 385 -> 535 ms with K = 256)
static int g_chol_ozaki_panel_kw = getenv("FERS_CHOL_OZAKI_PANEL_KW") ? atoi(getenv("FERS_CHOL_OZAKI_PANEL_KW")) : 1536;

struct EventPool {
    std::vector<cudaEvent_t> ev;
    size_t used = 0;
    cudaEvent_t get() {
        if (used == ev.size()) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); }
        return ev[used++];
    }
    void reset() { used = 0; }
};
static EventPool g_pool;

struct Timed {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans[7];      // 6 = int8 tensor-core Schur updates
    void clear() { for (auto& s : spans) s.clear(); }
};
static Timed g_timed;

static int set_smem_attrs() {
    static PerDeviceOnce once;
    if (!once.needed()) return 0;
    FERS_CK(cudaFuncSetAttribute(k_potrf_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM));
    FERS_CK(cudaFuncSetAttribute(k_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G84::SMEM));
    FERS_CK(cudaFuncSetAttribute(k_row_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRP::SMEM));
    FERS_CK(cudaFuncSetAttribute(k_update<G88>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G88::SMEM));
    FERS_CK(cudaFuncSetAttribute(k_syrk<G88>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G88::SMEM));
    FERS_CK(cudaFuncSetAttribute(k_update<G86>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G86::SMEM));
    FERS_CK(cudaFuncSetAttribute(k_syrk<G86>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G86::SMEM));
    FERS_CK(cudaFuncSetAttribute(k_fwd_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G44::SMEM));
    FERS_CK(cudaFuncSetAttribute(k_fwd_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G48::SMEM));
    FERS_CK(cudaFuncSetAttribute(k_fwd_bnd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G48::SMEM));
    FERS_CK(cudaFuncSetAttribute(k_bwd_bnd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G44N::SMEM));
    FERS_CK(cudaFuncSetAttribute(k_bwd_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G44N::SMEM));
    FERS_CK(cudaFuncSetAttribute(k_bwd_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G48N::SMEM));
    FERS_CK(cudaFuncSetAttribute(k_fwd_wave, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G44::SMEM));
#ifndef FERS_CHOL_DFMA
    FERS_CK(cudaFuncSetAttribute(k_wave2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WAVE_SMEM));
    FERS_CK(cudaFuncSetAttribute(k_wave2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WAVE_SMEM));
#endif
    FERS_CK(cudaFuncSetAttribute(k_bwd_wave, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G44N::SMEM));
    once.done();
    return 0;
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

struct IndexPtrs {
    const FrontDev* fronts;
    const int* nodes;
    const int* child;
    const long long* inv_off;
    const int* inv;
    const int* level_fronts;
    const KScatter* kscatter;
    const int* pos;
};
static IndexPtrs index_ptrs(const CholPlanHost& P, const void* index_dev) {
    const unsigned char* b = static_cast<const unsigned char*>(index_dev);
    IndexPtrs I;
    I.fronts = reinterpret_cast<const FrontDev*>(b + P.off_fronts);
    I.nodes = reinterpret_cast<const int*>(b + P.off_nodes);
    I.child = reinterpret_cast<const int*>(b + P.off_child);
    I.inv_off = reinterpret_cast<const long long*>(b + P.off_inv_off);
    I.inv = reinterpret_cast<const int*>(b + P.off_inv);
    I.level_fronts = reinterpret_cast<const int*>(b + P.off_level_fronts);
    I.kscatter = reinterpret_cast<const KScatter*>(b + P.off_kscatter);
    I.pos = reinterpret_cast<const int*>(b + P.off_pos);
    return I;
}

// counters of the wavefront substitution: per level and direction one ticket + one progress counter per front
static size_t wave_counter_ints(const CholPlanHost& P) { return 2 * ((size_t)P.level_ptr[P.n_levels] + (size_t)P.n_levels); }
static size_t wave_counter_bytes(const CholPlanHost& P) { return align_up(wave_counter_ints(P) * sizeof(int)) + 256; }

static size_t chol_work_bytes(const CholPlanHost& P, int nrhs) {
    const size_t schur = (size_t)(P.schur_doubles[0] + P.schur_doubles[1]) * 8 + 512;
    const size_t solve = ((size_t)P.rhs_rows + 6 * (size_t)P.n_nodes) * (size_t)nrhs * 8 + 512 + wave_counter_bytes(P);
    return align_up(schur > solve ? schur : solve) + 256;
}

// levels [l_begin, l_end) of the front tree (the whole factorisation: 0 .. n_levels); the row scales and the pivot flag
// are set up by the call that starts at level 0
static int chol_factor(const CholPlanHost& P, const void* index_dev, const double* vals, const int* diag_slot,
                       const uint8_t* fixed, const double* kspring, double* factor, void* work, int* info_dev,
                       cudaStream_t caller, int l_begin = 0, int l_end = 1 << 30) {
    if (set_smem_attrs()) return 1;
    const IndexPtrs I = index_ptrs(P, index_dev);
    double* sc = factor + P.factor_doubles;
    double* arena[2];
    arena[0] = static_cast<double*>(work);
    arena[1] = arena[0] + P.schur_doubles[0];
    const bool prof = g_chol_profile_on != 0;
    const bool prof_all = g_chol_profile_on == 1;
    std::unique_lock<std::mutex> prof_lock(g_profile_mutex, std::defer_lock);
    if (prof) { prof_lock.lock(); g_pool.reset(); g_timed.clear(); }
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    if (prof) { ev_begin = g_pool.get(); cudaEventRecord(ev_begin, caller); }
    // look-ahead needs two internal streams (high priority: the panel chain and everything else; low priority:
    // the overlapped panel updates); per-class profiling (mode 1) keeps everything on the caller's stream
    const bool lookahead = !(g_chol_variant & 4) && !prof_all;
    LookaheadState& LA = g_lookahead[current_device()];
    if (lookahead && LA.side == nullptr) {
        int lo = 0, hi = 0;
        FERS_CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        FERS_CK(cudaStreamCreateWithPriority(&LA.side, cudaStreamNonBlocking, lo));
        FERS_CK(cudaStreamCreateWithPriority(&LA.main, cudaStreamNonBlocking, hi));
        FERS_CK(cudaEventCreateWithFlags(&LA.ev_main, cudaEventDisableTiming));
        FERS_CK(cudaEventCreateWithFlags(&LA.ev_side, cudaEventDisableTiming));
        FERS_CK(cudaEventCreateWithFlags(&LA.ev_fork, cudaEventDisableTiming));
    }
    cudaStream_t const g_side_stream = LA.side, g_main_stream = LA.main;
    cudaEvent_t const ev_main = LA.ev_main, ev_side = LA.ev_side, ev_fork = LA.ev_fork;
    cudaStream_t st = caller;
    if (lookahead) {
        FERS_CK(cudaEventRecord(ev_fork, caller));
        FERS_CK(cudaStreamWaitEvent(g_main_stream, ev_fork, 0));
        st = g_main_stream;
    }
    bool side_pending = false;
    long long launches = 0;
    auto span_begin = [&](int cls) -> cudaEvent_t {
        if (!prof || !(prof_all || cls >= 5)) return nullptr;
        cudaEvent_t e = g_pool.get();
        cudaEventRecord(e, st);
        return e;
    };
    auto span_end = [&](int cls, cudaEvent_t b) {
        if (b == nullptr) return;
        cudaEvent_t e = g_pool.get();
        cudaEventRecord(e, st);
        g_timed.spans[cls].push_back({b, e});
    };
    if (l_begin <= 0) {
        FERS_CK(cudaMemsetAsync(info_dev, 0, sizeof(int), st));
        k_chol_scale<<<cdiv(6 * P.n_nodes, 256), 256, 0, st>>>(P.n_nodes, vals, diag_slot, fixed, kspring, sc);
        FERS_LAUNCH_CHECK();
        ++launches;
    }
    for (int l = std::max(l_begin, 0); l < std::min(l_end, P.n_levels); ++l) {
        const int nf = P.level_ptr[l + 1] - P.level_ptr[l];
        if (nf == 0) continue;
        LevelCtx L;
        L.fronts = I.fronts; L.level_fronts = I.level_fronts + P.level_ptr[l]; L.nodes = I.nodes; L.child = I.child;
        L.inv_off = I.inv_off; L.inv = I.inv; L.factor = factor;
        L.schur_cur = arena[l & 1]; L.schur_child = arena[(l & 1) ^ 1];
        L.n_small = nf <= LEVEL_SMALL_FRONTS ? nf : 0;
        for (int q = 0; q < L.n_small; ++q) L.small_fronts[q] = P.fronts[P.level_fronts[P.level_ptr[l] + q]];
        const int max_m = 6 * P.level_max_n[l], max_k1 = 6 * P.level_max_n1[l];
        cudaEvent_t e0 = span_begin(0);
        {
            if (g_init_cols == 12) {
                dim3 grid(cdiv(max_m, INIT_ROWS), cdiv(max_m, 12), nf);
                k_front_init<12, 4><<<grid, INIT_THREADS, 0, st>>>(L);
            } else if (g_init_cols == 18) {
                dim3 grid(cdiv(max_m, INIT_ROWS), cdiv(max_m, 18), nf);
                k_front_init<18, 3><<<grid, INIT_THREADS, 0, st>>>(L);
            } else if (g_init_cols == 6) {
                dim3 grid(cdiv(max_m, INIT_ROWS), cdiv(max_m, 6), nf);
                k_front_init<6, 5><<<grid, INIT_THREADS, 0, st>>>(L);
            } else {
                dim3 grid(cdiv(max_m, INIT_ROWS), cdiv(max_m, 24), nf);
                k_front_init<24, 1><<<grid, INIT_THREADS, 0, st>>>(L);
            }
            FERS_LAUNCH_CHECK();
            ++launches;
        }
        span_end(0, e0);
        const long long ks0 = P.kscatter_ptr[l], ks1 = P.kscatter_ptr[l + 1];
        if (ks1 > ks0) {
            e0 = span_begin(1);
            k_scatter_k<<<(unsigned)((ks1 - ks0 + 7) / 8), 256, 0, st>>>(L, I.kscatter + ks0, ks1 - ks0, vals, fixed, kspring, sc);
            FERS_LAUNCH_CHECK();
            ++launches;
            span_end(1, e0);
        }
        const int steps = cdiv(max_k1, NB);
        for (int s = 0; s < steps; ++s) {
            e0 = span_begin(2);
            k_potrf_inv<<<nf, 256, POTRF_SMEM, st>>>(L, s, info_dev);
            span_end(2, e0);
            // blocking: inside the current outer block (OUTER steps of 64 columns) the trsm and the inner update of
            // every step touch only the rows of the block; the rows below it are brought up to date once per block
            // by k_row_panel; the outer update then reaches all later columns
            const int cbeg = (s + 1) * NB;
            const int CHOL_OUTER = g_chol_outer < 1 ? 1 : g_chol_outer;
            const int blk_end = (s / CHOL_OUTER + 1) * CHOL_OUTER * NB;
            const int in_end = std::min(blk_end, max_k1);
            if (cbeg < in_end) {
                e0 = span_begin(3);
                dim3 grid(cdiv(in_end - cbeg, G84::BM), 1, nf);
                k_trsm<<<grid, GEMM_THREADS, G84::SMEM, st>>>(L, s, blk_end);
                span_end(3, e0);
                ++launches;
                e0 = span_begin(4);
                if (g_chol_variant & 1) {
                    dim3 grid2(cdiv(in_end - cbeg, G88::BM), cdiv(in_end - cbeg, G88::BN), nf);
                    k_update<G88><<<grid2, GEMM_THREADS, G88::SMEM, st>>>(L, s * NB, NB, cbeg, blk_end, blk_end);
                } else {
                    dim3 grid2(cdiv(in_end - cbeg, G86::BM), cdiv(in_end - cbeg, G86::BN), nf);
                    k_update<G86><<<grid2, GEMM_THREADS, G86::SMEM, st>>>(L, s * NB, NB, cbeg, blk_end, blk_end);
                }
                span_end(4, e0);
                ++launches;
            }
            if (cbeg >= in_end || s == steps - 1) {          // block finished: rows below it
                const int bs = blk_end - CHOL_OUTER * NB;
                if (max_m > bs) {
                    e0 = span_begin(3);
                    dim3 grid(cdiv(max_m - bs, GRP::BM), 1, nf);
                    k_row_panel<<<grid, GEMM_THREADS, GRP::SMEM, st>>>(L, bs, blk_end);
                    span_end(3, e0);
                    ++launches;
                }
            }
            if (cbeg == blk_end && blk_end < max_k1) {
                // third level: blocks are gathered in groups of g_chol_group; inside a group a finished block updates
                // only the rest of its group (K = block width), a finished group updates all later columns at once
                // (K = group width: fewer passes over the trailing panel, longer GEMM main loops)
                const int blk_w = CHOL_OUTER * NB;
                const int grp_w = (g_chol_group < 1 ? 1 : g_chol_group) * blk_w;
                const int grp_end = ((blk_end - 1) / grp_w + 1) * grp_w;
                const bool at_group_end = blk_end == grp_end;
                const int kw = at_group_end ? grp_w : blk_w;
                const int kc0 = blk_end - kw;
                const int lim = at_group_end ? max_k1 : std::min(grp_end, max_k1);     // columns [blk_end, lim) are updated
                const int lim_arg = at_group_end ? (1 << 30) : grp_end;
                // fronts with many own columns: the update of all later columns as int8 digit-plane products (ozaki_syrk.cu),
                // front by front; digit planes in the children's arena (the Schur update of the level reuses it afterwards)
                const bool oz_panel = g_chol_ozaki >= 4 && g_chol_ozaki_panel_k > 0 && nf <= g_chol_ozaki_panel_nf && kw < 16384 &&
                                      kw >= g_chol_ozaki_panel_kw &&
                                      6 * P.level_max_n1[l] >= g_chol_ozaki_panel_k &&
                                      oz_work_bytes(max_m, kw, g_chol_ozaki) <= (size_t)P.schur_doubles[(l & 1) ^ 1] * 8;
                auto outer = [&](cudaStream_t q, int c_from, int c_to_grid, int c_to) {
                    if (oz_panel && q != st) {
                        for (int z = P.level_ptr[l]; z < P.level_ptr[l + 1]; ++z) {
                            const FrontDev& f = P.fronts[P.level_fronts[z]];
                            const int fk1 = 6 * f.n1, fm = 6 * (f.n1 + f.n2);
                            const int hi = std::min(c_to, fk1);
                            if (c_from >= hi) continue;
                            double* panel = factor + f.panel_off;
                            oz_syrk(panel + (size_t)kc0 * fm + c_from, fm, panel + (size_t)c_from * fm + c_from, fm, fm - c_from, kw,
                                    g_chol_ozaki, arena[(l & 1) ^ 1], q, hi - c_from, false);
                            launches += 3;
                        }
                        return;
                    }
                    if (g_chol_variant & 1) {
                        dim3 grid(cdiv(max_m - c_from, G88::BM), cdiv(c_to_grid - c_from, G88::BN), nf);
                        k_update<G88><<<grid, GEMM_THREADS, G88::SMEM, q>>>(L, kc0, kw, c_from, c_to, 1 << 30);
                    } else {
                        dim3 grid(cdiv(max_m - c_from, G86::BM), cdiv(c_to_grid - c_from, G86::BN), nf);
                        k_update<G86><<<grid, GEMM_THREADS, G86::SMEM, q>>>(L, kc0, kw, c_from, c_to, 1 << 30);
                    }
                    ++launches;
                };
                if (!lookahead) {
                    e0 = span_begin(4);
                    outer(st, blk_end, lim, lim_arg);
                    span_end(4, e0);
                } else {
                    // look-ahead: the columns of the next block are updated first, on the main stream, so that its
                    // chain of diagonal tiles / trsm / inner updates overlaps the update of the remaining columns,
                    // which runs on the side stream.  Both touch disjoint columns; successive side updates are
                    // ordered by the side stream, and the urgent update of block b+1 waits for the side update of
                    // block b (same columns).
                    const int next_end = std::min(blk_end + blk_w, lim);
                    if (side_pending) { FERS_CK(cudaStreamWaitEvent(st, ev_side, 0)); side_pending = false; }
                    outer(st, blk_end, next_end, next_end);
                    if (next_end < lim) {
                        FERS_CK(cudaEventRecord(ev_main, st));
                        FERS_CK(cudaStreamWaitEvent(g_side_stream, ev_main, 0));
                        outer(g_side_stream, next_end, lim, lim_arg);
                        FERS_CK(cudaEventRecord(ev_side, g_side_stream));
                        side_pending = true;
                    }
                }
            }
            FERS_LAUNCH_CHECK();
            ++launches;
        }
        if (side_pending) { FERS_CK(cudaStreamWaitEvent(st, ev_side, 0)); side_pending = false; }
        const int max_m2 = max_m;   // upper bound on 6 * n2
        // levels whose fronts all have many own columns: every front on its own through the int8 tensor-core kernel (the
        // digit planes live in the children's Schur arena, free since this level's extend-add)
        bool ozaki = false;
        if (l + 1 < P.n_levels && g_chol_ozaki >= 4 && nf <= 64) {
            ozaki = true;
            const size_t room = (size_t)P.schur_doubles[(l & 1) ^ 1] * 8;
            for (int q = P.level_ptr[l]; q < P.level_ptr[l + 1]; ++q) {
                const FrontDev& f = P.fronts[P.level_fronts[q]];
                if (6 * f.n1 < g_chol_ozaki_min_k || 6 * f.n1 >= 16384 || 6 * f.n2 < 1024 ||
                    oz_work_bytes(6 * f.n2, 6 * f.n1, g_chol_ozaki) > room)
                    ozaki = false;
            }
        }
        if (ozaki) {
            e0 = span_begin(6);
            for (int q = P.level_ptr[l]; q < P.level_ptr[l + 1]; ++q) {
                const FrontDev& f = P.fronts[P.level_fronts[q]];
                const int k1 = 6 * f.n1, m2 = 6 * f.n2, m = k1 + m2;
                if (oz_syrk(factor + f.panel_off + k1, m, L.schur_cur + f.schur_off, m2, m2, k1, g_chol_ozaki,
                            arena[(l & 1) ^ 1], st))
                    return 1;
                launches += 3;
                if (prof) g_chol_profile.n_oz_fronts += 1;
            }
            span_end(6, e0);
        } else if (l + 1 < P.n_levels) {
            e0 = span_begin(5);
            if (g_chol_variant & 2) {
                dim3 grid(cdiv(max_m2, G88::BM), cdiv(max_m2, G88::BN), nf);
                k_syrk<G88><<<grid, GEMM_THREADS, G88::SMEM, st>>>(L);
            } else {
                dim3 grid(cdiv(max_m2, G86::BM), cdiv(max_m2, G86::BN), nf);
                k_syrk<G86><<<grid, GEMM_THREADS, G86::SMEM, st>>>(L);
            }
            FERS_LAUNCH_CHECK();
            span_end(5, e0);
            ++launches;
        }
    }
    if (lookahead) {                       // join: the caller's stream continues after the internal one
        FERS_CK(cudaEventRecord(ev_fork, st));
        FERS_CK(cudaStreamWaitEvent(caller, ev_fork, 0));
    }
    if (prof) { ev_end = g_pool.get(); cudaEventRecord(ev_end, caller); }
    count_launches(launches);
    if (prof) {
        FERS_CK(cudaStreamSynchronize(caller));
        double t[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < 7; ++c)
            for (auto& sp : g_timed.spans[c]) { float ms = 0; cudaEventElapsedTime(&ms, sp.first, sp.second); t[c] += ms; }
        float tot = 0;
        cudaEventElapsedTime(&tot, ev_begin, ev_end);
        g_chol_profile.init = t[0]; g_chol_profile.scatter = t[1]; g_chol_profile.potrf = t[2];
        g_chol_profile.trsm = t[3]; g_chol_profile.update = t[4]; g_chol_profile.syrk = t[5]; g_chol_profile.oz = t[6];
        g_chol_profile.oz_sum += t[6];
        g_chol_profile.factor_total = tot;
        g_chol_profile.launches_factor = launches;
        g_chol_profile.syrk_sum += t[5]; g_chol_profile.factor_sum += tot; g_chol_profile.n_factor += 1;
        g_chol_profile.n_syrk_launches += (long long)g_timed.spans[5].size();
    }
    return 0;
}

// phases: 1 = forward substitution over levels [l_begin, l_end), 2 = backward over the same levels (top down),
// 4 = permuted solution -> x; the whole substitution is 7 over 0 .. n_levels
static int chol_solve(const CholPlanHost& P, const void* index_dev, const double* factor, const uint8_t* fixed, int nrhs,
                      const double* b, double* x, int accumulate, void* work, cudaStream_t st, int phases = 7, int l_begin = 0,
                      int l_end = 1 << 30) {
    l_begin = std::max(l_begin, 0);
    l_end = std::min(l_end, P.n_levels);
    if (set_smem_attrs()) return 1;
    const IndexPtrs I = index_ptrs(P, index_dev);
    const double* sc = factor + P.factor_doubles;
    SolveCtx S;
    S.fronts = I.fronts; S.nodes = I.nodes; S.child = I.child; S.inv_off = I.inv_off; S.inv = I.inv; S.pos = I.pos;
    S.factor = factor;
    S.W = static_cast<double*>(work);
    S.Y = S.W + (size_t)P.rhs_rows * nrhs;
    S.nrhs = nrhs;
    long long launches = 0;
    const bool prof = g_chol_profile_on != 0;
    std::unique_lock<std::mutex> prof_lock(g_profile_mutex, std::defer_lock);
    if (prof) prof_lock.lock();
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    if (prof) { cudaEventCreate(&ev_begin); cudaEventCreate(&ev_end); cudaEventRecord(ev_begin, st); }
    const bool wave = !(g_chol_variant & 8);
    int* ctr = reinterpret_cast<int*>(
        (reinterpret_cast<uintptr_t>(S.Y + (size_t)6 * P.n_nodes * nrhs) + 255) & ~(uintptr_t)255);
    const size_t ctr_half = (size_t)P.level_ptr[P.n_levels] + (size_t)P.n_levels;
    if (wave && (phases & 1) && l_begin == 0) FERS_CK(cudaMemsetAsync(ctr, 0, wave_counter_ints(P) * sizeof(int), st));
    for (int l = l_begin; (phases & 1) && l < l_end; ++l) {                // forward: children -> parents
        const int nf = P.level_ptr[l + 1] - P.level_ptr[l];
        if (nf == 0) continue;
        S.level_fronts = I.level_fronts + P.level_ptr[l];
        const int max_m = 6 * P.level_max_n[l], max_k1 = 6 * P.level_max_n1[l];
        {
            const long long total = (long long)max_m * nrhs;
            dim3 grid((unsigned)std::min<long long>((total + 255) / 256, 1024), nf);
            k_fwd_gather<<<grid, 256, 0, st>>>(S, b, fixed, sc);
            ++launches;
        }
        const int steps = cdiv(max_k1, NB);
        if (wave) {
#ifndef FERS_CHOL_DFMA
            if (!(g_chol_variant & 16)) k_wave2<false><<<nf * steps, GEMM_THREADS, WAVE_SMEM, st>>>(S, nf, ctr + P.level_ptr[l] + l);
            else
#endif
            k_fwd_wave<<<nf * steps, GEMM_THREADS, G44::SMEM, st>>>(S, nf, ctr + P.level_ptr[l] + l);
            ++launches;
        } else
        for (int s = 0; s < steps; ++s) {
            k_fwd_diag<<<nf, GEMM_THREADS, G44::SMEM, st>>>(S, s);
            ++launches;
            const int rows_below = max_k1 - (s + 1) * NB;     // own rows after this step
            if (rows_below > 0) {
                dim3 grid(cdiv(rows_below, G48::BN), nf);
                k_fwd_update<<<grid, GEMM_THREADS, G48::SMEM, st>>>(S, s);
                ++launches;
            }
        }
        if (l + 1 < P.n_levels) {                              // the root level has no boundary
            dim3 grid(cdiv(max_m, G48::BN), nf);
            k_fwd_bnd<<<grid, GEMM_THREADS, G48::SMEM, st>>>(S);
            ++launches;
        }
        FERS_LAUNCH_CHECK();
    }
    for (int l = l_end - 1; (phases & 2) && l >= l_begin; --l) {           // backward: parents -> children
        const int nf = P.level_ptr[l + 1] - P.level_ptr[l];
        if (nf == 0) continue;
        S.level_fronts = I.level_fronts + P.level_ptr[l];
        const int max_m = 6 * P.level_max_n[l], max_k1 = 6 * P.level_max_n1[l];
        if (l + 1 < P.n_levels) {
            const long long total = (long long)max_m * nrhs;
            dim3 grid((unsigned)std::min<long long>((total + 255) / 256, 1024), nf);
            k_bwd_gather<<<grid, 256, 0, st>>>(S);
            dim3 grid2(cdiv(max_k1, G44N::BN), nf);
            k_bwd_bnd<<<grid2, GEMM_THREADS, G44N::SMEM, st>>>(S);
            launches += 2;
        }
        const int steps = cdiv(max_k1, NB);
        if (wave) {
#ifndef FERS_CHOL_DFMA
            if (!(g_chol_variant & 16)) k_wave2<true><<<nf * steps, GEMM_THREADS, WAVE_SMEM, st>>>(S, nf, ctr + ctr_half + P.level_ptr[l] + l);
            else
#endif
            k_bwd_wave<<<nf * steps, GEMM_THREADS, G44N::SMEM, st>>>(S, nf, ctr + ctr_half + P.level_ptr[l] + l);
            ++launches;
        } else
        for (int s = steps - 1; s >= 0; --s) {
            k_bwd_diag<<<nf, GEMM_THREADS, G44N::SMEM, st>>>(S, s);
            ++launches;
            if (s > 0) {
                dim3 grid(cdiv(s * NB, G48N::BN), nf);
                k_bwd_update<<<grid, GEMM_THREADS, G48N::SMEM, st>>>(S, s);
                ++launches;
            }
        }
        FERS_LAUNCH_CHECK();
    }
    if (phases & 4) {
        const long long total = (long long)P.n_nodes * 6 * nrhs;
        k_chol_unpermute<<<(unsigned)std::min<long long>((total + 255) / 256, 65535), 256, 0, st>>>(
            P.n_nodes, nrhs, I.pos, fixed, sc, S.Y, x, accumulate);
        FERS_LAUNCH_CHECK();
        ++launches;
    }
    count_launches(launches);
    if (prof) {
        cudaEventRecord(ev_end, st);
        FERS_CK(cudaStreamSynchronize(st));
        float tot = 0;
        cudaEventElapsedTime(&tot, ev_begin, ev_end);
        g_chol_profile.solve_total = tot;
        g_chol_profile.launches_solve = launches;
        g_chol_profile.solve_sum += tot; g_chol_profile.n_solve += 1;
        cudaEventDestroy(ev_begin); cudaEventDestroy(ev_end);
    }
    return 0;
}

}  // namespace fers

using namespace fers;

extern "C" {

size_t fers_b200_chol_factor_bytes(void* plan) {
    if (!plan) return 0;
    const CholPlanHost& P = *static_cast<CholPlanHost*>(plan);
    return align_up(((size_t)P.factor_doubles + 6 * (size_t)P.n_nodes) * 8);
}

size_t fers_b200_chol_work_bytes(void* plan, int nrhs) {
    if (!plan) return 0;
    return chol_work_bytes(*static_cast<CholPlanHost*>(plan), nrhs < 1 ? 1 : nrhs);
}

int fers_b200_chol_upload(void* plan, void* index_dev, void* stream) {
    FERS_REQUIRE(plan && index_dev, "chol_upload: bad arguments");
    CholPlanHost& P = *static_cast<CholPlanHost*>(plan);
    // section by section straight from the plan's arrays (no zero-filled staging copy of the tens of MB of maps)
    cudaStream_t st = as_stream(stream);
    unsigned char* d = static_cast<unsigned char*>(index_dev);
    auto put = [&](size_t off, const void* src, size_t bytes) -> int {
        if (bytes == 0) return 0;
        FERS_CK(cudaMemcpyAsync(d + off, src, bytes, cudaMemcpyHostToDevice, st));
        return 0;
    };
    if (put(P.off_fronts, P.fronts.data(), sizeof(FrontDev) * P.fronts.size())) return 1;
    if (put(P.off_nodes, P.nodes.data(), sizeof(int) * P.nodes.size())) return 1;
    if (put(P.off_child, P.child.data(), sizeof(int) * P.child.size())) return 1;
    if (put(P.off_inv_off, P.inv_off.data(), sizeof(long long) * P.inv_off.size())) return 1;
    if (put(P.off_inv, P.inv.data(), sizeof(int) * P.inv.size())) return 1;
    if (put(P.off_level_fronts, P.level_fronts.data(), sizeof(int) * P.level_fronts.size())) return 1;
    if (put(P.off_kscatter, P.kscatter.data(), sizeof(KScatter) * P.kscatter.size())) return 1;
    if (put(P.off_pos, P.pos.data(), sizeof(int) * P.pos.size())) return 1;
    FERS_CK(cudaStreamSynchronize(st));
    return 0;
}

int fers_b200_chol_factor(void* plan, const void* index_dev, const double* vals, const int* diag_slot,
                          const uint8_t* fixed, const double* kspring, double* factor, void* work, size_t work_bytes,
                          int* info_dev, int* info_out, void* stream) {
    FERS_REQUIRE(plan && index_dev && vals && diag_slot && fixed && factor && work && info_dev, "chol_factor: bad arguments");
    const CholPlanHost& P = *static_cast<CholPlanHost*>(plan);
    FERS_REQUIRE(work_bytes >= chol_work_bytes(P, 1), "chol_factor: workspace too small");
    const int rc = chol_factor(P, index_dev, vals, diag_slot, fixed, kspring, factor, work, info_dev, as_stream(stream));
    if (rc) return rc;
    if (info_out) {
        FERS_CK(cudaMemcpyAsync(info_out, info_dev, sizeof(int), cudaMemcpyDeviceToHost, as_stream(stream)));
        FERS_CK(cudaStreamSynchronize(as_stream(stream)));
    }
    return 0;
}

int fers_b200_chol_factor_levels(void* plan, const void* index_dev, const double* vals, const int* diag_slot,
                                 const uint8_t* fixed, const double* kspring, double* factor, void* work, size_t work_bytes,
                                 int* info_dev, int l_begin, int l_end, void* stream) {
    FERS_REQUIRE(plan && index_dev && vals && diag_slot && fixed && factor && work && info_dev, "chol_factor_levels: bad arguments");
    const CholPlanHost& P = *static_cast<CholPlanHost*>(plan);
    FERS_REQUIRE(work_bytes >= chol_work_bytes(P, 1), "chol_factor_levels: workspace too small");
    FERS_REQUIRE(0 <= l_begin && l_begin <= l_end, "chol_factor_levels: bad level range");
    return chol_factor(P, index_dev, vals, diag_slot, fixed, kspring, factor, work, info_dev, as_stream(stream), l_begin, l_end);
}

int fers_b200_chol_solve_phase(void* plan, const void* index_dev, const double* factor, const uint8_t* fixed, int nrhs,
                               const double* b, double* x, int accumulate, void* work, size_t work_bytes, int phases,
                               int l_begin, int l_end, void* stream) {
    FERS_REQUIRE(plan && index_dev && factor && fixed && b && x && work, "chol_solve_phase: bad arguments");
    FERS_REQUIRE(nrhs >= 2 && nrhs <= 64 && (nrhs % 2) == 0, "chol_solve_phase: nrhs must be even and in 2..64");
    FERS_REQUIRE(phases > 0 && phases < 8 && 0 <= l_begin && l_begin <= l_end, "chol_solve_phase: bad phases / level range");
    const CholPlanHost& P = *static_cast<CholPlanHost*>(plan);
    FERS_REQUIRE(work_bytes >= chol_work_bytes(P, nrhs), "chol_solve_phase: workspace too small");
    return chol_solve(P, index_dev, factor, fixed, nrhs, b, x, accumulate, work, as_stream(stream), phases, l_begin, l_end);
}

int fers_b200_chol_solve(void* plan, const void* index_dev, const double* factor, const uint8_t* fixed, int nrhs,
                         const double* b, double* x, int accumulate, void* work, size_t work_bytes, void* stream) {
    FERS_REQUIRE(plan && index_dev && factor && fixed && b && x && work, "chol_solve: bad arguments");
    FERS_REQUIRE(nrhs >= 2 && nrhs <= 64 && (nrhs % 2) == 0, "chol_solve: nrhs must be even and in 2..64 (pad the right-hand sides)");
    const CholPlanHost& P = *static_cast<CholPlanHost*>(plan);
    FERS_REQUIRE(work_bytes >= chol_work_bytes(P, nrhs), "chol_solve: workspace too small");
    return chol_solve(P, index_dev, factor, fixed, nrhs, b, x, accumulate, work, as_stream(stream));
}

int fers_b200_residual(long long n_dof, int nrhs, const double* f, const double* kx, const uint8_t* fixed, double* r,
                       double* norms2, void* stream) {
    FERS_REQUIRE(f && kx && fixed && r && norms2, "residual: bad arguments");
    FERS_REQUIRE(nrhs >= 1 && nrhs <= 64 && 256 % nrhs == 0, "residual: nrhs must divide 256 and be <= 64");
    cudaStream_t st = as_stream(stream);
    FERS_CK(cudaMemsetAsync(norms2, 0, 2 * (size_t)nrhs * sizeof(double), st));
    const int rows_per_pass = 256 / nrhs;
    long long grid = (n_dof + rows_per_pass - 1) / rows_per_pass;
    const long long cap = (long long)sm_count() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    k_chol_residual<<<(unsigned)grid, 256, 0, st>>>(n_dof, nrhs, f, kx, fixed, r, norms2);
    FERS_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int fers_b200_chol_set_variant(int variant) {
    FERS_REQUIRE(variant >= 0 && variant <= 31, "chol_set_variant: bit 0 = panel updates, bit 1 = Schur update use the 128x128 tile, bit 2 = no look-ahead, bit 3 = per-step substitution kernels instead of the wavefront, bit 4 = first wavefront version");
    g_chol_variant = variant;
    return 0;
}

int fers_b200_chol_set_outer(int steps, int group) {
    FERS_REQUIRE(steps >= 1 && steps <= 64, "chol_set_outer: 1..64 steps of 64 columns per outer block");
    FERS_REQUIRE(group >= 1 && group <= 64, "chol_set_outer: 1..64 outer blocks per group");
    g_chol_outer = steps;
    g_chol_group = group;
    return 0;
}

int fers_b200_chol_set_ozaki(int slices, int min_k) {
    FERS_REQUIRE(slices == 0 || (slices >= 4 && slices <= 8), "chol_set_ozaki: slices 0 (off) or 4..8");
    FERS_REQUIRE(min_k >= 64, "chol_set_ozaki: min_k >= 64");
    g_chol_ozaki = slices;
    g_chol_ozaki_min_k = min_k;
    return 0;
}

int fers_b200_chol_set_profile(int on) {
    g_chol_profile_on = on < 0 ? 0 : (on > 2 ? 1 : on);
    g_chol_profile.syrk_sum = g_chol_profile.factor_sum = g_chol_profile.solve_sum = g_chol_profile.oz_sum = 0;
    g_chol_profile.n_factor = g_chol_profile.n_solve = g_chol_profile.n_syrk_launches = g_chol_profile.n_oz_fronts = 0;
    return 0;
}

/* out[10]: ms of init, scatter, potrf, trsm, update, syrk, factor total, solve total; launches factor, solve;
 * n_out >= 16 adds sums since the last set_profile: syrk ms, factor ms, solve ms, factorisations, substitutions, syrk launches;
 * n_out >= 20 adds the int8 tensor-core Schur updates: ms (last), ms (sum), fronts (sum), slices * 100000 + min_k */
int fers_b200_chol_profile(double* out, int n_out) {
    FERS_REQUIRE(out && n_out >= 10, "chol_profile: need 10 outputs");
    const CholProfile& p = g_chol_profile;
    out[0] = p.init; out[1] = p.scatter; out[2] = p.potrf; out[3] = p.trsm; out[4] = p.update; out[5] = p.syrk;
    out[6] = p.factor_total; out[7] = p.solve_total; out[8] = (double)p.launches_factor; out[9] = (double)p.launches_solve;
    if (n_out >= 16) {
        out[10] = p.syrk_sum; out[11] = p.factor_sum; out[12] = p.solve_sum;
        out[13] = (double)p.n_factor; out[14] = (double)p.n_solve; out[15] = (double)p.n_syrk_launches;
    }
    if (n_out >= 20) {      // int8 tensor-core Schur updates: ms of the last profiled factorisation, ms and fronts since set_profile, settings
        out[16] = p.oz; out[17] = p.oz_sum; out[18] = (double)p.n_oz_fronts; out[19] = (double)(g_chol_ozaki * 100000 + g_chol_ozaki_min_k);
    }
    return 0;
}

}  // extern "C"
