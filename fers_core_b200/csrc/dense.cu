// dense.cu -- dense direct path for small models and batches of small models
// (BASELINE.json configs[0..2] and the per-frame systems of configs[4]).
//
// Blocked right-looking Cholesky with 32x32 tiles, batched over blockIdx.z; row-major storage,
// lower triangle holds L on exit.  Three kernels per block column (diagonal factor, panel solve,
// trailing update), two per block column and direction for the multi-RHS substitution.
// The matrices of this path live in L2 (n <= ~4k), so the kernels are latency/FP64-pipe bound,
// not HBM bound; the HBM-bound large-model path is solve_pcg.cu.
#include "common.cuh"

namespace fers {

constexpr int NB = 32;
constexpr int NBP = NB + 1;   // padded shared-memory stride

// expand block-CSR to dense with constrained DOFs replaced by identity rows/cols
__global__ void k_dense_expand(int n_nodes, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                               const double* __restrict__ vals, const uint8_t* __restrict__ fixed,
                               const double* __restrict__ kspring, double* __restrict__ a) {
    const int n = 6 * n_nodes;
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_nodes) return;
    for (int k = rowptr[row]; k < rowptr[row + 1]; ++k) {
        const int col = colidx[k];
        for (int e = lane; e < 36; e += 32) {
            const int r = e / 6, c = e % 6;
            const int gi = 6 * row + r, gj = 6 * col + c;
            double v = vals[(size_t)k * 36 + e];
            const bool fx = (fixed[gi] != 0) || (fixed[gj] != 0);
            if (fx) v = (gi == gj) ? 1.0 : 0.0;
            else if (gi == gj && kspring != nullptr) v += kspring[gi];
            a[(size_t)gi * n + gj] = v;
        }
    }
}

// batched variant: the model is a set of `batch` disconnected frames of nodes_per_frame nodes each
// (numbered frame by frame); frame f is expanded into its own dense matrix a[f][n][n], n = 6*npf
__global__ void k_dense_expand_batched(int npf, int batch, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                       const double* __restrict__ vals, const uint8_t* __restrict__ fixed,
                                       const double* __restrict__ kspring, double* __restrict__ a, int* __restrict__ bad,
                                       const uint8_t* __restrict__ active) {
    const int n = 6 * npf;
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= npf * batch) return;
    const int f = row / npf, lr = row - f * npf;
    if (active != nullptr && !active[f]) return;           // frozen frame of a P-Delta batch: its matrix is not needed again
    double* af = a + (size_t)f * n * n;
    for (int k = rowptr[row]; k < rowptr[row + 1]; ++k) {
        const int col = colidx[k];
        const int lc = col - f * npf;
        if (lc < 0 || lc >= npf) {           // a member crosses frames: not a block-diagonal batch
            if (lane == 0) atomicExch(bad, 1);
            continue;
        }
        for (int e = lane; e < 36; e += 32) {
            const int r = e / 6, c = e % 6;
            const int gi = 6 * row + r, gj = 6 * col + c;
            double v = vals[(size_t)k * 36 + e];
            const bool fx = (fixed[gi] != 0) || (fixed[gj] != 0);
            if (fx) v = (gi == gj) ? 1.0 : 0.0;
            else if (gi == gj && kspring != nullptr) v += kspring[gi];
            af[(size_t)(6 * lr + r) * n + 6 * lc + c] = v;
        }
    }
}

// factor the diagonal tile [j0, j0+nb)
__global__ void __launch_bounds__(NB* NB) k_potrf_diag(double* __restrict__ a_all, int n, int j0, int* __restrict__ info) {
    __shared__ double t[NB][NBP];
    double* a = a_all + (size_t)blockIdx.z * n * n;
    const int tx = threadIdx.x % NB, ty = threadIdx.x / NB;
    const int nb = min(NB, n - j0);
    const bool in = (ty < nb && tx < nb);
    t[ty][tx] = (in && tx <= ty) ? a[(size_t)(j0 + ty) * n + j0 + tx] : ((ty == tx) ? 1.0 : 0.0);
    __syncthreads();
    for (int k = 0; k < nb; ++k) {
        if (tx == k && ty == k) {
            double d = t[k][k];
            if (!(d > 0.0)) {
                if (info && info[blockIdx.z] == 0) info[blockIdx.z] = j0 + k + 1;
                d = 1.0;
            }
            t[k][k] = sqrt(d);
        }
        __syncthreads();
        if (tx == k && ty > k && ty < nb) t[ty][k] /= t[k][k];
        __syncthreads();
        if (tx > k && tx <= ty && ty < nb) t[ty][tx] -= t[ty][k] * t[tx][k];
        __syncthreads();
    }
    if (in && tx <= ty) a[(size_t)(j0 + ty) * n + j0 + tx] = t[ty][tx];
}

// panel: rows below the diagonal tile, X * L_jj^T = A_panel
__global__ void __launch_bounds__(NB* NB) k_trsm_panel(double* __restrict__ a_all, int n, int j0) {
    __shared__ double l[NB][NBP];
    __shared__ double b[NB][NBP];
    double* a = a_all + (size_t)blockIdx.z * n * n;
    const int tx = threadIdx.x % NB, ty = threadIdx.x / NB;
    const int nb = min(NB, n - j0);
    const int i0 = j0 + NB + blockIdx.x * NB;
    const int rows = min(NB, n - i0);
    l[ty][tx] = (ty < nb && tx < nb && tx <= ty) ? a[(size_t)(j0 + ty) * n + j0 + tx] : ((ty == tx) ? 1.0 : 0.0);
    b[ty][tx] = (ty < rows && tx < nb) ? a[(size_t)(i0 + ty) * n + j0 + tx] : 0.0;
    __syncthreads();
    for (int k = 0; k < nb; ++k) {
        if (tx == k) b[ty][k] /= l[k][k];
        __syncthreads();
        if (tx > k && tx < nb) b[ty][tx] -= b[ty][k] * l[tx][k];
        __syncthreads();
    }
    if (ty < rows && tx < nb) a[(size_t)(i0 + ty) * n + j0 + tx] = b[ty][tx];
}

// trailing update: A[ti][tk] -= P_ti P_tk^T for tiles ti >= tk below/right of the panel
__global__ void __launch_bounds__(NB* NB) k_syrk_update(double* __restrict__ a_all, int n, int j0) {
    if (blockIdx.y > blockIdx.x) return;
    __shared__ double pi[NB][NBP];
    __shared__ double pk[NB][NBP];
    double* a = a_all + (size_t)blockIdx.z * n * n;
    const int tx = threadIdx.x % NB, ty = threadIdx.x / NB;
    const int nb = min(NB, n - j0);
    const int i0 = j0 + NB + blockIdx.x * NB, k0 = j0 + NB + blockIdx.y * NB;
    const int ri = min(NB, n - i0), rk = min(NB, n - k0);
    pi[ty][tx] = (ty < ri && tx < nb) ? a[(size_t)(i0 + ty) * n + j0 + tx] : 0.0;
    pk[ty][tx] = (ty < rk && tx < nb) ? a[(size_t)(k0 + ty) * n + j0 + tx] : 0.0;
    __syncthreads();
    if (ty < ri && tx < rk && (k0 + tx) <= (i0 + ty)) {
        double s = 0.0;
#pragma unroll 8
        for (int q = 0; q < NB; ++q) s = fma(pi[ty][q], pk[tx][q], s);
        a[(size_t)(i0 + ty) * n + k0 + tx] -= s;
    }
}

// ---- multi-RHS substitution; bx [batch][n][nrhs] -----------------------------------------------
// diagonal solve of block j: forward (L) or backward (L^T); one CTA per 32 rhs columns
__global__ void __launch_bounds__(NB* NB) k_trsv_diag(const double* __restrict__ a_all, double* __restrict__ bx_all, int n,
                                                      int nrhs, int j0, int backward) {
    __shared__ double l[NB][NBP];
    __shared__ double b[NB][NBP];
    const double* a = a_all + (size_t)blockIdx.z * n * n;
    double* bx = bx_all + (size_t)blockIdx.z * n * nrhs;
    const int tx = threadIdx.x % NB, ty = threadIdx.x / NB;
    const int nb = min(NB, n - j0);
    const int c0 = blockIdx.x * NB;
    const int col = c0 + tx;
    l[ty][tx] = (ty < nb && tx < nb && tx <= ty) ? a[(size_t)(j0 + ty) * n + j0 + tx] : ((ty == tx) ? 1.0 : 0.0);
    b[ty][tx] = (ty < nb && col < nrhs) ? bx[(size_t)(j0 + ty) * nrhs + col] : 0.0;
    __syncthreads();
    if (!backward) {
        for (int k = 0; k < nb; ++k) {
            if (ty == k) b[k][tx] /= l[k][k];
            __syncthreads();
            if (ty > k && ty < nb) b[ty][tx] -= l[ty][k] * b[k][tx];
            __syncthreads();
        }
    } else {
        for (int k = nb - 1; k >= 0; --k) {
            if (ty == k) b[k][tx] /= l[k][k];
            __syncthreads();
            if (ty < k) b[ty][tx] -= l[k][ty] * b[k][tx];
            __syncthreads();
        }
    }
    if (ty < nb && col < nrhs) bx[(size_t)(j0 + ty) * nrhs + col] = b[ty][tx];
}

// forward: b_i -= L_ij x_j (i > j);  backward: b_k -= L_jk^T x_j (k < j)
__global__ void __launch_bounds__(NB* NB) k_subst_update(const double* __restrict__ a_all, double* __restrict__ bx_all, int n,
                                                         int nrhs, int j0, int backward) {
    __shared__ double l[NB][NBP];
    __shared__ double xj[NB][NBP];
    const double* a = a_all + (size_t)blockIdx.z * n * n;
    double* bx = bx_all + (size_t)blockIdx.z * n * nrhs;
    const int tx = threadIdx.x % NB, ty = threadIdx.x / NB;
    const int nb = min(NB, n - j0);
    const int col = blockIdx.x * NB + tx;
    const int i0 = backward ? blockIdx.y * NB : j0 + NB + blockIdx.y * NB;
    const int rows = backward ? min(NB, j0 - i0) : min(NB, n - i0);
    // l[ty][tx] = coefficient of x_j[tx] in row i0+ty
    if (!backward) l[ty][tx] = (ty < rows && tx < nb) ? a[(size_t)(i0 + ty) * n + j0 + tx] : 0.0;
    else l[ty][tx] = (ty < rows && tx < nb) ? a[(size_t)(j0 + tx) * n + i0 + ty] : 0.0;
    xj[ty][tx] = (ty < nb && col < nrhs) ? bx[(size_t)(j0 + ty) * nrhs + col] : 0.0;
    __syncthreads();
    if (ty < rows && col < nrhs) {
        double s = 0.0;
#pragma unroll 8
        for (int q = 0; q < NB; ++q) s = fma(l[ty][q], xj[q][tx], s);
        bx[(size_t)(i0 + ty) * nrhs + col] -= s;
    }
}

// ---- small matrices (n <= SMALL_MAX_N): one CTA factorises one matrix in a single launch ----------------
// Right-looking with 32-column panels.  The panel (all rows from the diagonal tile down x 32 columns) sits in
// shared memory TRANSPOSED (pT[c][r], rows contiguous, zero-padded to whole 64-row macro-tiles) so that every
// access of the three phases is conflict-free and the trailing update can use unguarded 16-byte loads:
//   diagonal tile   one warp, lane = row, the row in registers, column broadcasts by shuffle (no barriers);
//   rows below      one thread per row solves its 32 entries against the tile (reciprocal diagonal from the
//                   tile phase: no divisions);
//   trailing matrix (L2-resident, row-major) A -= P P^T in 64x64 macro-tiles, a 4x4 micro-tile per thread,
//                   four 16-byte shared loads per 16 DFMA.
// configs[4]: 1 024 frames x 270 DOF -- the kernel is instruction-issue bound (ncu, profiles/r1q_*).
constexpr int SMALL_MAX_N = 512;
constexpr int SMALL_THREADS = 256;

// padded panel rows: 32 + whole 64-row macro-tiles + 2 (16-byte aligned columns, stride == 2 mod 16); a template
// parameter so that every shared-memory address of the unrolled phases is an immediate offset
template <int MAXN> struct SmallLdp { static constexpr int value = NB + ((MAXN - NB + 63) / 64) * 64 + 2; };

template <int LDP>
__global__ void __launch_bounds__(SMALL_THREADS, 2) k_chol_small(double* __restrict__ a_all, int n, int* __restrict__ info,
                                                                 const uint8_t* __restrict__ active) {
    extern __shared__ __align__(16) double pT[];         // [NB][LDP] transposed panel, then rinv[NB]
    if (active != nullptr && !active[blockIdx.x]) return;
    double* rinv = pT + NB * LDP;
    double* a = a_all + (size_t)blockIdx.x * n * n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool even = (n & 1) == 0;
    __shared__ int s_bad;
    if (tid == 0) s_bad = 0;
    for (int j0 = 0; j0 < n; j0 += NB) {
        const int nb = min(NB, n - j0), rows = n - j0;
        const int used = NB + ((max(rows - NB, 0) + 63) / 64) * 64;      // rows the trailing update may read
        // panel -> shared (coalesced along the 32 columns of a row; padding rows and the tile's upper triangle = 0)
        for (int r = warp; r < used; r += SMALL_THREADS / 32) {
            double v = 0.0;
            if (r < rows && lane < nb && (r >= nb || lane <= r)) v = a[(size_t)(j0 + r) * n + j0 + lane];
            pT[lane * LDP + r] = v;
        }
        __syncthreads();
        if (warp == 0) {                                  // diagonal tile, left-looking, lane = row
            for (int k = 0; k < NB; ++k) {
                double s = pT[k * LDP + lane];
#pragma unroll 4
                for (int q = 0; q < k; ++q) s = fma(-pT[q * LDP + lane], pT[q * LDP + k], s);
                double skk = __shfl_sync(FULL, s, k);
                double rd = 0.0;
                if (k < nb) {
                    if (!(skk > 0.0)) {
                        if (lane == 0 && s_bad == 0) s_bad = j0 + k + 1;
                        skk = 1.0;
                    }
                    rd = rsqrt(skk);
                }
                pT[k * LDP + lane] = (lane == k) ? skk * rd : (lane > k ? s * rd : 0.0);
                if (lane == k) rinv[k] = rd;              // columns past nb: zero
                __syncwarp();
            }
        }
        __syncthreads();
        // rows below: x L_jj^T = p, one thread per row, the row in registers
        for (int r = nb + tid; r < rows; r += SMALL_THREADS) {
            double x[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) x[c] = pT[c * LDP + r];
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                double v = x[c];
#pragma unroll
                for (int q = 0; q < c; ++q) v = fma(-x[q], pT[q * LDP + c], v);
                x[c] = v * rinv[c];
            }
#pragma unroll
            for (int c = 0; c < NB; ++c) pT[c * LDP + r] = x[c];
        }
        __syncthreads();
        for (int r = warp; r < rows; r += SMALL_THREADS / 32)
            if (lane < nb && (r >= nb || lane <= r)) a[(size_t)(j0 + r) * n + j0 + lane] = pT[lane * LDP + r];
        // trailing update: rows / columns j0 + nb .. n - 1 (panel rows nb .. rows - 1), lower macro-tiles of 64 x 64;
        // the C micro-tile is fetched before the product so that its L2 latency hides behind the 512 DFMA
        const int rem = rows - nb;
        if (rem > 0) {
            const int tx = tid & 15, ty = tid >> 4;
            const int nt = (rem + 63) / 64;
            for (int ti = 0; ti < nt; ++ti)
                for (int tk = 0; tk <= ti; ++tk) {
                    const int i0 = nb + 64 * ti + 4 * ty, k0 = nb + 64 * tk + 4 * tx;     // panel-row indices
                    if (i0 >= rows || k0 >= rows || k0 > i0 + 3) continue;
                    const bool full = even && i0 + 3 < rows && k0 + 3 <= i0;             // whole 4x4 inside the triangle
                    double* cp = a + (size_t)(j0 + i0) * n + j0 + k0;
                    double2 c0[4][2];
                    if (full) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            c0[u][0] = *reinterpret_cast<const double2*>(cp + (size_t)u * n);
                            c0[u][1] = *reinterpret_cast<const double2*>(cp + (size_t)u * n + 2);
                        }
                    }
                    double acc[4][4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
#pragma unroll 8
                    for (int q = 0; q < NB; ++q) {
                        const double2 pa = *reinterpret_cast<const double2*>(pT + q * LDP + i0);
                        const double2 pb = *reinterpret_cast<const double2*>(pT + q * LDP + i0 + 2);
                        const double2 ka = *reinterpret_cast<const double2*>(pT + q * LDP + k0);
                        const double2 kb = *reinterpret_cast<const double2*>(pT + q * LDP + k0 + 2);
                        const double pi[4] = {pa.x, pa.y, pb.x, pb.y}, pk[4] = {ka.x, ka.y, kb.x, kb.y};
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int v = 0; v < 4; ++v) acc[u][v] = fma(pi[u], pk[v], acc[u][v]);
                    }
                    if (full) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            *reinterpret_cast<double2*>(cp + (size_t)u * n) = make_double2(c0[u][0].x - acc[u][0], c0[u][0].y - acc[u][1]);
                            *reinterpret_cast<double2*>(cp + (size_t)u * n + 2) = make_double2(c0[u][1].x - acc[u][2], c0[u][1].y - acc[u][3]);
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int v = 0; v < 4; ++v)
                                if (i0 + u < rows && k0 + v <= i0 + u) cp[(size_t)u * n + v] -= acc[u][v];
                    }
                }
        }
        __syncthreads();
    }
    if (tid == 0 && info != nullptr) info[blockIdx.x] = s_bad;
}

template <int MAXN>
static int launch_chol_small(double* a, int n, int batch, int* info_dev, cudaStream_t st, const uint8_t* active = nullptr) {
    constexpr int LDP = SmallLdp<MAXN>::value;
    constexpr size_t smem = ((size_t)NB * LDP + NB) * sizeof(double);
    static PerDeviceOnce attr_set;
    if (attr_set.needed()) {
        FERS_CK(cudaFuncSetAttribute(k_chol_small<LDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set.done();
    }
    k_chol_small<LDP><<<batch, SMALL_THREADS, smem, st>>>(a, n, info_dev, active);
    return 0;
}

// single right-hand side, small matrices: one CTA per matrix, blocked by 32 rows.  Per block the diagonal tile
// goes to shared memory and one warp substitutes through it (lane = row, shuffles, reciprocal diagonal); the other
// rows (forward) / columns (backward) are then updated by one thread each with 32 independent loads of L.
constexpr int SOLVE1_THREADS = 256;
__global__ void __launch_bounds__(SOLVE1_THREADS) k_solve_small1(const double* __restrict__ a_all, double* __restrict__ bx_all,
                                                                 int n, int batch, const uint8_t* __restrict__ active) {
    extern __shared__ double xs[];                       // [n]
    __shared__ double tile[NB][NBP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.x;
    if (f >= batch) return;
    if (active != nullptr && !active[f]) return;
    const double* a = a_all + (size_t)f * n * n;
    double* bx = bx_all + (size_t)f * n;
    for (int i = tid; i < n; i += SOLVE1_THREADS) xs[i] = bx[i];
    for (int j0 = 0; j0 < n; j0 += NB) {                  // forward: L y = b
        const int nb = min(NB, n - j0);
        for (int r = warp; r < nb; r += SOLVE1_THREADS / 32)
            if (lane < nb) tile[r][lane] = a[(size_t)(j0 + r) * n + j0 + lane];
        __syncthreads();
        if (warp == 0) {
            double xr = lane < nb ? xs[j0 + lane] : 0.0;
            const double rd = lane < nb ? 1.0 / tile[lane][lane] : 0.0;
            for (int k = 0; k < nb; ++k) {
                const double xk = __shfl_sync(FULL, xr * rd, k);
                if (lane == k) xr = xk;
                else if (lane > k && lane < nb) xr = fma(-tile[lane][k], xk, xr);
            }
            if (lane < nb) xs[j0 + lane] = xr;
        }
        __syncthreads();
        for (int r = j0 + nb + tid; r < n; r += SOLVE1_THREADS) {
            const double* row = a + (size_t)r * n + j0;
            double s = 0.0;
#pragma unroll 8
            for (int c = 0; c < NB; ++c) s = fma(row[c], xs[j0 + c], s);     // nb == NB whenever rows remain below
            xs[r] -= s;
        }
        __syncthreads();
    }
    for (int j0 = ((n - 1) / NB) * NB; j0 >= 0; j0 -= NB) {   // backward: L^T x = y
        const int nb = min(NB, n - j0);
        for (int r = warp; r < nb; r += SOLVE1_THREADS / 32)
            if (lane < nb) tile[r][lane] = a[(size_t)(j0 + r) * n + j0 + lane];
        __syncthreads();
        if (warp == 0) {
            double xr = lane < nb ? xs[j0 + lane] : 0.0;
            const double rd = lane < nb ? 1.0 / tile[lane][lane] : 0.0;
            for (int k = nb - 1; k >= 0; --k) {
                const double xk = __shfl_sync(FULL, xr * rd, k);
                if (lane == k) xr = xk;
                else if (lane < k) xr = fma(-tile[k][lane], xk, xr);
            }
            if (lane < nb) xs[j0 + lane] = xr;
        }
        __syncthreads();
        for (int c = tid; c < j0; c += SOLVE1_THREADS) {
            double s = 0.0;
            for (int i = 0; i < nb; ++i) s = fma(a[(size_t)(j0 + i) * n + c], xs[j0 + i], s);
            xs[c] -= s;
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += SOLVE1_THREADS) bx[i] = xs[i];
}

}  // namespace fers

using namespace fers;

static inline int rem_count(int n, int j0) { return (n - j0 - NB) > 0 ? 2 : 1; }

// One CTA per frame: relative change r = |un - u| / |un| of this P-Delta iteration.  An active frame takes the new
// displacements, counts the iteration and freezes once r < tol; frozen frames keep everything.  flags[0] |= some frame is
// still active, flags[1] |= some frame hit a non-positive pivot (bad[] remembers which).
__global__ void __launch_bounds__(128) k_frame_batch_step(int n, const double* __restrict__ un_all, double* __restrict__ u_all,
                                                          uint8_t* __restrict__ active, long long* __restrict__ iters,
                                                          double* __restrict__ resid, double tol, const int* __restrict__ info,
                                                          uint8_t* __restrict__ bad, int* __restrict__ flags) {
    const int f = blockIdx.x, tid = threadIdx.x;
    if (!active[f]) return;
    const double* un = un_all + (size_t)f * n;
    double* u = u_all + (size_t)f * n;
    double num = 0.0, den = 0.0;
    for (int i = tid; i < n; i += 128) {
        const double a = un[i], d = a - u[i];
        num += d * d; den += a * a;
    }
    __shared__ double s_num[4], s_den[4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { num += __shfl_xor_sync(0xffffffffu, num, o); den += __shfl_xor_sync(0xffffffffu, den, o); }
    if ((tid & 31) == 0) { s_num[tid >> 5] = num; s_den[tid >> 5] = den; }
    __syncthreads();
    num = (s_num[0] + s_num[1]) + (s_num[2] + s_num[3]);
    den = (s_den[0] + s_den[1]) + (s_den[2] + s_den[3]);
    for (int i = tid; i < n; i += 128) u[i] = un[i];
    if (tid == 0) {
        const double r = den > 0.0 ? sqrt(num) / sqrt(den) : 0.0;
        iters[f] += 1;
        resid[f] = r;
        const bool go_on = !(r < tol);
        active[f] = go_on ? 1 : 0;
        if (go_on) atomicOr(flags, 1);
        if (info[f] != 0) { bad[f] = 1; atomicOr(flags + 1, 1); }
    }
}

extern "C" {

int fers_b200_dense_expand(int n_nodes, const int* rowptr, const int* colidx, const double* vals,
                           const uint8_t* fixed, const double* kspring, double* a, void* stream) {
    cudaStream_t st = as_stream(stream);
    const size_t n = 6 * (size_t)n_nodes;
    FERS_CK(cudaMemsetAsync(a, 0, n * n * sizeof(double), st));
    const int T = 256;
    int grid = (n_nodes * 32 + T - 1) / T;
    k_dense_expand<<<grid, T, 0, st>>>(n_nodes, rowptr, colidx, vals, fixed, kspring, a);
    FERS_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int fers_b200_dense_expand_batched(int nodes_per_frame, int batch, const int* rowptr, const int* colidx,
                                   const double* vals, const uint8_t* fixed, const double* kspring, double* a,
                                   int* flag_dev, void* stream) {
    cudaStream_t st = as_stream(stream);
    FERS_REQUIRE(nodes_per_frame > 0 && batch > 0 && flag_dev != nullptr, "dense_expand_batched: bad arguments");
    const size_t n = 6 * (size_t)nodes_per_frame;
    FERS_CK(cudaMemsetAsync(a, 0, (size_t)batch * n * n * sizeof(double), st));
    FERS_CK(cudaMemsetAsync(flag_dev, 0, sizeof(int), st));
    const int T = 256;
    const long long rows = (long long)nodes_per_frame * batch;
    int grid = (int)((rows * 32 + T - 1) / T);
    k_dense_expand_batched<<<grid, T, 0, st>>>(nodes_per_frame, batch, rowptr, colidx, vals, fixed, kspring, a, flag_dev, nullptr);
    FERS_LAUNCH_CHECK();
    count_launches(1);
    int bad = 0;
    FERS_CK(cudaMemcpyAsync(&bad, flag_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    FERS_CK(cudaStreamSynchronize(st));
    FERS_REQUIRE(bad == 0, "dense_expand_batched: the model is not a batch of disconnected equal-sized frames");
    return 0;
}

int fers_b200_dense_cholesky(int n, int batch, double* a, int* info_dev, int* info_out, void* stream) {
    cudaStream_t st = as_stream(stream);
    FERS_REQUIRE(n > 0 && batch > 0, "dense_cholesky: empty problem");
    FERS_REQUIRE(batch <= 65535 || n <= SMALL_MAX_N, "dense_cholesky: batch exceeds gridDim.z");
    if (info_dev) FERS_CK(cudaMemsetAsync(info_dev, 0, sizeof(int) * batch, st));
    if (n <= SMALL_MAX_N) {                       // one CTA per matrix, one launch
        int rc;
        if (n <= 96) rc = launch_chol_small<96>(a, n, batch, info_dev, st);
        else if (n <= 160) rc = launch_chol_small<160>(a, n, batch, info_dev, st);
        else if (n <= 288) rc = launch_chol_small<288>(a, n, batch, info_dev, st);
        else rc = launch_chol_small<SMALL_MAX_N>(a, n, batch, info_dev, st);
        if (rc) return rc;
        count_launches(1);
    } else
    for (int j0 = 0; j0 < n; j0 += NB) {
        k_potrf_diag<<<dim3(1, 1, batch), NB * NB, 0, st>>>(a, n, j0, info_dev);
        count_launches(1);
        const int rem = n - j0 - NB;
        if (rem > 0) {
            const int nt = (rem + NB - 1) / NB;
            k_trsm_panel<<<dim3(nt, 1, batch), NB * NB, 0, st>>>(a, n, j0);
            k_syrk_update<<<dim3(nt, nt, batch), NB * NB, 0, st>>>(a, n, j0);
            count_launches(2);
        }
    }
    FERS_LAUNCH_CHECK();
    if (info_dev && info_out) {
        FERS_CK(cudaMemcpyAsync(info_out, info_dev, sizeof(int) * batch, cudaMemcpyDeviceToHost, st));
        FERS_CK(cudaStreamSynchronize(st));
    }
    return 0;
}

// ---- batch of frames in a P-Delta fixed point (configs[4]): masked, asynchronous ------------------------------------------
int fers_b200_frame_batch_factor_solve(int nodes_per_frame, int batch, const int* rowptr, const int* colidx, const double* vals,
                                       const uint8_t* fixed, const double* kspring, double* a, const uint8_t* active,
                                       int* info_dev, int* flag_dev, double* bx, void* stream) {
    cudaStream_t st = as_stream(stream);
    const int n = 6 * nodes_per_frame;
    FERS_REQUIRE(nodes_per_frame > 0 && batch > 0 && a && info_dev && flag_dev && bx, "frame_batch_factor_solve: bad arguments");
    FERS_REQUIRE(n <= SMALL_MAX_N, "frame_batch_factor_solve: frames of at most 512 DOF (one CTA per frame)");
    FERS_CK(cudaMemsetAsync(a, 0, (size_t)batch * n * n * sizeof(double), st));
    FERS_CK(cudaMemsetAsync(info_dev, 0, sizeof(int) * batch, st));
    const int T = 256;
    const long long rows = (long long)nodes_per_frame * batch;
    k_dense_expand_batched<<<(int)((rows * 32 + T - 1) / T), T, 0, st>>>(nodes_per_frame, batch, rowptr, colidx, vals, fixed, kspring,
                                                                         a, flag_dev, active);
    int rc;
    if (n <= 96) rc = launch_chol_small<96>(a, n, batch, info_dev, st, active);
    else if (n <= 160) rc = launch_chol_small<160>(a, n, batch, info_dev, st, active);
    else if (n <= 288) rc = launch_chol_small<288>(a, n, batch, info_dev, st, active);
    else rc = launch_chol_small<SMALL_MAX_N>(a, n, batch, info_dev, st, active);
    if (rc) return rc;
    k_solve_small1<<<batch, SOLVE1_THREADS, (size_t)n * sizeof(double), st>>>(a, bx, n, batch, active);
    FERS_LAUNCH_CHECK();
    count_launches(3);
    return 0;
}

int fers_b200_frame_batch_step(int batch, int n, const double* un, double* u, uint8_t* active, long long* iters, double* resid,
                               double tol, const int* info_dev, uint8_t* bad, int* flags, void* stream) {
    FERS_REQUIRE(batch > 0 && n > 0 && un && u && active && iters && resid && info_dev && bad && flags, "frame_batch_step: bad arguments");
    k_frame_batch_step<<<batch, 128, 0, as_stream(stream)>>>(n, un, u, active, iters, resid, tol, info_dev, bad, flags);
    FERS_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int fers_b200_dense_solve(int n, int nrhs, int batch, const double* a, double* bx, void* stream) {
    cudaStream_t st = as_stream(stream);
    FERS_REQUIRE(n > 0 && nrhs > 0 && batch > 0, "dense_solve: empty problem");
    if (nrhs == 1 && n <= 4096) {                 // one CTA per matrix (n doubles of shared memory <= 32 KB)
        k_solve_small1<<<batch, SOLVE1_THREADS, (size_t)n * sizeof(double), st>>>(a, bx, n, batch, nullptr);
        FERS_LAUNCH_CHECK();
        count_launches(1);
        return 0;
    }
    const int nc = (nrhs + NB - 1) / NB;
    for (int j0 = 0; j0 < n; j0 += NB) {
        k_trsv_diag<<<dim3(nc, 1, batch), NB * NB, 0, st>>>(a, bx, n, nrhs, j0, 0);
        count_launches(rem_count(n, j0));
        const int rem = n - j0 - NB;
        if (rem > 0) k_subst_update<<<dim3(nc, (rem + NB - 1) / NB, batch), NB * NB, 0, st>>>(a, bx, n, nrhs, j0, 0);
    }
    const int last = ((n - 1) / NB) * NB;
    for (int j0 = last; j0 >= 0; j0 -= NB) {
        k_trsv_diag<<<dim3(nc, 1, batch), NB * NB, 0, st>>>(a, bx, n, nrhs, j0, 1);
        count_launches(j0 > 0 ? 2 : 1);
        if (j0 > 0) k_subst_update<<<dim3(nc, j0 / NB, batch), NB * NB, 0, st>>>(a, bx, n, nrhs, j0, 1);
    }
    FERS_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
