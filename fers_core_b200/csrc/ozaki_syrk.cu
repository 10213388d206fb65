// ozaki_syrk.cu -- fp64 Schur update F22 -= L21 L21^T of the multifrontal Cholesky on the 5th-generation tensor cores
// (tcgen05.mma kind::i8, int32 accumulators in TMEM) by integer slicing of the fp64 operand (Ozaki scheme I).
//
// tcgen05 has no fp64 kind; the fp64 tensor path of sm_100a is mma.sync (DMMA), 37 TFLOP/s.  The int8 path runs at more
// than a hundred times that rate, so a product of fp64 matrices split into S signed 7-bit digit planes, with the S(S+1)/2
// significant digit-plane products formed exactly in int32, is an fp64-equivalent GEMM whose accuracy is set by S:
//
//   row i of A = L21:   a_ik = 2^E_i * sum_{p<S} q_p[i,k] 2^(-6-Bp),   2^E_i > max_k |a_ik|,   B = 8 bits per digit plane:
//                       q_0 in [-65, 65], q_p in [-128, 127] (signed digits, carries normalised)
//   (A A^T)_ij = 2^(E_i+E_j) * sum_{d<S} 2^(-12-Bd) * ( sum_{p+q=d} Q_p Q_q^T )_ij  +  O(S 2^(2-BS)) |a_i||a_j|
//
// Every bracket is an int8 x int8 -> int32 GEMM chain accumulated in ONE TMEM accumulator per digit-plane diagonal d
// (exact: |q_p q_q| <= 2^14, (d+1) K 2^14 < 2^31 for K < 16 384 with S <= 8).  S = 7 carries 6 + 6*8 = 54 bits (fp64 grade),
// S = 6 46 bits.  (7-bit digits, the textbook choice, need S = 8 planes and 36 instead of 28 products for the same bits.)
//
// Kernels:
//   k_oz_rowmax     max_k |a_ik| per row of L21 (E_i = its exponent)
//   k_oz_split      fp64 -> S int8 digit planes, written in the shared-memory image the tensor cores read (K-major core
//                   matrices, 8 rows x 16 B), so that ONE bulk copy (cp.async.bulk, TMA engine) per plane and K-chunk
//                   brings a whole operand tile; no tensor maps, any front shape
//   k_oz_syrk       persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane,
//                   tcgen05.mma.cta_group::1.kind::i8, 128 x 64 x 32 per instruction), warps 2-5 = epilogue
//                   (tcgen05.ld, int32 -> fp64, scaling, C -= ...).  Ring of shared-memory stages with full / empty
//                   mbarriers, tcgen05.commit releases a stage when the MMAs that read it have retired.
//
// Reference row: SURVEY.md 8a A10 (the factorisation the reference describes at fers_core/types/pydantic_models.py:1216-1219).
#include <stdlib.h>

#include <algorithm>

#include "chol.cuh"
#include "common.cuh"

namespace fers {

constexpr int OZ_M = 128;                 // rows of an output tile (TMEM lanes)
constexpr int OZ_N = 64;                  // columns of an output tile (TMEM columns per accumulator)
constexpr int OZ_KC = 32;                 // K bytes per shared-memory stage (one MMA of K = 32 per digit-plane pair): small stages, deep ring
constexpr int OZ_DIGIT_BITS = 8;          // bits per digit plane after the first (6 bits)
constexpr double OZ_BASE = 256.0;         // 2^OZ_DIGIT_BITS
constexpr int OZ_THREADS = 192;
constexpr int OZ_BAND = 12;               // row blocks per band of the tile order (L2 reuse of the operand tiles)
constexpr int OZ_WIDE_DEFAULT = 1;
constexpr int OZ_TS_DEFAULT = 0;          // A operand through tensor memory unless FERS_OZ_TS says otherwise
constexpr uint32_t OZ_SPIN_LIMIT = 1u << 28;   // a barrier that never completes traps instead of hanging the device

__host__ __device__ inline size_t oz_stage_bytes(int S) { return (size_t)S * (OZ_M + OZ_N) * OZ_KC; }
__host__ __device__ inline int oz_stages(int S) { return (int)((200 * 1024) / oz_stage_bytes(S)); }

// digit planes of one front: plane p at p * plane_bytes; inside a plane [k chunk][8-row group][16-byte K piece][8 rows][16 B]
struct OzFront {
    const double* A;        // L21: m2 x k1, column-major, leading dimension lda
    double* C;              // F22: m2 x m2, column-major, leading dimension ldc (lower triangle)
    int8_t* planes;
    double* rowscale;       // [m2_pad] 2^E_i
    double* rowmax;         // [m2_pad] max_k |a_ik| (scratch of k_oz_rowmax)
    long long lda, ldc;
    int m2, k1;
    int m2_pad, k1_pad;     // multiples of OZ_M / OZ_KC
    int ncols;              // only columns < ncols of C are updated (ncols = m2: the whole lower triangle)
};

__host__ __device__ inline size_t oz_plane_bytes(int m2_pad, int k1_pad) { return (size_t)m2_pad * k1_pad; }

// ---- row exponents ---------------------------------------------------------------------------------------------------------
// row maxima by a 2D grid (rows x slices of K; one thread per row alone is a chain of k1 dependent fmax over 50 CTAs:
// 17.6 ms per factorisation at configs[3]); non-negative doubles order like their bit patterns, so the slices meet in an
// integer atomicMax.  k_oz_split turns the maximum into 2^E_i.
__global__ void __launch_bounds__(256)
k_oz_rowmax(OzFront f, int k_per_slice) {
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= f.m2) return;
    const int k0 = blockIdx.y * k_per_slice, k1 = min(k0 + k_per_slice, f.k1);
    double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;
    int k = k0;
    for (; k + 3 < k1; k += 4) {
        m0 = fmax(m0, fabs(f.A[(size_t)k * f.lda + r]));
        m1 = fmax(m1, fabs(f.A[(size_t)(k + 1) * f.lda + r]));
        m2 = fmax(m2, fabs(f.A[(size_t)(k + 2) * f.lda + r]));
        m3 = fmax(m3, fabs(f.A[(size_t)(k + 3) * f.lda + r]));
    }
    for (; k < k1; ++k) m0 = fmax(m0, fabs(f.A[(size_t)k * f.lda + r]));
    const double mx = fmax(fmax(m0, m1), fmax(m2, m3));
    if (mx > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(f.rowmax + r), (unsigned long long)__double_as_longlong(mx));
}
__device__ __forceinline__ double oz_rowscale_of(double mx) {
    int e = 0;
    if (mx > 0.0) { frexp(mx, &e); }           // mx = m 2^e, 0.5 <= m < 1  ->  |a| < 2^e
    return ldexp(1.0, e);
}

// ---- fp64 -> digit planes ---------------------------------------------------------------------------------------------------
// thread = (row, 16-byte K piece): reads 16 doubles of its row (coalesced across the rows of a warp), writes 16 B per plane
template <int S>
__global__ void __launch_bounds__(256)
k_oz_split(OzFront f) {
    const int r = blockIdx.x * 256 + threadIdx.x;
    const int piece = blockIdx.y;                       // 16 K values
    if (r >= f.m2_pad) return;
    const int k0 = piece * 16;
    const double rs = (r < f.m2) ? oz_rowscale_of(f.rowmax[r]) : 1.0;
    if (piece == 0) f.rowscale[r] = rs;
    const double inv = (r < f.m2) ? 64.0 / rs : 0.0;
    uint32_t w[S][4];
#pragma unroll
    for (int p = 0; p < S; ++p) w[p][0] = w[p][1] = w[p][2] = w[p][3] = 0u;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int k = k0 + j;
        double t = 0.0;
        if (r < f.m2 && k < f.k1) t = f.A[(size_t)k * f.lda + r] * inv;       // |t| < 64, exact scaling
        int q[S];
#pragma unroll
        for (int p = 0; p < S; ++p) {
            const double qd = rint(t);
            t = (t - qd) * OZ_BASE;                                           // exact: |t - qd| <= 1/2
            q[p] = (int)qd;                                                   // p > 0: in [-128, 128]
        }
#pragma unroll
        for (int p = S - 1; p > 0; --p)                                       // signed digits in [-128, 127]: +128 carries upwards
            if (q[p] >= 128) { q[p] -= 256; q[p - 1] += 1; }
#pragma unroll
        for (int p = 0; p < S; ++p) w[p][j >> 2] |= ((uint32_t)q[p] & 0xffu) << (8 * (j & 3));
    }
    const size_t plane = oz_plane_bytes(f.m2_pad, f.k1_pad);
    const int chunk = k0 / OZ_KC, kc16 = (k0 % OZ_KC) / 16;
    const size_t off = (size_t)chunk * f.m2_pad * OZ_KC + (size_t)(r >> 3) * (OZ_KC * 8) + (size_t)kc16 * 128 + (size_t)(r & 7) * 16;
#pragma unroll
    for (int p = 0; p < S; ++p)
        *reinterpret_cast<uint4*>(f.planes + p * plane + off) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
}

// ---- PTX helpers ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t oz_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void oz_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void oz_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > OZ_SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ void oz_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void oz_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void oz_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
// one lane of a CONVERGED warp; the code under it may use warp-uniform operands without a divergence loop
__device__ __forceinline__ bool oz_elect() {
    uint32_t pred;
    __syncwarp();           // the lanes may leave a barrier-wait loop at different iterations: reconverge before electing
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void oz_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void oz_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void oz_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 B (128 contiguous bytes);
// LBO = distance of the two core matrices an MMA spans along K, SBO = distance of consecutive 8-row groups
// (shared-memory addresses are below 256 KB, so address >> 4 fits the 14-bit field and tile offsets can simply be added to it)
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr16, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint32_t lo = saddr16 | ((lbo_bytes >> 4) << 16);
    const uint32_t hi = (sbo_bytes >> 4) | (1u << 14);      // descriptor version of sm_100 at bit 46; base offset 0, no swizzle
    return ((uint64_t)hi << 32) | lo;
}
// instruction descriptor of kind::i8: D = S32, A = B = signed 8 bit, both K-major, M = 128, N = 64
constexpr uint32_t OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_N >> 3) << 17) | ((uint32_t)(OZ_M >> 4) << 24);
__device__ __forceinline__ void oz_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(OZ_IDESC), "r"(accumulate)
        : "memory");
}
// A operand from tensor memory (128 lanes x K / 4 columns): the tile is copied there once per digit plane and K chunk
// (tcgen05.cp, 128 rows x 256 bits) and read by every product of that plane, instead of S - p reads of shared memory
__device__ __forceinline__ void oz_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(OZ_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void oz_cp_a(uint32_t tmem_a, uint64_t da) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_a), "l"(da) : "memory");
}
__device__ __forceinline__ void oz_tmem_ld8(uint32_t taddr, int (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void oz_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Tile t of the lower triangle in band order: bands of OZ_BAND row blocks; inside a band column block by column block, so that
// the ~148 tiles in flight share about a dozen row-block and a dozen column-block operand streams in L2.  Column block bj
// exists in row block r when bj <= 2 r + 1.
__device__ __forceinline__ bool oz_tile(int t, int nbi, int nbj, int& bi, int& bj) {
    for (int b0 = 0; b0 < nbi; b0 += OZ_BAND) {
        const int b1 = min(b0 + OZ_BAND, nbi);
        const int maxbj = min(2 * (b1 - 1) + 1, nbj - 1);
        for (int j = 0; j <= maxbj; ++j) {
            const int rmin = max(b0, j / 2);                    // smallest r with 2 r + 1 >= j
            const int cnt = b1 - rmin;
            if (cnt <= 0) continue;
            if (t < cnt) { bi = rmin + t; bj = j; return true; }
            t -= cnt;
        }
    }
    return false;
}

// (Thread-block clusters with TMA multicast of the row-block operand were measured and dropped: 2 CTAs per cluster changed
// nothing -- L2 already serves the dozen CTAs that stream the same tile -- and 4 per cluster strand SMs: half the throughput.)
// TS: the A tile of a digit plane goes through tensor memory (columns S * OZ_N ..: 8 per plane; S <= 7).  tcgen05.cp and
// tcgen05.mma of one thread execute in issue order, so the copies of the next K chunk wait for the products of this one.
template <int S, bool TS>
__global__ void __launch_bounds__(OZ_THREADS, 1)
k_oz_syrk(OzFront f, int n_tiles) {
    static_assert(!TS || (S * OZ_N + S * (OZ_KC / 4) <= 512 && OZ_KC == 32), "A staging does not fit the tensor memory");
    extern __shared__ __align__(1024) uint8_t oz_smem_raw[];
    constexpr int STAGES = (200 * 1024) / (S * (OZ_M + OZ_N) * OZ_KC);
    constexpr uint32_t STAGE_BYTES = S * (OZ_M + OZ_N) * OZ_KC;
    constexpr uint32_t A_BYTES = OZ_M * OZ_KC, B_BYTES = OZ_N * OZ_KC;
    static_assert(STAGES >= 2, "ring too shallow");
    __shared__ __align__(8) uint64_t bars[2 * STAGES + 2];
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ring = oz_smem(oz_smem_raw);
    const uint32_t bar_full = oz_smem(&bars[0]), bar_empty = oz_smem(&bars[STAGES]);
    const uint32_t bar_acc_full = oz_smem(&bars[2 * STAGES]), bar_acc_empty = oz_smem(&bars[2 * STAGES + 1]);
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { oz_mbar_init(bar_full + 8u * i, 1); oz_mbar_init(bar_empty + 8u * i, 1); }
        oz_mbar_init(bar_acc_full, 1);
        oz_mbar_init(bar_acc_empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                    // the MMA warp owns the tensor memory: all 512 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    oz_fence_before();
    __syncthreads();
    oz_fence_after();
    const uint32_t tmem = s_tmem;
    const int unit0 = blockIdx.x, unit_stride = gridDim.x;
    const int nbi = f.m2_pad / OZ_M, nbj = (f.ncols + OZ_N - 1) / OZ_N;
    const int nkc = f.k1_pad / OZ_KC;
    const size_t plane = oz_plane_bytes(f.m2_pad, f.k1_pad);
    const size_t chunk_bytes = (size_t)f.m2_pad * OZ_KC;

    if (warp == 0) {
        // ===== TMA producer (the whole warp walks the loop, one elected lane issues) =====
        uint32_t it = 0;
        for (int t = unit0; t < n_tiles; t += unit_stride) {
            int bi, bj;
            if (!oz_tile(t, nbi, nbj, bi, bj)) break;
            const int8_t* a0 = f.planes + (size_t)bi * OZ_M * OZ_KC;
            const int8_t* b0 = f.planes + (size_t)bj * OZ_N * OZ_KC;
            for (int c = 0; c < nkc; ++c, ++it) {
                const uint32_t st = it % STAGES, ph = (it / STAGES) & 1u;
                oz_mbar_wait(bar_empty + 8u * st, ph ^ 1u);
                if (oz_elect()) {
                    const uint32_t full = bar_full + 8u * st;
                    oz_mbar_expect_tx(full, STAGE_BYTES);
                    const uint32_t dst = ring + st * STAGE_BYTES;
#pragma unroll
                    for (int p = 0; p < S; ++p) {
                        oz_bulk_g2s(dst + p * A_BYTES, a0 + p * plane + (size_t)c * chunk_bytes, A_BYTES, full);
                        oz_bulk_g2s(dst + S * A_BYTES + p * B_BYTES, b0 + p * plane + (size_t)c * chunk_bytes, B_BYTES, full);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp in the loop, one elected lane issues the MMAs of a stage) =====
        uint32_t it = 0, tile_i = 0;
        for (int t = unit0; t < n_tiles; t += unit_stride, ++tile_i) {
            int bi, bj;
            if (!oz_tile(t, nbi, nbj, bi, bj)) break;
            oz_mbar_wait(bar_acc_empty, (tile_i & 1u) ^ 1u);
            oz_fence_after();
            for (int c = 0; c < nkc; ++c, ++it) {
                const uint32_t st = it % STAGES, ph = (it / STAGES) & 1u;
                oz_mbar_wait(bar_full + 8u * st, ph);
                oz_fence_after();
                if (oz_elect()) {
                    const uint32_t sa = (ring + st * STAGE_BYTES) >> 4, sb = sa + ((S * A_BYTES) >> 4);
                    if (TS) {
#pragma unroll
                        for (int p = 0; p < S; ++p)
                            oz_cp_a(tmem + (uint32_t)(S * OZ_N + 8 * p), oz_desc(sa + ((p * A_BYTES) >> 4), 128, OZ_KC * 8));
#pragma unroll
                        for (int p = 0; p < S; ++p)
#pragma unroll
                            for (int q = 0; q + p < S; ++q) {
                                const uint64_t db = oz_desc(sb + ((q * B_BYTES) >> 4), 128, OZ_KC * 8);
                                oz_mma_ts(tmem + (uint32_t)(p + q) * OZ_N, tmem + (uint32_t)(S * OZ_N + 8 * p), db, (c > 0 || p > 0) ? 1u : 0u);
                            }
                    } else
#pragma unroll
                    for (int kk = 0; kk < OZ_KC / 32; ++kk) {
#pragma unroll
                        for (int p = 0; p < S; ++p) {
                            const uint64_t da = oz_desc(sa + ((p * A_BYTES + kk * 256) >> 4), 128, OZ_KC * 8);
#pragma unroll
                            for (int q = 0; q + p < S; ++q) {
                                const uint64_t db = oz_desc(sb + ((q * B_BYTES + kk * 256) >> 4), 128, OZ_KC * 8);
                                // the first product of a digit-plane diagonal d = p + q overwrites its accumulator
                                const uint32_t acc = (c > 0 || kk > 0 || p > 0) ? 1u : 0u;
                                oz_mma(tmem + (uint32_t)(p + q) * OZ_N, da, db, acc);
                            }
                        }
                    }
                    oz_commit(bar_empty + 8u * st);                  // the stage is free once these MMAs have read it
                    if (c == nkc - 1) oz_commit(bar_acc_full);       // all accumulators of the tile are final
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue: TMEM -> fp64 -> C -= 2^(E_i + E_j) sum_d 2^(-12 - 7 d) acc_d =====
        const int quad = warp & 3;                           // the TMEM lanes this warp may read: 32 quad .. 32 quad + 31
        uint32_t tile_i = 0;
        for (int t = unit0; t < n_tiles; t += unit_stride, ++tile_i) {
            int bi, bj;
            if (!oz_tile(t, nbi, nbj, bi, bj)) break;
            oz_mbar_wait(bar_acc_full, tile_i & 1u);
            oz_fence_after();
            const int r = bi * OZ_M + 32 * quad + lane;
            const int c0 = bj * OZ_N;
            const double rs = f.rowscale[min(r, f.m2_pad - 1)];
            const uint32_t lane_addr = tmem + ((uint32_t)(32 * quad) << 16);
#pragma unroll 1
            for (int cb = 0; cb < OZ_N; cb += 8) {
                double sum[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) sum[j] = 0.0;
                int v[S][8];
#pragma unroll
                for (int d = 0; d < S; ++d) oz_tmem_ld8(lane_addr + (uint32_t)(d * OZ_N + cb), v[d]);
                oz_tmem_wait_ld();
                double wgt = ldexp(1.0, -12 - OZ_DIGIT_BITS * (S - 1));
#pragma unroll
                for (int d = S - 1; d >= 0; --d) {           // smallest weights first
#pragma unroll
                    for (int j = 0; j < 8; ++j) sum[j] = fma((double)v[d][j], wgt, sum[j]);
                    wgt *= OZ_BASE;
                }
                if (r < f.m2) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int col = c0 + cb + j;
                        if (col <= r && col < f.ncols) {
                            double* dst = f.C + (size_t)col * f.ldc + r;
                            *dst -= sum[j] * rs * f.rowscale[col];
                        }
                    }
                }
            }
            oz_fence_before();
            __syncwarp();
            if (lane == 0) oz_mbar_arrive(bar_acc_empty);
        }
    }
    oz_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// ---- wide tiles: 128 x 128 per CTA, two passes over K -----------------------------------------------------------------------
// Measured on B200 (tools/micro/mma_rate.cu): a tcgen05.mma with both operands in shared memory takes
// max(N / 2, (4096 + 32 N) / 128) cycles at M = 128, K = 32 bytes -- the tensor core reads its operands at 128 B per cycle.
// N = 64 costs 48 cycles (2/3 of the int8 rate), N = 128 costs 64 (full rate).  S accumulators of 128 columns do not fit the
// 512 columns of tensor memory, so a tile is done in two passes over K: first the low-order diagonals d >= 4 (all S planes,
// S - 4 accumulators), then d < 4 (planes 0..3, 4 accumulators); each pass ends with its own C -= ... (small terms first).
// The operand bytes per product are the same as with 128 x 64 tiles (11 plane loads of 256 rows against 7 of 192 for half
// the tile), the products run twice as wide.
constexpr int OZW_N = 128;
constexpr int OZW_EPI_WARPS = 8;
constexpr int OZW_THREADS = 64 + 32 * OZW_EPI_WARPS;
constexpr uint32_t OZW_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZW_N >> 3) << 17) | ((uint32_t)(OZ_M >> 4) << 24);
__device__ __forceinline__ void oz_mma_w(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(OZW_IDESC), "r"(accumulate)
        : "memory");
}
// tile t of the lower triangle of 128 x 128 blocks, bands of OZ_BAND row blocks, column by column inside a band
__device__ __forceinline__ bool ozw_tile(int t, int nb, int nbj, int& bi, int& bj) {
    for (int b0 = 0; b0 < nb; b0 += OZ_BAND) {
        const int b1 = min(b0 + OZ_BAND, nb);
        const int maxbj = min(b1 - 1, nbj - 1);
        for (int j = 0; j <= maxbj; ++j) {
            const int rmin = max(b0, j);
            const int cnt = b1 - rmin;
            if (t < cnt) { bi = rmin + t; bj = j; return true; }
            t -= cnt;
        }
    }
    return false;
}

// DS: diagonals d >= DS in the first pass (all S planes), d < DS in the second (planes 0 .. DS - 1)
template <int S, int DS>
__global__ void __launch_bounds__(OZW_THREADS, 1)
k_oz_syrk_wide(OzFront f, int n_tiles) {
    extern __shared__ __align__(1024) uint8_t oz_smem_raw[];
    constexpr uint32_t T_BYTES = OZ_M * OZ_KC;                      // one 128-row tile of one plane and K chunk
    constexpr uint32_t STAGE_BYTES = 2 * S * T_BYTES;               // A planes, then B planes
    constexpr int STAGES = (200 * 1024) / STAGE_BYTES;
    constexpr int NPASS = S > DS ? 2 : 1;
    static_assert(STAGES >= 2 && OZ_KC == 32 && (S - DS) * OZW_N <= 512 && DS * OZW_N <= 512, "wide int8 tile geometry");
    __shared__ __align__(8) uint64_t bars[2 * STAGES + 2];
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ring = oz_smem(oz_smem_raw);
    const uint32_t bar_full = oz_smem(&bars[0]), bar_empty = oz_smem(&bars[STAGES]);
    const uint32_t bar_acc_full = oz_smem(&bars[2 * STAGES]), bar_acc_empty = oz_smem(&bars[2 * STAGES + 1]);
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { oz_mbar_init(bar_full + 8u * i, 1); oz_mbar_init(bar_empty + 8u * i, 1); }
        oz_mbar_init(bar_acc_full, 1);
        oz_mbar_init(bar_acc_empty, OZW_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    oz_fence_before();
    __syncthreads();
    oz_fence_after();
    const uint32_t tmem = s_tmem;
    const int unit0 = blockIdx.x, unit_stride = gridDim.x;
    const int nb = f.m2_pad / OZ_M, nbj = (f.ncols + OZW_N - 1) / OZW_N;
    const int nkc = f.k1_pad / OZ_KC;
    const size_t plane = oz_plane_bytes(f.m2_pad, f.k1_pad);
    const size_t chunk_bytes = (size_t)f.m2_pad * OZ_KC;

    if (warp == 0) {
        // ===== TMA producer =====
        uint32_t it = 0;
        for (int t = unit0; t < n_tiles; t += unit_stride) {
            int bi, bj;
            if (!ozw_tile(t, nb, nbj, bi, bj)) break;
            const int8_t* a0 = f.planes + (size_t)bi * T_BYTES;
            const int8_t* b0 = f.planes + (size_t)bj * T_BYTES;
            for (int pass = 0; pass < NPASS; ++pass) {
                const int np = (NPASS == 2 && pass == 0) ? S : (S < DS ? S : DS);
                for (int c = 0; c < nkc; ++c, ++it) {
                    const uint32_t st = it % STAGES, ph = (it / STAGES) & 1u;
                    oz_mbar_wait(bar_empty + 8u * st, ph ^ 1u);
                    if (oz_elect()) {
                        const uint32_t full = bar_full + 8u * st;
                        oz_mbar_expect_tx(full, 2u * np * T_BYTES);
                        const uint32_t dst = ring + st * STAGE_BYTES;
                        for (int p = 0; p < np; ++p) {
                            oz_bulk_g2s(dst + p * T_BYTES, a0 + p * plane + (size_t)c * chunk_bytes, T_BYTES, full);
                            oz_bulk_g2s(dst + (S + p) * T_BYTES, b0 + p * plane + (size_t)c * chunk_bytes, T_BYTES, full);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        uint32_t it = 0, acc_i = 0;
        for (int t = unit0; t < n_tiles; t += unit_stride) {
            int bi, bj;
            if (!ozw_tile(t, nb, nbj, bi, bj)) break;
            for (int pass = 0; pass < NPASS; ++pass, ++acc_i) {
                const bool low = (NPASS == 2 && pass == 0);         // diagonals d >= DS
                oz_mbar_wait(bar_acc_empty, (acc_i & 1u) ^ 1u);
                oz_fence_after();
                for (int c = 0; c < nkc; ++c, ++it) {
                    const uint32_t st = it % STAGES, ph = (it / STAGES) & 1u;
                    oz_mbar_wait(bar_full + 8u * st, ph);
                    oz_fence_after();
                    if (oz_elect()) {
                        const uint32_t sa = (ring + st * STAGE_BYTES) >> 4, sb = sa + ((S * T_BYTES) >> 4);
                        if (low) {
#pragma unroll
                            for (int p = 0; p < S; ++p) {
                                const uint64_t da = oz_desc(sa + ((p * T_BYTES) >> 4), 128, OZ_KC * 8);
#pragma unroll
                                for (int q = 0; q + p < S; ++q) {
                                    if (p + q < DS) continue;
                                    const uint64_t db = oz_desc(sb + ((q * T_BYTES) >> 4), 128, OZ_KC * 8);
                                    oz_mma_w(tmem + (uint32_t)(p + q - DS) * OZW_N, da, db, (c > 0 || p > 0) ? 1u : 0u);
                                }
                            }
                        } else {
#pragma unroll
                            for (int p = 0; p < S && p < DS; ++p) {
                                const uint64_t da = oz_desc(sa + ((p * T_BYTES) >> 4), 128, OZ_KC * 8);
#pragma unroll
                                for (int q = 0; q + p < S && q + p < DS; ++q) {
                                    const uint64_t db = oz_desc(sb + ((q * T_BYTES) >> 4), 128, OZ_KC * 8);
                                    oz_mma_w(tmem + (uint32_t)(p + q) * OZW_N, da, db, (c > 0 || p > 0) ? 1u : 0u);
                                }
                            }
                        }
                        oz_commit(bar_empty + 8u * st);
                        if (c == nkc - 1) oz_commit(bar_acc_full);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===== epilogue: 8 warps, warp w reads TMEM lanes 32 (w % 4) .., columns 64 (w / 4) .. of every accumulator =====
        const int ew = warp - 2, quad = warp & 3, half = ew >> 2;
        uint32_t acc_i = 0;
        for (int t = unit0; t < n_tiles; t += unit_stride) {
            int bi, bj;
            if (!ozw_tile(t, nb, nbj, bi, bj)) break;
            const int r = bi * OZ_M + 32 * quad + lane;
            const int c0 = bj * OZW_N + 64 * half;
            const double rs = f.rowscale[min(r, f.m2_pad - 1)];
            const uint32_t lane_addr = tmem + ((uint32_t)(32 * quad) << 16) + (uint32_t)(64 * half);
            // the C tile is read-modify-written from DRAM (a front's F22 is gigabytes): ask for it in L2 now, while the
            // products of this tile run, so that the epilogue -- which the next pass waits for -- finds it there.  A warp's
            // 32 rows of one column are one 256-byte run: every fourth lane asks for its 32-byte sector.
            if (r < f.m2 && (lane & 3) == 0) {
#pragma unroll 4
                for (int j = 0; j < 64; ++j) {
                    const int col = c0 + j;
                    if (col <= r + 3 && col < f.ncols)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(f.C + (size_t)col * f.ldc + r));
                }
            }
            for (int pass = 0; pass < NPASS; ++pass, ++acc_i) {
                const bool low = (NPASS == 2 && pass == 0);
                oz_mbar_wait(bar_acc_full, acc_i & 1u);
                oz_fence_after();
                if (low) {
                    constexpr int ND = S > DS ? S - DS : 1;
#pragma unroll 1
                    for (int cb = 0; cb < 64; cb += 8) {
                        double sum[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) sum[j] = 0.0;
                        int v[ND][8];
#pragma unroll
                        for (int d = 0; d < ND; ++d) oz_tmem_ld8(lane_addr + (uint32_t)(d * OZW_N + cb), v[d]);
                        oz_tmem_wait_ld();
                        double wgt = ldexp(1.0, -12 - OZ_DIGIT_BITS * (S - 1));
#pragma unroll
                        for (int d = ND - 1; d >= 0; --d) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) sum[j] = fma((double)v[d][j], wgt, sum[j]);
                            wgt *= OZ_BASE;
                        }
                        if (r < f.m2) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int col = c0 + cb + j;
                                if (col <= r && col < f.ncols) {
                                    double* dst = f.C + (size_t)col * f.ldc + r;
                                    *dst -= sum[j] * rs * f.rowscale[col];
                                }
                            }
                        }
                    }
                } else {
                    constexpr int ND = S < DS ? S : DS;
#pragma unroll 1
                    for (int cb = 0; cb < 64; cb += 8) {
                        double sum[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) sum[j] = 0.0;
                        int v[ND][8];
#pragma unroll
                        for (int d = 0; d < ND; ++d) oz_tmem_ld8(lane_addr + (uint32_t)(d * OZW_N + cb), v[d]);
                        oz_tmem_wait_ld();
                        double wgt = ldexp(1.0, -12 - OZ_DIGIT_BITS * (ND - 1));
#pragma unroll
                        for (int d = ND - 1; d >= 0; --d) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) sum[j] = fma((double)v[d][j], wgt, sum[j]);
                            wgt *= OZ_BASE;
                        }
                        if (r < f.m2) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int col = c0 + cb + j;
                                if (col <= r && col < f.ncols) {
                                    double* dst = f.C + (size_t)col * f.ldc + r;
                                    *dst -= sum[j] * rs * f.rowscale[col];
                                }
                            }
                        }
                    }
                }
                oz_fence_before();
                __syncwarp();
                if (lane == 0) oz_mbar_arrive(bar_acc_empty);
            }
        }
    }
    oz_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

static int ozw_tile_count(int ncols, int m2_pad) {
    const int nb = m2_pad / OZ_M, nbj = (ncols + OZW_N - 1) / OZW_N;
    long long n = 0;
    for (int r = 0; r < nb; ++r) n += std::min(r + 1, nbj);
    return (int)n;
}

static int oz_tile_count(int ncols, int m2_pad) {
    const int nbi = m2_pad / OZ_M, nbj = (ncols + OZ_N - 1) / OZ_N;
    long long n = 0;
    for (int r = 0; r < nbi; ++r) n += std::min(2 * r + 2, nbj);
    return (int)n;
}

size_t oz_work_bytes(int m2, int k1, int S) {
    const int m2_pad = (m2 + OZ_M - 1) / OZ_M * OZ_M, k1_pad = (k1 + OZ_KC - 1) / OZ_KC * OZ_KC;
    return align_up((size_t)S * oz_plane_bytes(m2_pad, k1_pad)) + 2 * align_up((size_t)m2_pad * sizeof(double)) + 256;
}

static int g_oz_dsplit = 3;
static int g_oz_wide = -1;    // 128 x 128 tiles in two passes over K (S <= 7); FERS_OZ_WIDE=0 keeps the 128 x 64 single-pass kernel
static int g_oz_ts = -1;      // A operand through tensor memory (S <= 7); FERS_OZ_TS=0 keeps both operands in shared memory

template <int S>
static int oz_launch(OzFront f, cudaStream_t st, bool persistent) {
    static PerDeviceOnce once;
    constexpr bool TS_OK = (S <= 7);
    const size_t smem = (size_t)oz_stages(S) * oz_stage_bytes(S);
    if (g_oz_ts < 0) {
        const char* e = getenv("FERS_OZ_TS");
        g_oz_ts = e ? atoi(e) : OZ_TS_DEFAULT;
    }
    if (once.needed()) {
        FERS_CK(cudaFuncSetAttribute(k_oz_syrk<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FERS_CK(cudaFuncSetAttribute(k_oz_syrk<S, TS_OK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        once.done();
    }
    FERS_CK(cudaMemsetAsync(f.rowmax, 0, (size_t)f.m2_pad * sizeof(double), st));
    {
        const int row_blocks = (f.m2_pad + 255) / 256;
        const int slices = std::max(1, std::min((f.k1 + 63) / 64, (4 * sm_count() + row_blocks - 1) / row_blocks));
        const int k_per_slice = (f.k1 + slices - 1) / slices;
        k_oz_rowmax<<<dim3(row_blocks, (f.k1 + k_per_slice - 1) / k_per_slice), 256, 0, st>>>(f, k_per_slice);
    }
    dim3 grid((f.m2_pad + 255) / 256, f.k1_pad / 16);
    k_oz_split<S><<<grid, 256, 0, st>>>(f);
    if (g_oz_wide < 0) {
        const char* e = getenv("FERS_OZ_WIDE");
        g_oz_wide = e ? atoi(e) : OZ_WIDE_DEFAULT;
        const char* d = getenv("FERS_OZ_DSPLIT");
        g_oz_dsplit = d ? atoi(d) : 3;
    }
    if (g_oz_wide) {
        static PerDeviceOnce once_w;
        constexpr size_t smem_w = (size_t)((200 * 1024) / (2 * S * OZ_M * OZ_KC)) * (2 * S * OZ_M * OZ_KC);
        constexpr int DS_ALT = (S - 3) * OZW_N <= 512 ? 3 : 4;       // split at d = 3 (default): one plane load less per K chunk than at d = 4
        if (once_w.needed()) {
            FERS_CK(cudaFuncSetAttribute(k_oz_syrk_wide<S, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
            FERS_CK(cudaFuncSetAttribute(k_oz_syrk_wide<S, DS_ALT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
            once_w.done();
        }
        const int n_w = ozw_tile_count(f.ncols, f.m2_pad);
        const int n_cta_w = persistent ? std::min(n_w, sm_count()) : n_w;
        if (g_oz_dsplit == 3) k_oz_syrk_wide<S, DS_ALT><<<n_cta_w, OZW_THREADS, smem_w, st>>>(f, n_w);
        else k_oz_syrk_wide<S, 4><<<n_cta_w, OZW_THREADS, smem_w, st>>>(f, n_w);
        FERS_LAUNCH_CHECK();
        count_launches(3);
        return 0;
    }
    const int n_tiles = oz_tile_count(f.ncols, f.m2_pad);
    // persistent: one CTA per SM walks the tile list (long K: the Schur updates).  Otherwise one CTA per tile, so that SMs
    // come free every few microseconds for the kernels of a higher-priority stream (the panel chain of the look-ahead)
    const int n_cta = persistent ? std::min(n_tiles, sm_count()) : n_tiles;
    if (TS_OK && g_oz_ts) k_oz_syrk<S, TS_OK><<<n_cta, OZ_THREADS, smem, st>>>(f, n_tiles);
    else k_oz_syrk<S, false><<<n_cta, OZ_THREADS, smem, st>>>(f, n_tiles);
    FERS_LAUNCH_CHECK();
    count_launches(3);
    return 0;
}

// C (m2 x m2, lower) -= A A^T with A m2 x k1, both column-major fp64; `work` holds the digit planes and row scales
int oz_syrk(const double* A, long long lda, double* C, long long ldc, int m2, int k1, int S, void* work, cudaStream_t st,
            int ncols, bool persistent) {
    OzFront f;
    f.A = A; f.C = C; f.lda = lda; f.ldc = ldc; f.m2 = m2; f.k1 = k1;
    f.ncols = (ncols <= 0 || ncols > m2) ? m2 : ncols;
    f.m2_pad = (m2 + OZ_M - 1) / OZ_M * OZ_M;
    f.k1_pad = (k1 + OZ_KC - 1) / OZ_KC * OZ_KC;
    f.planes = static_cast<int8_t*>(work);
    f.rowscale = reinterpret_cast<double*>(static_cast<uint8_t*>(work) + align_up((size_t)S * oz_plane_bytes(f.m2_pad, f.k1_pad)));
    f.rowmax = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(f.rowscale) + align_up((size_t)f.m2_pad * sizeof(double)));
    switch (S) {
        case 4: return oz_launch<4>(f, st, persistent);
        case 5: return oz_launch<5>(f, st, persistent);
        case 6: return oz_launch<6>(f, st, persistent);
        case 7: return oz_launch<7>(f, st, persistent);
        case 8: return oz_launch<8>(f, st, persistent);
    }
    set_error("oz_syrk: 4 <= slices <= 8");
    return 2;
}

}  // namespace fers

extern "C" {

size_t fers_b200_oz_syrk_work_bytes(int m2, int k1, int slices) { return fers::oz_work_bytes(m2, k1, slices); }



int fers_b200_oz_syrk(const double* A, long long lda, double* C, long long ldc, int m2, int k1, int slices, void* work,
                      size_t work_bytes, void* stream) {
    FERS_REQUIRE(A && C && work && m2 > 0 && k1 > 0 && lda >= m2 && ldc >= m2, "oz_syrk: bad arguments");
    FERS_REQUIRE(k1 < 16384, "oz_syrk: K must stay below 16 384 for exact int32 accumulation");
    FERS_REQUIRE(work_bytes >= fers::oz_work_bytes(m2, k1, slices), "oz_syrk: workspace too small");
    return fers::oz_syrk(A, lda, C, ldc, m2, k1, slices, work, fers::as_stream(stream), m2, true);
}

}  // extern "C"
