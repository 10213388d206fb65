// chol_gemm.cuh -- fp64 tile GEMM core of the multifrontal Cholesky (C -= A * op(B)).
//
// All matrices are column-major.  A is M x K with rows contiguous; B is either N x K with rows
// contiguous ("NT": C -= A B^T, the form of every factorisation update and of the forward
// substitution) or K x N with K contiguous ("NN": the backward substitution, which multiplies by L^T).
// One CTA of 256 threads computes a (16 TM) x (16 TN) tile; a thread owns TM x TN accumulators laid
// out as 2-wide row / column pairs every 32 rows / columns, so that
//   - the A fragment of a warp is 16 consecutive 16-byte words (conflict-free LDS.128, the two
//     column groups of the warp read the same words = broadcast),
//   - a C column segment written by a warp is 256 contiguous bytes.
// Operand tiles move global -> shared with cp.async (16 B, zero-fill outside the matrix) through a
// STAGES-deep ring; the fp64 pipe is the intended limiter (64 DFMA per 8 shared loads per thread).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fers {

constexpr int GEMM_THREADS = 256;
constexpr int GEMM_BK = 16;
constexpr int GEMM_STAGES = 3;
constexpr int GEMM_BKP = GEMM_BK + 2;   // padded K extent of the K-major B tile

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int TM, int TN, bool B_KMAJOR>
struct GemmTile {
    static constexpr int BM = 16 * TM, BN = 16 * TN;
    static constexpr int MIN_CTAS = 1;
    static constexpr int A_STAGE = GEMM_BK * BM;                                   // doubles
    static constexpr int B_STAGE = B_KMAJOR ? BN * GEMM_BKP : GEMM_BK * BN;        // doubles
    static constexpr size_t SMEM = (size_t)GEMM_STAGES * (A_STAGE + B_STAGE) * sizeof(double);
    struct Acc { double v[TM][TN]; };
    __device__ static __forceinline__ void clear(Acc& acc) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc.v[i][j] = 0.0;
    }

    // stage loader: k-tile kt (columns kt*BK .. ) of A (mrem x K, lda) and B into ring slot `slot`
    __device__ static __forceinline__ void load_stage(double* smem, int slot, int kt, const double* __restrict__ A, int lda,
                                                      int mrem, const double* __restrict__ B, int ldb, int nrem, int K) {
        const int tid = threadIdx.x;
        const int k0 = kt * GEMM_BK;
        double* as = smem + (size_t)slot * (A_STAGE + B_STAGE);
        double* bs = as + A_STAGE;
        const uint32_t as_addr = (uint32_t)__cvta_generic_to_shared(as);
        const uint32_t bs_addr = (uint32_t)__cvta_generic_to_shared(bs);
        constexpr int A_CHUNKS = GEMM_BK * BM / 2;
#pragma unroll
        for (int c = tid; c < A_CHUNKS; c += GEMM_THREADS) {
            const int k = c / (BM / 2), i = 2 * (c % (BM / 2));
            const bool ok = (i < mrem) && (k0 + k < K);
            const double* src = ok ? A + (size_t)(k0 + k) * lda + i : A;
            cp_async16(as_addr + (uint32_t)(k * BM + i) * 8u, src, ok);
        }
        if (!B_KMAJOR) {
            constexpr int B_CHUNKS = GEMM_BK * BN / 2;
#pragma unroll
            for (int c = tid; c < B_CHUNKS; c += GEMM_THREADS) {
                const int k = c / (BN / 2), j = 2 * (c % (BN / 2));
                const bool ok = (j < nrem) && (k0 + k < K);
                const double* src = ok ? B + (size_t)(k0 + k) * ldb + j : B;
                cp_async16(bs_addr + (uint32_t)(k * BN + j) * 8u, src, ok);
            }
        } else {
            constexpr int B_CHUNKS = BN * GEMM_BK / 2;
#pragma unroll
            for (int c = tid; c < B_CHUNKS; c += GEMM_THREADS) {
                const int j = c / (GEMM_BK / 2), k = 2 * (c % (GEMM_BK / 2));
                const bool ok = (j < nrem) && (k0 + k < K);
                const double* src = ok ? B + (size_t)j * ldb + k0 + k : B;
                cp_async16(bs_addr + (uint32_t)(j * GEMM_BKP + k) * 8u, src, ok);
            }
        }
    }

    // acc += A(mrem x K) * op(B)(K x nrem) for this CTA's tile (pointers at the tile origin)
    template <bool CLEAR = true>
    __device__ static __forceinline__ void mainloop(Acc& acc_, double* smem, const double* __restrict__ A, int lda,
                                                    int mrem, const double* __restrict__ B, int ldb, int nrem, int K) {
        const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
        double (&acc)[TM][TN] = acc_.v;
        if (CLEAR) {
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = 0.0;
        }
        const int nk = (K + GEMM_BK - 1) / GEMM_BK;
#pragma unroll
        for (int s = 0; s < GEMM_STAGES - 1; ++s) {
            if (s < nk) load_stage(smem, s, s, A, lda, mrem, B, ldb, nrem, K);
            cp_async_commit();
        }
        for (int kt = 0; kt < nk; ++kt) {
            cp_async_wait<GEMM_STAGES - 2>();
            __syncthreads();
            {
                const int nxt = kt + GEMM_STAGES - 1;
                if (nxt < nk) load_stage(smem, nxt % GEMM_STAGES, nxt, A, lda, mrem, B, ldb, nrem, K);
                cp_async_commit();
            }
            const double* as = smem + (size_t)(kt % GEMM_STAGES) * (A_STAGE + B_STAGE);
            const double* bs = as + A_STAGE;
#pragma unroll
            for (int kk = 0; kk < GEMM_BK; ++kk) {
                double a[TM], b[TN];
#pragma unroll
                for (int q = 0; q < TM / 2; ++q) {
                    const double2 v = *reinterpret_cast<const double2*>(as + kk * BM + 32 * q + 2 * tx);
                    a[2 * q] = v.x; a[2 * q + 1] = v.y;
                }
                if (!B_KMAJOR) {
#pragma unroll
                    for (int q = 0; q < TN / 2; ++q) {
                        const double2 v = *reinterpret_cast<const double2*>(bs + kk * BN + 32 * q + 2 * ty);
                        b[2 * q] = v.x; b[2 * q + 1] = v.y;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < TN / 2; ++q) {
                        b[2 * q] = bs[(32 * q + 2 * ty) * GEMM_BKP + kk];
                        b[2 * q + 1] = bs[(32 * q + 2 * ty + 1) * GEMM_BKP + kk];
                    }
                }
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
            }
        }
        cp_async_wait<0>();
        __syncthreads();
    }

    // C = C - acc (SET == false) or C = acc (SET == true) on the valid part of the tile
    template <bool SET>
    __device__ static __forceinline__ void store(const Acc& acc_, double* __restrict__ C, size_t ldc, int mrem,
                                                 int nrem) {
        const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
        const double (&acc)[TM][TN] = acc_.v;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int col = 32 * (j / 2) + 2 * ty + (j & 1);
            if (col >= nrem) continue;
#pragma unroll
            for (int q = 0; q < TM / 2; ++q) {
                const int row = 32 * q + 2 * tx;
                if (row >= mrem) continue;
                double2* p = reinterpret_cast<double2*>(C + (size_t)col * ldc + row);
                double2 v;
                if (SET) {
                    v.x = acc[2 * q][j]; v.y = acc[2 * q + 1][j];
                } else {
                    v = *p;
                    v.x -= acc[2 * q][j]; v.y -= acc[2 * q + 1][j];
                }
                *p = v;
            }
        }
    }
};

// ---- DMMA variant -------------------------------------------------------------------------------
// Same interface, products on the fp64 tensor pipe (mma.sync m8n8k4 = SASS DMMA.8x8x4).  Measured on
// B200 (tools/micro/fp64_peak.cu): DFMA and DMMA both peak at 37.0 TFLOP/s, but one DMMA replaces
// eight DFMA per lane and a 64 x 32 warp tile needs 12 shared loads per 32 DMMA instead of 8 per 64
// DFMA, so the shared-memory pipe and the issue slots stop being co-limiters (ncu r1d: the DFMA
// kernel kept the fp64 pipe 65 % busy).  The MMA "A" operand is fed from the B tile and the MMA "B"
// operand from the A tile: a lane then owns two consecutive ROWS of C per 8x8 block and the epilogue
// uses 16-byte accesses along the contiguous dimension of the column-major C.
// Shared tiles are padded (+4 doubles per k row, 20 doubles per column of the K-major B tile) so that
// the fragment loads of a half-warp (4 k x 4 consecutive elements) hit 32 distinct banks.
__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1])
                 : "d"(a), "d"(b));
}

template <int BM_, int BN_, int WARPS_M, int WARPS_N, bool B_KMAJOR, int BK_ = 32, int STAGES_ = 3, int MIN_CTAS_ = 1>
struct GemmMma {
    static constexpr int MIN_CTAS = MIN_CTAS_;                           // CTAs per SM the register budget is set for
    static_assert(WARPS_M * WARPS_N * 32 == GEMM_THREADS, "8 warps per CTA");
    static constexpr int BM = BM_, BN = BN_, BK = BK_, STAGES = STAGES_;
    static constexpr int MT = BM / WARPS_M / 8, NT = BN / WARPS_N / 8;   // 8x8 blocks per warp
    static constexpr int BKP = BK + 4;                                   // padded K extent of the K-major B tile
    static constexpr int LDA = BM + 4;
    static constexpr int LDB = B_KMAJOR ? BKP : BN + 4;
    static constexpr int A_STAGE = BK * LDA;
    static constexpr int B_STAGE = B_KMAJOR ? BN * BKP : BK * LDB;
    static constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double);
    static constexpr int A_CHUNKS = BK * BM / 2 / GEMM_THREADS;          // 16-byte chunks per thread and stage
    static constexpr int B_CHUNKS = BK * BN / 2 / GEMM_THREADS;
    static constexpr int KSTEPS = BK / 4;
    static_assert((BK * BM / 2) % GEMM_THREADS == 0 && (BK * BN / 2) % GEMM_THREADS == 0, "whole chunks per thread");
    struct Acc { double v[NT][MT][2]; };
    __device__ static __forceinline__ void clear(Acc& acc) {
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int i = 0; i < MT; ++i) acc.v[j][i][0] = acc.v[j][i][1] = 0.0;
    }

    // chunk `q` (0 .. A_CHUNKS + B_CHUNKS - 1) of this thread for k-tile kt -> ring slot
    __device__ static __forceinline__ void load_chunk(int q, uint32_t stage_addr, int k0, const double* __restrict__ A, int lda,
                                                      int mrem, const double* __restrict__ B, int ldb, int nrem, int K) {
        const int tid = threadIdx.x;
        if (q < A_CHUNKS) {
            const int c = tid + q * GEMM_THREADS;
            const int k = c / (BM / 2), i = 2 * (c % (BM / 2));
            const bool ok = (i < mrem) && (k0 + k < K);
            const double* src = ok ? A + (size_t)(k0 + k) * lda + i : A;
            cp_async16(stage_addr + (uint32_t)(k * LDA + i) * 8u, src, ok);
        } else {
            const int c = tid + (q - A_CHUNKS) * GEMM_THREADS;
            const uint32_t bs_addr = stage_addr + A_STAGE * 8u;
            if (!B_KMAJOR) {
                const int k = c / (BN / 2), j = 2 * (c % (BN / 2));
                const bool ok = (j < nrem) && (k0 + k < K);
                const double* src = ok ? B + (size_t)(k0 + k) * ldb + j : B;
                cp_async16(bs_addr + (uint32_t)(k * LDB + j) * 8u, src, ok);
            } else {
                const int j = c / (BK / 2), k = 2 * (c % (BK / 2));
                const bool ok = (j < nrem) && (k0 + k < K);
                const double* src = ok ? B + (size_t)j * ldb + k0 + k : B;
                cp_async16(bs_addr + (uint32_t)(j * BKP + k) * 8u, src, ok);
            }
        }
    }

    // CLEAR = false continues an accumulation started by an earlier call (same tile, next K range)
    template <bool CLEAR = true>
    __device__ static __forceinline__ void mainloop(Acc& acc, double* smem, const double* __restrict__ A, int lda, int mrem,
                                                    const double* __restrict__ B, int ldb, int nrem, int K) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int wm0 = (warp % WARPS_M) * (MT * 8), wn0 = (warp / WARPS_M) * (NT * 8);
        const int lk = lane & 3, lr = lane >> 2;
        if (CLEAR) {
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int i = 0; i < MT; ++i) acc.v[j][i][0] = acc.v[j][i][1] = 0.0;
        }
        const int nk = (K + BK - 1) / BK;
        const uint32_t smem_addr = (uint32_t)__cvta_generic_to_shared(smem);
        constexpr uint32_t STAGE_BYTES = (A_STAGE + B_STAGE) * 8u;
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            if (s < nk) {
#pragma unroll
                for (int q = 0; q < A_CHUNKS + B_CHUNKS; ++q)
                    load_chunk(q, smem_addr + s * STAGE_BYTES, s * BK, A, lda, mrem, B, ldb, nrem, K);
            }
            cp_async_commit();
        }
        const int a_off = lk * LDA + wm0 + lr;
        const int b_off = B_KMAJOR ? (wn0 + lr) * BKP + lk : lk * LDB + wn0 + lr;
        int slot = 0, nslot = STAGES - 1;       // ring slot being computed / being filled
        for (int kt = 0; kt < nk; ++kt) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            const int nxt = kt + STAGES - 1;
            const bool do_load = nxt < nk;
            const uint32_t fill_addr = smem_addr + nslot * STAGE_BYTES;
            const double* as = smem + (size_t)slot * (A_STAGE + B_STAGE) + a_off;
            const double* bs = smem + (size_t)slot * (A_STAGE + B_STAGE) + A_STAGE + b_off;
            double am[2][MT], bn[2][NT];
#pragma unroll
            for (int i = 0; i < MT; ++i) am[0][i] = as[8 * i];
#pragma unroll
            for (int j = 0; j < NT; ++j) bn[0][j] = B_KMAJOR ? bs[8 * j * BKP] : bs[8 * j];
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const int cur = ks & 1;
                if (ks + 1 < KSTEPS) {          // fragments of the next k4 step
                    const int k4 = 4 * (ks + 1);
#pragma unroll
                    for (int i = 0; i < MT; ++i) am[cur ^ 1][i] = as[k4 * LDA + 8 * i];
#pragma unroll
                    for (int j = 0; j < NT; ++j) bn[cur ^ 1][j] = B_KMAJOR ? bs[8 * j * BKP + k4] : bs[k4 * LDB + 8 * j];
                }
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int i = 0; i < MT; ++i) dmma884(acc.v[j][i], bn[cur][j], am[cur][i]);
                // this step's share of the next stage's copies, issued while the tensor pipe drains the DMMAs
                if (do_load) {
                    constexpr int TOTAL = A_CHUNKS + B_CHUNKS;
#pragma unroll
                    for (int q = ks * TOTAL / KSTEPS; q < (ks + 1) * TOTAL / KSTEPS; ++q)
                        load_chunk(q, fill_addr, nxt * BK, A, lda, mrem, B, ldb, nrem, K);
                }
            }
            cp_async_commit();
            slot = (slot + 1 == STAGES) ? 0 : slot + 1;
            nslot = (nslot + 1 == STAGES) ? 0 : nslot + 1;
        }
        cp_async_wait<0>();
        __syncthreads();
    }

    // ---- helpers of the wavefront substitution (square 64 x 64 tiles: BM == BN == K) ----------------------
    static constexpr int SQ_LD = BN + 4;                     // padded row of a resident 64 x 64 operand tile
    static constexpr size_t SQ_BYTES = (size_t)BN * SQ_LD * sizeof(double);

    // acc = -C (tile origin C, column stride ldc; entries outside mrem x nrem are zero)
    __device__ static __forceinline__ void load_neg(Acc& acc, const double* __restrict__ C, size_t ldc, int mrem, int nrem) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int wm0 = (warp % WARPS_M) * (MT * 8), wn0 = (warp / WARPS_M) * (NT * 8);
        const int lk = lane & 3, lr = lane >> 2;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int col = wn0 + 8 * j + lr;
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const int row = wm0 + 8 * i + 2 * lk;
                double2 v = make_double2(0.0, 0.0);
                if (col < nrem && row < mrem) v = __ldcg(reinterpret_cast<const double2*>(C + (size_t)col * ldc + row));
                acc.v[j][i][0] = -v.x; acc.v[j][i][1] = -v.y;
            }
        }
    }

    // the accumulator tile as the A operand of a following product: as[n * SQ_LD + m] (n = tile column = K index)
    __device__ static __forceinline__ void acc_to_operand(const Acc& acc, double* __restrict__ as) {
        static_assert(BM == BN, "square tiles only");
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int wm0 = (warp % WARPS_M) * (MT * 8), wn0 = (warp / WARPS_M) * (NT * 8);
        const int lk = lane & 3, lr = lane >> 2;
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int i = 0; i < MT; ++i)
                *reinterpret_cast<double2*>(as + (wn0 + 8 * j + lr) * SQ_LD + wm0 + 8 * i + 2 * lk) =
                    make_double2(acc.v[j][i][0], acc.v[j][i][1]);
    }

    // a 64 x 64 tile (column stride ld in global memory) -> resident operand rows of SQ_LD doubles (one commit group)
    __device__ static __forceinline__ void load_square(double* __restrict__ bs, const double* __restrict__ B, int ld) {
        const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bs);
#pragma unroll
        for (int q = 0; q < BN * BN / 2 / GEMM_THREADS; ++q) {
            const int c = threadIdx.x + q * GEMM_THREADS;
            const int k = c / (BN / 2), j = 2 * (c % (BN / 2));
            cp_async16(addr + (uint32_t)(k * SQ_LD + j) * 8u, B + (size_t)k * ld + j, true);
        }
        cp_async_commit();
    }

    // out[n][m] = sum_k as[k][m] * (BT ? bs[n][k] : bs[k][n]),  k = 0 .. BN - 1, both operands resident
    template <bool BT>
    __device__ static __forceinline__ void square_product(Acc& out, const double* __restrict__ as, const double* __restrict__ bs) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int wm0 = (warp % WARPS_M) * (MT * 8), wn0 = (warp / WARPS_M) * (NT * 8);
        const int lk = lane & 3, lr = lane >> 2;
        clear(out);
        const double* ap = as + lk * SQ_LD + wm0 + lr;
        const double* bp = BT ? bs + (wn0 + lr) * SQ_LD + lk : bs + lk * SQ_LD + wn0 + lr;
#pragma unroll 4
        for (int ks = 0; ks < BN / 4; ++ks) {
            double am[MT], bn[NT];
#pragma unroll
            for (int i = 0; i < MT; ++i) am[i] = ap[4 * ks * SQ_LD + 8 * i];
#pragma unroll
            for (int j = 0; j < NT; ++j) bn[j] = BT ? bp[8 * j * SQ_LD + 4 * ks] : bp[4 * ks * SQ_LD + 8 * j];
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int i = 0; i < MT; ++i) dmma884(out.v[j][i], bn[j], am[i]);
        }
    }

    template <bool SET>
    __device__ static __forceinline__ void store(const Acc& acc, double* __restrict__ C, size_t ldc, int mrem, int nrem) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int wm0 = (warp % WARPS_M) * (MT * 8), wn0 = (warp / WARPS_M) * (NT * 8);
        const int lk = lane & 3, lr = lane >> 2;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int col = wn0 + 8 * j + lr;
            if (col >= nrem) continue;
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const int row = wm0 + 8 * i + 2 * lk;
                if (row >= mrem) continue;
                double2* p = reinterpret_cast<double2*>(C + (size_t)col * ldc + row);
                double2 v;
                if (SET) {
                    v.x = acc.v[j][i][0]; v.y = acc.v[j][i][1];
                } else {
                    v = *p;
                    v.x -= acc.v[j][i][0]; v.y -= acc.v[j][i][1];
                }
                *p = v;
            }
        }
    }
};

}  // namespace fers
