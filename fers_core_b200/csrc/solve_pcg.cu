// solve_pcg.cu -- block-CSR (6x6) SpMM / SpMV, block-Jacobi preconditioner and the batched
// multi-right-hand-side preconditioned CG (the reference engine's
// "batch first-order solve ... across many load cases", pydantic_models.py:1216-1219).
//
// Layout: multi-vectors are [n_dof][NR] with the right-hand-side index fastest, so a warp
// reading one DOF row touches NR*8 contiguous bytes.  Thread mapping of every multi-vector
// kernel: a group of CP = NR/2 lanes owns one node (six DOF rows), each lane two adjacent
// columns (one 16-byte access per row); G = 32/CP node groups share a warp.  A 6x6 matrix block
// is read through the read-only path as a broadcast (all lanes of a group load the same 16
// bytes), six x rows of the neighbour node are read once per block, 72 DFMA per lane follow.
// HBM-bound by design: 0.23 GB of matrix + 2 * n_dof * NR * 8 bytes of vectors per product
// (DESIGN.md "SpMM").  Reductions are deterministic: per-CTA partials + ordered final sum.
#include <stdlib.h>

#include "common.cuh"

namespace fers {

constexpr int SPMM_THREADS = 256;
constexpr int SPMM_WARPS = SPMM_THREADS / 32;
constexpr int CTAS_PER_SM = 3;

__device__ __forceinline__ void fma2(double2& a, double k, const double2& x) {
    a.x = fma(k, x.x, a.x);
    a.y = fma(k, x.y, a.y);
}

// NS partial sums per column are reduced over the node groups of a warp, then over the warps of
// the CTA, and written to partial[blockIdx.x][NS][NR].
template <int NR, int NS>
__device__ __forceinline__ void cta_reduce_store(double2 (&v)[NS], double* __restrict__ partial) {
    constexpr int CP = NR / 2;
    __shared__ double2 s_red[SPMM_WARPS][NS][CP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
#pragma unroll
        for (int off = CP; off < 32; off <<= 1) {
            v[s].x += __shfl_xor_sync(FULL, v[s].x, off);
            v[s].y += __shfl_xor_sync(FULL, v[s].y, off);
        }
        if (lane < CP) s_red[warp][s][lane] = v[s];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < NS * CP; t += SPMM_THREADS) {
        const int s = t / CP, cp = t % CP;
        double2 a = s_red[0][s][cp];
#pragma unroll
        for (int w = 1; w < SPMM_WARPS; ++w) { a.x += s_red[w][s][cp].x; a.y += s_red[w][s][cp].y; }
        reinterpret_cast<double2*>(partial + ((size_t)blockIdx.x * NS + s) * NR)[cp] = a;
    }
}

template <int NR, bool DOT>
__global__ void __launch_bounds__(SPMM_THREADS, CTAS_PER_SM)
k_spmm(int n_nodes, const int* __restrict__ rowptr, const int* __restrict__ colidx, const double* __restrict__ vals,
       const uint8_t* __restrict__ fixed, const double* __restrict__ kspring, const double* __restrict__ x,
       double* __restrict__ y, double* __restrict__ partial) {
    constexpr int CP = NR / 2, G = 32 / CP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / CP, cp = lane % CP;
    double2 dot[1];
    dot[0] = make_double2(0.0, 0.0);
    // interleaved row groups: at any instant all CTAs work inside one narrow window of consecutive
    // rows, so the x rows of neighbouring nodes are shared through L2 (ncu r1: chunked mapping gave
    // 8 % L2 hit rate and 2.4x the algorithmic DRAM traffic)
    for (int row = (blockIdx.x * SPMM_WARPS + warp) * G + g; row < n_nodes; row += gridDim.x * SPMM_WARPS * G) {
        double2 acc[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) acc[r] = make_double2(0.0, 0.0);
        const int kb = rowptr[row], ke = rowptr[row + 1];
        for (int k = kb; k < ke; ++k) {
            const int col = __ldg(colidx + k);
            const double* xb = x + (size_t)col * 6 * NR + 2 * cp;
            double2 xv[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) xv[c] = ldg2(xb + (size_t)c * NR);
            const double* blk = vals + (size_t)k * 36;
#pragma unroll
            for (int r = 0; r < 6; ++r) {
#pragma unroll
                for (int c = 0; c < 6; c += 2) {
                    const double2 kk = ldg2(blk + r * 6 + c);
                    fma2(acc[r], kk.x, xv[c]);
                    fma2(acc[r], kk.y, xv[c + 1]);
                }
            }
        }
        const size_t d0 = (size_t)row * 6;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const size_t d = d0 + r;
            double2 xo = make_double2(0.0, 0.0);
            if (kspring != nullptr || DOT) xo = ldg2(x + d * NR + 2 * cp);
            if (fixed != nullptr && fixed[d]) {
                acc[r] = make_double2(0.0, 0.0);
            } else if (kspring != nullptr) {
                const double ks = kspring[d];
                acc[r].x = fma(ks, xo.x, acc[r].x);
                acc[r].y = fma(ks, xo.y, acc[r].y);
            }
            reinterpret_cast<double2*>(y + d * NR)[cp] = acc[r];
            if (DOT) { dot[0].x = fma(xo.x, acc[r].x, dot[0].x); dot[0].y = fma(xo.y, acc[r].y, dot[0].y); }
        }
    }
    if (DOT) cta_reduce_store<NR, 1>(dot, partial);
}

// ---- TMA-staged variant for NR = 64 ------------------------------------------------------------
// Each warp owns a ring of STAGES shared-memory slots; one slot holds one 6x6 matrix block (288 B)
// and the six x rows of its column node (6*64*8 = 3072 contiguous bytes).  Lane 0 issues two bulk
// copies (cp.async.bulk, completion on an mbarrier) per block, STAGES blocks ahead of the FMA loop,
// so the DFMA pipe never waits on an L2/HBM round trip; the FMA loop reads 16 B per lane per x row
// (conflict-free) and the matrix block as shared-memory broadcasts.
constexpr int TMA_STAGES = 3;
constexpr int TMA_KB = 36 * 8;
constexpr int TMA_XB = 6 * 64 * 8;
constexpr int TMA_SLOT = TMA_KB + TMA_XB + 16;       // + 16-byte header holding the block's column node
constexpr size_t TMA_SMEM = (size_t)SPMM_WARPS * TMA_STAGES * TMA_SLOT + (size_t)SPMM_WARPS * TMA_STAGES * 8;

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// 16-byte shared-memory load at base + compile-time offset (keeps the address math out of the loop)
template <int OFFSET>
__device__ __forceinline__ void lds_f64x2(double2& v, uint32_t base) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(base), "n"(OFFSET));
}

// No global load sits on the critical path of the FMA loop: row pointers and the BC mask of the NEXT
// row are fetched while the current row is processed, the producer lane keeps the next column index
// prefetched, and the row's own x (needed for the fused p.Ap and the support springs) is taken from
// the shared-memory slot of its diagonal block.
template <bool DOT>
__global__ void __launch_bounds__(SPMM_THREADS, 2)
k_spmm64_tma(int n_nodes, const int* __restrict__ rowptr, const int* __restrict__ colidx, const double* __restrict__ vals,
             const uint8_t* __restrict__ fixed, const double* __restrict__ kspring, const double* __restrict__ x,
             double* __restrict__ y, double* __restrict__ partial) {
    constexpr int NR = 64;
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* ring = smem_dyn + (size_t)warp * TMA_STAGES * TMA_SLOT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_dyn + (size_t)SPMM_WARPS * TMA_STAGES * TMA_SLOT) + warp * TMA_STAGES;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < TMA_STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_addr(&bars[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int first = blockIdx.x * SPMM_WARPS + warp, stride = gridDim.x * SPMM_WARPS;

    // ---- producer cursor (lane 0): block pk of row prow, its column prefetched in pcol; the
    //      pointers of the following row prefetched in (nkb, nke)
    int prow = first, pk = 0, pke = 0, pcol = 0, nkb = 0, nke = 0;
    if (lane == 0 && prow < n_nodes) {
        pk = rowptr[prow]; pke = rowptr[prow + 1];
        pcol = __ldg(colidx + pk);
        if (prow + stride < n_nodes) { nkb = rowptr[prow + stride]; nke = rowptr[prow + stride + 1]; }
    }
    auto issue = [&](int stage) {
        if (prow >= n_nodes) return;
        const uint32_t bar = s_addr(&bars[stage]);
        const uint32_t dk = s_addr(ring + (size_t)stage * TMA_SLOT), dx = dk + TMA_KB;
        asm volatile("st.shared.s32 [%0], %1;" ::"r"(dk + TMA_KB + TMA_XB), "r"(pcol) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(TMA_KB + TMA_XB))
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dk),
                     "l"(vals + (size_t)pk * 36), "r"((uint32_t)TMA_KB), "r"(bar)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dx),
                     "l"(x + (size_t)pcol * 6 * NR), "r"((uint32_t)TMA_XB), "r"(bar)
                     : "memory");
        if (++pk == pke) {
            prow += stride;
            pk = nkb; pke = nke;
            if (prow + stride < n_nodes) { nkb = rowptr[prow + stride]; nke = rowptr[prow + stride + 1]; }
        }
        if (prow < n_nodes) pcol = __ldg(colidx + pk);      // consumed at the next issue
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < TMA_STAGES; ++s) issue(s);
    }

    // ---- consumer: current row state + prefetched next row state (uniform across the warp)
    auto load_mask = [&](int r) -> uint32_t {
        if (fixed == nullptr) return 0u;
        uint32_t m = 0;
#pragma unroll
        for (int q = 0; q < 6; ++q) m |= (uint32_t)(fixed[(size_t)r * 6 + q] != 0) << q;
        return m;
    };
    int row = first, kb = 0, ke = 0;
    uint32_t fx = 0;
    if (row < n_nodes) { kb = rowptr[row]; ke = rowptr[row + 1]; fx = load_mask(row); }
    const uint32_t ring_addr = s_addr(ring);
    const uint32_t lane16 = (uint32_t)lane * 16u;
    double2 dot[1];
    dot[0] = make_double2(0.0, 0.0);
    int stage = 0;
    uint32_t phase = 0;
    while (row < n_nodes) {
        const int next = row + stride;
        int nkb2 = 0, nke2 = 0;
        uint32_t nfx = 0;
        if (next < n_nodes) { nkb2 = rowptr[next]; nke2 = rowptr[next + 1]; nfx = load_mask(next); }
        double2 acc[6], xo[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) { acc[r] = make_double2(0.0, 0.0); xo[r] = make_double2(0.0, 0.0); }
        for (int k = kb; k < ke; ++k) {
            {   // wait for this slot's bytes
                const uint32_t bar = s_addr(&bars[stage]);
                uint32_t done = 0;
                while (!done) {
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                        "selp.u32 %0, 1, 0, p;\n\t}"
                        : "=r"(done)
                        : "r"(bar), "r"(phase)
                        : "memory");
                }
            }
            // explicit 32-bit shared addressing: one base per slot, every load is base + immediate
            const uint32_t kbase = ring_addr + (uint32_t)stage * TMA_SLOT;
            const uint32_t xbase = kbase + TMA_KB + lane16;
            int col;
            asm volatile("ld.shared.s32 %0, [%1+%2];" : "=r"(col) : "r"(kbase), "n"(TMA_KB + TMA_XB));
            double2 xv[6];
#define FERS_LDSX(c) lds_f64x2<(c) * 64 * 8>(xv[c], xbase);
            FERS_LDSX(0) FERS_LDSX(1) FERS_LDSX(2) FERS_LDSX(3) FERS_LDSX(4) FERS_LDSX(5)
#undef FERS_LDSX
            if (col == row) {
#pragma unroll
                for (int c = 0; c < 6; ++c) xo[c] = xv[c];
            }
#define FERS_ROW(r)                                                                              \
    {                                                                                            \
        double2 k0, k1, k2;                                                                      \
        lds_f64x2<((r) * 6 + 0) * 8>(k0, kbase);                                                 \
        lds_f64x2<((r) * 6 + 2) * 8>(k1, kbase);                                                 \
        lds_f64x2<((r) * 6 + 4) * 8>(k2, kbase);                                                 \
        fma2(acc[r], k0.x, xv[0]); fma2(acc[r], k0.y, xv[1]);                                    \
        fma2(acc[r], k1.x, xv[2]); fma2(acc[r], k1.y, xv[3]);                                    \
        fma2(acc[r], k2.x, xv[4]); fma2(acc[r], k2.y, xv[5]);                                    \
    }
            FERS_ROW(0) FERS_ROW(1) FERS_ROW(2) FERS_ROW(3) FERS_ROW(4) FERS_ROW(5)
#undef FERS_ROW
            __syncwarp();                       // every lane has consumed the slot
            if (lane == 0) issue(stage);        // refill it with the block TMA_STAGES ahead
            if (++stage == TMA_STAGES) { stage = 0; phase ^= 1u; }
        }
        const size_t d0 = (size_t)row * 6;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const size_t d = d0 + r;
            if (fx & (1u << r)) {
                acc[r] = make_double2(0.0, 0.0);
            } else if (kspring != nullptr) {
                const double ks = kspring[d];
                acc[r].x = fma(ks, xo[r].x, acc[r].x);
                acc[r].y = fma(ks, xo[r].y, acc[r].y);
            }
            reinterpret_cast<double2*>(y + d * NR)[lane] = acc[r];
            if (DOT) { dot[0].x = fma(xo[r].x, acc[r].x, dot[0].x); dot[0].y = fma(xo[r].y, acc[r].y, dot[0].y); }
        }
        row = next; kb = nkb2; ke = nke2; fx = nfx;
    }
    if (DOT) cta_reduce_store<NR, 1>(dot, partial);
}

// 0 = register/LDG variant, 1 = TMA-staged variant (NR = 64 only, default); FERS_SPMM_VARIANT overrides
// ---- warp-specialised TMA variant (variant 2) ----------------------------------------------------
// 8 consumer warps + 2 producer warps per CTA.  Producer lane l of warp 8 (9) feeds the ring of
// consumer warp l (4 + l): it waits on the slot's "empty" mbarrier, then issues the two bulk copies
// that complete on the slot's "full" mbarrier.  Consumers never execute copy-issue code: per block they
// wait on "full", read the slot (25 shared loads), run 72 DFMA and arrive on "empty".
constexpr int WS_CONSUMERS = 8;
constexpr int WS_PRODUCERS = 2;
constexpr int WS_THREADS = (WS_CONSUMERS + WS_PRODUCERS) * 32;
constexpr int WS_STAGES = 3;
constexpr int WS_XO = 6 * 64 * 8;                       // per-consumer copy of the row's own x
constexpr size_t WS_SMEM = (size_t)WS_CONSUMERS * (WS_STAGES * TMA_SLOT + WS_XO) + (size_t)WS_CONSUMERS * WS_STAGES * 16;

__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ uint32_t mbar_test_parity(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done;
}

template <bool DOT>
__global__ void __launch_bounds__(WS_THREADS, 2)
k_spmm64_ws(int n_nodes, const int* __restrict__ rowptr, const int* __restrict__ colidx, const double* __restrict__ vals,
            const uint8_t* __restrict__ fixed, const double* __restrict__ kspring, const double* __restrict__ x,
            double* __restrict__ y, double* __restrict__ partial) {
    constexpr int NR = 64;
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t smem0 = s_addr(smem_dyn);
    constexpr uint32_t RING = WS_STAGES * TMA_SLOT + WS_XO;              // bytes per consumer
    const uint32_t bars0 = smem0 + WS_CONSUMERS * RING;                  // [consumer][stage] {full, empty}
    if (threadIdx.x == 0) {
        for (int i = 0; i < WS_CONSUMERS * WS_STAGES * 2; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars0 + 8u * i));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int stride = gridDim.x * WS_CONSUMERS;

    if (warp >= WS_CONSUMERS) {
        // ================= producer warps =================
        const int cw = (warp - WS_CONSUMERS) * (WS_CONSUMERS / WS_PRODUCERS) + lane;    // consumer served by this lane
        const bool mine = lane < (WS_CONSUMERS / WS_PRODUCERS);
        int prow = mine ? blockIdx.x * WS_CONSUMERS + cw : n_nodes;
        int pk = 0, pke = 0, pcol = 0, nkb = 0, nke = 0, stage = 0;
        uint32_t phase = 0;
        if (prow < n_nodes) {
            pk = rowptr[prow]; pke = rowptr[prow + 1];
            pcol = __ldg(colidx + pk);
            if (prow + stride < n_nodes) { nkb = rowptr[prow + stride]; nke = rowptr[prow + stride + 1]; }
        }
        const uint32_t ring = smem0 + (uint32_t)cw * RING;
        const uint32_t bars = bars0 + (uint32_t)cw * WS_STAGES * 16u;
        while (__any_sync(FULL, prow < n_nodes)) {
            bool issued = false;
            if (prow < n_nodes) {
                const uint32_t full = bars + (uint32_t)stage * 16u, empty = full + 8u;
                if (mbar_test_parity(empty, phase ^ 1u)) {
                    const uint32_t dk = ring + (uint32_t)stage * TMA_SLOT, dx = dk + TMA_KB;
                    asm volatile("st.shared.s32 [%0], %1;" ::"r"(dk + TMA_KB + TMA_XB), "r"(pcol) : "memory");
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full),
                                 "r"((uint32_t)(TMA_KB + TMA_XB))
                                 : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dk),
                                 "l"(vals + (size_t)pk * 36), "r"((uint32_t)TMA_KB), "r"(full)
                                 : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dx),
                                 "l"(x + (size_t)pcol * 6 * NR), "r"((uint32_t)TMA_XB), "r"(full)
                                 : "memory");
                    if (++pk == pke) {
                        prow += stride;
                        pk = nkb; pke = nke;
                        if (prow + stride < n_nodes) { nkb = rowptr[prow + stride]; nke = rowptr[prow + stride + 1]; }
                    }
                    if (prow < n_nodes) pcol = __ldg(colidx + pk);
                    if (++stage == WS_STAGES) { stage = 0; phase ^= 1u; }
                    issued = true;
                }
            }
            if (!__any_sync(FULL, issued)) __nanosleep(64);
        }
        return;
    }

    // ================= consumer warps =================
    const uint32_t ring = smem0 + (uint32_t)warp * RING;
    const uint32_t xo_addr = ring + WS_STAGES * TMA_SLOT + (uint32_t)lane * 16u;
    const uint32_t bars = bars0 + (uint32_t)warp * WS_STAGES * 16u;
    const uint32_t lane16 = (uint32_t)lane * 16u;
    auto load_mask = [&](int r) -> uint32_t {
        if (fixed == nullptr) return 0u;
        uint32_t m = 0;
#pragma unroll
        for (int q = 0; q < 6; ++q) m |= (uint32_t)(fixed[(size_t)r * 6 + q] != 0) << q;
        return m;
    };
    int row = blockIdx.x * WS_CONSUMERS + warp, kb = 0, ke = 0;
    uint32_t fx = 0;
    if (row < n_nodes) { kb = rowptr[row]; ke = rowptr[row + 1]; fx = load_mask(row); }
    double2 dot = make_double2(0.0, 0.0);
    int stage = 0;
    uint32_t phase = 0;
    while (row < n_nodes) {
        const int next = row + stride;
        int nkb2 = 0, nke2 = 0;
        uint32_t nfx = 0;
        if (next < n_nodes) { nkb2 = rowptr[next]; nke2 = rowptr[next + 1]; nfx = load_mask(next); }
        double2 acc[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) acc[r] = make_double2(0.0, 0.0);
        for (int k = kb; k < ke; ++k) {
            const uint32_t full = bars + (uint32_t)stage * 16u;
            mbar_wait_parity(full, phase);
            const uint32_t kbase = ring + (uint32_t)stage * TMA_SLOT;
            const uint32_t xbase = kbase + TMA_KB + lane16;
            int col;
            asm volatile("ld.shared.s32 %0, [%1+%2];" : "=r"(col) : "r"(kbase), "n"(TMA_KB + TMA_XB));
            double2 xv[6];
#define FERS_LDSX(c) lds_f64x2<(c) * 64 * 8>(xv[c], xbase);
            FERS_LDSX(0) FERS_LDSX(1) FERS_LDSX(2) FERS_LDSX(3) FERS_LDSX(4) FERS_LDSX(5)
#undef FERS_LDSX
            if (col == row) {            // diagonal block: keep the row's own x for the epilogue
#pragma unroll
                for (int c = 0; c < 6; ++c)
                    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(xo_addr + (uint32_t)c * 512u), "d"(xv[c].x), "d"(xv[c].y)
                                 : "memory");
            }
#define FERS_ROW(r)                                                                              \
    {                                                                                            \
        double2 k0, k1, k2;                                                                      \
        lds_f64x2<((r) * 6 + 0) * 8>(k0, kbase);                                                 \
        lds_f64x2<((r) * 6 + 2) * 8>(k1, kbase);                                                 \
        lds_f64x2<((r) * 6 + 4) * 8>(k2, kbase);                                                 \
        fma2(acc[r], k0.x, xv[0]); fma2(acc[r], k0.y, xv[1]);                                    \
        fma2(acc[r], k1.x, xv[2]); fma2(acc[r], k1.y, xv[3]);                                    \
        fma2(acc[r], k2.x, xv[4]); fma2(acc[r], k2.y, xv[5]);                                    \
    }
            FERS_ROW(0) FERS_ROW(1) FERS_ROW(2) FERS_ROW(3) FERS_ROW(4) FERS_ROW(5)
#undef FERS_ROW
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full + 8u) : "memory");
            if (++stage == WS_STAGES) { stage = 0; phase ^= 1u; }
        }
        const size_t d0 = (size_t)row * 6;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const size_t d = d0 + r;
            double2 xo = make_double2(0.0, 0.0);
            if (kspring != nullptr || DOT)
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(xo.x), "=d"(xo.y) : "r"(xo_addr + (uint32_t)r * 512u));
            if (fx & (1u << r)) {
                acc[r] = make_double2(0.0, 0.0);
            } else if (kspring != nullptr) {
                const double ks = kspring[d];
                acc[r].x = fma(ks, xo.x, acc[r].x);
                acc[r].y = fma(ks, xo.y, acc[r].y);
            }
            reinterpret_cast<double2*>(y + d * NR)[lane] = acc[r];
            if (DOT) { dot.x = fma(xo.x, acc[r].x, dot.x); dot.y = fma(xo.y, acc[r].y, dot.y); }
        }
        row = next; kb = nkb2; ke = nke2; fx = nfx;
    }
    if (DOT) {
        // consumers only: reduce over the 8 consumer warps through shared memory (named barrier 1)
        __shared__ double2 s_red[WS_CONSUMERS][32];
        s_red[warp][lane] = dot;
        asm volatile("bar.sync 1, %0;" ::"n"(WS_CONSUMERS * 32));
        if (warp == 0) {
            double2 a = s_red[0][lane];
#pragma unroll
            for (int w = 1; w < WS_CONSUMERS; ++w) { a.x += s_red[w][lane].x; a.y += s_red[w][lane].y; }
            reinterpret_cast<double2*>(partial + (size_t)blockIdx.x * NR)[lane] = a;
        }
    }
}

static int g_spmm_variant = -1;
static int spmm_variant() {
    if (g_spmm_variant < 0) {
        const char* e = getenv("FERS_SPMM_VARIANT");
        g_spmm_variant = e ? atoi(e) : 1;   // TMA-staged kernel is the default for nrhs = 64 (profiles/r1_*)
    }
    return g_spmm_variant;
}

// single right-hand side.  Lanes 6g..6g+5 of a warp own one node (five nodes per warp, two lanes idle).
// Per 6x6 block the six lanes of a group issue three 16-byte streaming loads that each cover 96
// CONTIGUOUS bytes of the block (every 32-byte sector is requested exactly once per block: the L2 -> L1
// traffic equals the matrix size), which leaves lane r = 3 rr + cc with the entries (row 2q + rr,
// columns 2cc, 2cc+1), q = 0..2: three partial row sums per lane, combined over cc by two shuffles
// per block ROW.  x of the column node comes through the read-only path (one 16-byte load per block).
constexpr int SPMV_NODES_PER_WARP = 5;
constexpr int SPMV_THREADS = 64;     // small CTAs: short tail on the last wave
__global__ void __launch_bounds__(SPMV_THREADS)
k_spmv1(int n_nodes, const int* __restrict__ rowptr, const int* __restrict__ colidx, const double* __restrict__ vals,
        const uint8_t* __restrict__ fixed, const double* __restrict__ kspring, const double* __restrict__ x,
        double* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int g = lane / 6, r = lane - 6 * g, rr = r / 3, cc = r - 3 * rr;
    const int wglobal = (blockIdx.x * SPMV_THREADS + threadIdx.x) >> 5;
    const int nw = (gridDim.x * SPMV_THREADS) >> 5;
    for (int base = wglobal * SPMV_NODES_PER_WARP; base < n_nodes; base += nw * SPMV_NODES_PER_WARP) {
        const int row = base + g;
        const bool valid = g < SPMV_NODES_PER_WARP && row < n_nodes;
        const int kb = valid ? rowptr[row] : 0, ke = valid ? rowptr[row + 1] : 0;
        double acc[3] = {0.0, 0.0, 0.0};
        for (int k0 = kb; k0 < ke; k0 += 4) {
            double2 a[4][3], xv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = k0 + u;
                xv[u] = make_double2(0.0, 0.0);
#pragma unroll
                for (int q = 0; q < 3; ++q) a[u][q] = make_double2(0.0, 0.0);
                if (k < ke) {                       // predicated: nothing is fetched past the end of the row
                    const int col = __ldg(colidx + k);
                    const double2* ap = reinterpret_cast<const double2*>(vals + (size_t)k * 36) + r;
#pragma unroll
                    for (int q = 0; q < 3; ++q) a[u][q] = __ldcs(ap + 6 * q);
                    xv[u] = __ldg(reinterpret_cast<const double2*>(x + (size_t)col * 6) + cc);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int q = 0; q < 3; ++q) acc[q] = fma(a[u][q].y, xv[u].y, fma(a[u][q].x, xv[u].x, acc[q]));
            }
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const double s1 = __shfl_down_sync(FULL, acc[q], 1), s2 = __shfl_down_sync(FULL, acc[q], 2);
            acc[q] += s1 + s2;                                    // complete on the cc == 0 lanes
        }
        if (valid && cc == 0) {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const size_t d = (size_t)row * 6 + 2 * q + rr;
                double out = acc[q];
                if (fixed != nullptr && fixed[d]) out = 0.0;
                else if (kspring != nullptr) out = fma(kspring[d], x[d], out);
                y[d] = out;
            }
        }
    }
}

// Inverse of each node's diagonal block restricted to the free DOFs (constrained rows/cols ->
// identity).  Gauss-Jordan without pivoting: the restricted block is symmetric positive definite.
__global__ void k_block_jacobi(int n_nodes, const double* __restrict__ vals, const int* __restrict__ diag_slot,
                               const uint8_t* __restrict__ fixed, const double* __restrict__ kspring,
                               double* __restrict__ minv) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_nodes) return;
    double a[6][6], b[6][6];
    const double* blk = vals + (size_t)diag_slot[n] * 36;
    bool fx[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) fx[r] = fixed[(size_t)n * 6 + r] != 0;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double v = blk[r * 6 + c];
            if (fx[r] || fx[c]) v = (r == c) ? 1.0 : 0.0;
            else if (r == c && kspring != nullptr) v += kspring[(size_t)n * 6 + r];
            a[r][c] = v;
            b[r][c] = (r == c) ? 1.0 : 0.0;
        }
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        double piv = a[p][p];
        double inv = (piv != 0.0) ? 1.0 / piv : 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) { a[p][c] *= inv; b[p][c] *= inv; }
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            if (r == p) continue;
            const double f = a[r][p];
#pragma unroll
            for (int c = 0; c < 6; ++c) { a[r][c] -= f * a[p][c]; b[r][c] -= f * b[p][c]; }
        }
    }
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) minv[(size_t)n * 36 + r * 6 + c] = (fx[r] || fx[c]) ? 0.0 : b[r][c];
}

// z = Minv r for one node / two columns
__device__ __forceinline__ void apply_minv(const double* __restrict__ mb, const double2 (&rv)[6], double2 (&z)[6]) {
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        z[r] = make_double2(0.0, 0.0);
#pragma unroll
        for (int c = 0; c < 6; c += 2) {
            const double2 kk = ldg2(mb + r * 6 + c);
            fma2(z[r], kk.x, rv[c]);
            fma2(z[r], kk.y, rv[c + 1]);
        }
    }
}

// r = mask (b - ax), p = Minv r; partial sums: [0] r.z  [1] r.r  [2] (mask b).(mask b)
template <int NR>
__global__ void __launch_bounds__(SPMM_THREADS, CTAS_PER_SM)
k_pcg_init(int n_nodes, const uint8_t* __restrict__ fixed, const double* __restrict__ minv,
           const double* __restrict__ b, const double* __restrict__ ax, double* __restrict__ r,
           double* __restrict__ p, double* __restrict__ partial) {
    constexpr int CP = NR / 2, G = 32 / CP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / CP, cp = lane % CP;
    const int per_cta = (n_nodes + gridDim.x - 1) / gridDim.x;
    const int row0 = blockIdx.x * per_cta, row1 = min(n_nodes, row0 + per_cta);
    double2 s[3];
    s[0] = s[1] = s[2] = make_double2(0.0, 0.0);
    for (int row = row0 + warp * G + g; row < row1; row += SPMM_WARPS * G) {
        double2 rv[6], z[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const size_t d = (size_t)row * 6 + q;
            double2 bv = ldg2(b + d * NR + 2 * cp);
            if (fixed[d]) bv = make_double2(0.0, 0.0);
            s[2].x = fma(bv.x, bv.x, s[2].x); s[2].y = fma(bv.y, bv.y, s[2].y);
            if (ax != nullptr) {
                const double2 av = ldg2(ax + d * NR + 2 * cp);
                bv.x -= av.x; bv.y -= av.y;
                if (fixed[d]) bv = make_double2(0.0, 0.0);
            }
            rv[q] = bv;
            reinterpret_cast<double2*>(r + d * NR)[cp] = bv;
            s[1].x = fma(bv.x, bv.x, s[1].x); s[1].y = fma(bv.y, bv.y, s[1].y);
        }
        apply_minv(minv + (size_t)row * 36, rv, z);
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            reinterpret_cast<double2*>(p + ((size_t)row * 6 + q) * NR)[cp] = z[q];
            s[0].x = fma(rv[q].x, z[q].x, s[0].x); s[0].y = fma(rv[q].y, z[q].y, s[0].y);
        }
    }
    cta_reduce_store<NR, 3>(s, partial);
}

// alpha = rz / pAp ; x += alpha p ; r -= alpha Ap ; partial sums: [0] r.(Minv r)  [1] r.r
template <int NR>
__global__ void __launch_bounds__(SPMM_THREADS, CTAS_PER_SM)
k_pcg_update(int n_nodes, const double* __restrict__ minv, const double* __restrict__ rz,
             const double* __restrict__ pAp, const double* __restrict__ p, const double* __restrict__ Ap,
             double* __restrict__ x, double* __restrict__ r, double* __restrict__ partial) {
    constexpr int CP = NR / 2, G = 32 / CP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / CP, cp = lane % CP;
    const int per_cta = (n_nodes + gridDim.x - 1) / gridDim.x;
    const int row0 = blockIdx.x * per_cta, row1 = min(n_nodes, row0 + per_cta);
    const double2 num = ldg2(rz + 2 * cp), den = ldg2(pAp + 2 * cp);
    double2 alpha;
    alpha.x = (den.x > 0.0 && num.x > 0.0) ? num.x / den.x : 0.0;
    alpha.y = (den.y > 0.0 && num.y > 0.0) ? num.y / den.y : 0.0;
    double2 s[2];
    s[0] = s[1] = make_double2(0.0, 0.0);
    for (int row = row0 + warp * G + g; row < row1; row += SPMM_WARPS * G) {
        double2 rv[6], z[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const size_t o = ((size_t)row * 6 + q) * NR + 2 * cp;
            const double2 pv = ldg2(p + o), av = ldg2(Ap + o);
            double2 xv = *reinterpret_cast<const double2*>(x + o);
            double2 rr = *reinterpret_cast<const double2*>(r + o);
            xv.x = fma(alpha.x, pv.x, xv.x); xv.y = fma(alpha.y, pv.y, xv.y);
            rr.x = fma(-alpha.x, av.x, rr.x); rr.y = fma(-alpha.y, av.y, rr.y);
            *reinterpret_cast<double2*>(x + o) = xv;
            *reinterpret_cast<double2*>(r + o) = rr;
            rv[q] = rr;
            s[1].x = fma(rr.x, rr.x, s[1].x); s[1].y = fma(rr.y, rr.y, s[1].y);
        }
        apply_minv(minv + (size_t)row * 36, rv, z);
#pragma unroll
        for (int q = 0; q < 6; ++q) { s[0].x = fma(rv[q].x, z[q].x, s[0].x); s[0].y = fma(rv[q].y, z[q].y, s[0].y); }
    }
    cta_reduce_store<NR, 2>(s, partial);
}

// beta = rz_new / rz_old ; p = Minv r + beta p
template <int NR>
__global__ void __launch_bounds__(SPMM_THREADS, CTAS_PER_SM)
k_pcg_direction(int n_nodes, const double* __restrict__ minv, const double* __restrict__ rz_new,
                const double* __restrict__ rz_old, const double* __restrict__ r, double* __restrict__ p) {
    constexpr int CP = NR / 2, G = 32 / CP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / CP, cp = lane % CP;
    const int per_cta = (n_nodes + gridDim.x - 1) / gridDim.x;
    const int row0 = blockIdx.x * per_cta, row1 = min(n_nodes, row0 + per_cta);
    const double2 num = ldg2(rz_new + 2 * cp), den = ldg2(rz_old + 2 * cp);
    double2 beta;
    beta.x = (den.x > 0.0) ? num.x / den.x : 0.0;
    beta.y = (den.y > 0.0) ? num.y / den.y : 0.0;
    for (int row = row0 + warp * G + g; row < row1; row += SPMM_WARPS * G) {
        double2 rv[6], z[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) rv[q] = ldg2(r + ((size_t)row * 6 + q) * NR + 2 * cp);
        apply_minv(minv + (size_t)row * 36, rv, z);
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const size_t o = ((size_t)row * 6 + q) * NR + 2 * cp;
            double2 pv = *reinterpret_cast<const double2*>(p + o);
            pv.x = fma(beta.x, pv.x, z[q].x); pv.y = fma(beta.y, pv.y, z[q].y);
            *reinterpret_cast<double2*>(p + o) = pv;
        }
    }
}

// out[j] = sum_b partial[b][j] in a fixed order (deterministic).  One CTA of 32 x RED_GROUPS threads
// per 32 outputs: thread (j, g) sums blocks g, g+RED_GROUPS, ... (independent, coalesced loads), the
// groups are then combined through shared memory in index order.
constexpr int RED_GROUPS = 16;
__global__ void __launch_bounds__(32 * RED_GROUPS)
k_reduce_partials(int n_blocks, int width, const double* __restrict__ partial, double* __restrict__ out) {
    __shared__ double s[RED_GROUPS][33];
    const int tj = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + tj;
    double a = 0.0;
    if (j < width) {
#pragma unroll 4
        for (int b = g; b < n_blocks; b += RED_GROUPS) a += partial[(size_t)b * width + j];
    }
    s[g][tj] = a;
    __syncthreads();
    if (g == 0 && j < width) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < RED_GROUPS; ++q) t += s[q][tj];
        out[j] = t;
    }
}
static inline void launch_reduce(int n_blocks, int width, const double* partial, double* out, cudaStream_t st) {
    k_reduce_partials<<<(width + 31) / 32, 32 * RED_GROUPS, 0, st>>>(n_blocks, width, partial, out);
}

// Device-time profile of the last PCG solve (fers_b200_pcg_profile).
constexpr int PROFILE_STRIDE = 16;
constexpr int PROFILE_MAX = 64;
struct Profile {
    cudaEvent_t ev[6 * PROFILE_MAX];
    bool have_events = false;
    double spmm_ms = 0, update_ms = 0, direction_ms = 0, iter_ms = 0;
    int samples = 0, iterations = 0;
    int ensure_events() {
        if (have_events) return 0;
        for (int i = 0; i < 6 * PROFILE_MAX; ++i) FERS_CK(cudaEventCreate(&ev[i]));
        have_events = true;
        return 0;
    }
    // call after the stream has been synchronised
    void collect(int n, int iters) {
        double a = 0, b = 0, c = 0, d = 0;
        for (int s = 0; s < n; ++s) {
            float t = 0;
            cudaEventElapsedTime(&t, ev[6 * s + 0], ev[6 * s + 1]); a += t;
            cudaEventElapsedTime(&t, ev[6 * s + 2], ev[6 * s + 3]); b += t;
            cudaEventElapsedTime(&t, ev[6 * s + 4], ev[6 * s + 5]); c += t;
            cudaEventElapsedTime(&t, ev[6 * s + 0], ev[6 * s + 5]); d += t;
        }
        samples = n; iterations = iters;
        spmm_ms = n ? a / n : 0; update_ms = n ? b / n : 0; direction_ms = n ? c / n : 0; iter_ms = n ? d / n : 0;
    }
};
static Profile g_profile;

static int persistent_grid(int n_nodes, int nr) {
    const int groups_per_cta = SPMM_WARPS * (32 / (nr / 2));
    int need = (n_nodes + groups_per_cta - 1) / groups_per_cta;
    int cap = sm_count() * CTAS_PER_SM;
    return need < cap ? (need > 0 ? need : 1) : cap;
}

// Launch y = K x with fused x.y partials (PCG); returns the grid size used (= rows of `partial`).
template <int NR>
static int launch_spmm_dot(int n_nodes, const int* rowptr, const int* colidx, const double* vals, const uint8_t* fixed,
                           const double* kspring, const double* x, double* y, double* partial, cudaStream_t st) {
    if (NR == 64 && spmm_variant() == 2) {
        static PerDeviceOnce attr;
        if (attr.needed()) {
            cudaFuncSetAttribute(k_spmm64_ws<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM);
            cudaFuncSetAttribute(k_spmm64_ws<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM);
            attr.done();
        }
        int grid = sm_count() * 2;
        const int need = (n_nodes + WS_CONSUMERS - 1) / WS_CONSUMERS;
        if (grid > need) grid = need;
        k_spmm64_ws<true><<<grid, WS_THREADS, WS_SMEM, st>>>(n_nodes, rowptr, colidx, vals, fixed, kspring, x, y, partial);
        return grid;
    }
    if (NR == 64 && spmm_variant() == 1) {
        static PerDeviceOnce attr;
        if (attr.needed()) {
            cudaFuncSetAttribute(k_spmm64_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM);
            cudaFuncSetAttribute(k_spmm64_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM);
            attr.done();
        }
        int grid = sm_count() * 2;
        const int need = (n_nodes + SPMM_WARPS - 1) / SPMM_WARPS;
        if (grid > need) grid = need;
        k_spmm64_tma<true><<<grid, SPMM_THREADS, TMA_SMEM, st>>>(n_nodes, rowptr, colidx, vals, fixed, kspring, x, y, partial);
        return grid;
    }
    const int grid = persistent_grid(n_nodes, NR);
    k_spmm<NR, true><<<grid, SPMM_THREADS, 0, st>>>(n_nodes, rowptr, colidx, vals, fixed, kspring, x, y, partial);
    return grid;
}

template <int NR>
static int run_spmm(int n_nodes, const int* rowptr, const int* colidx, const double* vals, const uint8_t* fixed,
                    const double* kspring, const double* x, double* y, double* partial, bool dot, cudaStream_t st) {
    if (dot) {
        launch_spmm_dot<NR>(n_nodes, rowptr, colidx, vals, fixed, kspring, x, y, partial, st);
    } else if (NR == 64 && spmm_variant() == 2) {
        static PerDeviceOnce attr;
        if (attr.needed()) {
            cudaFuncSetAttribute(k_spmm64_ws<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM);
            attr.done();
        }
        int grid = sm_count() * 2;
        const int need = (n_nodes + WS_CONSUMERS - 1) / WS_CONSUMERS;
        if (grid > need) grid = need;
        k_spmm64_ws<false><<<grid, WS_THREADS, WS_SMEM, st>>>(n_nodes, rowptr, colidx, vals, fixed, kspring, x, y, partial);
    } else if (NR == 64 && spmm_variant() == 1) {
        static PerDeviceOnce attr;
        if (attr.needed()) {
            cudaFuncSetAttribute(k_spmm64_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM);
            attr.done();
        }
        int grid = sm_count() * 2;
        const int need = (n_nodes + SPMM_WARPS - 1) / SPMM_WARPS;
        if (grid > need) grid = need;
        k_spmm64_tma<false><<<grid, SPMM_THREADS, TMA_SMEM, st>>>(n_nodes, rowptr, colidx, vals, fixed, kspring, x, y, partial);
    } else {
        // plain products: one node group per warp slot (fine-grained grid)
        const int groups_per_cta = SPMM_WARPS * (32 / (NR / 2));
        int grid = (n_nodes + groups_per_cta - 1) / groups_per_cta;
        if (grid < 1) grid = 1;
        k_spmm<NR, false><<<grid, SPMM_THREADS, 0, st>>>(n_nodes, rowptr, colidx, vals, fixed, kspring, x, y, partial);
    }
    FERS_LAUNCH_CHECK();
    return 0;
}

template <int NR>
static int run_pcg(int n_nodes, const int* rowptr, const int* colidx, const double* vals, const uint8_t* fixed,
                   const double* kspring, const double* minv, const double* b, double* x, int warm_start, double tol,
                   int max_iter, int check_every, int* iters_out, double* relres_out, void* workspace, cudaStream_t st) {
    const size_t nv = (size_t)n_nodes * 6 * NR;
    const int grid = persistent_grid(n_nodes, NR);
    char* w = (char*)workspace;
    double* r = (double*)w; w += align_up(nv * 8);
    double* p = (double*)w; w += align_up(nv * 8);
    double* Ap = (double*)w; w += align_up(nv * 8);
    double* partial = (double*)w; w += align_up((size_t)sm_count() * CTAS_PER_SM * 3 * NR * 8);
    // device scalars, NR wide each: [0] pAp | [1..3] init sums (r.z, r.r, b.b) | [4,5] and [6,7] update sums
    double* scal = (double*)w;
    double* s_pAp = scal;
    double* s_init = scal + NR;
    double* s_pair[2] = {scal + 4 * NR, scal + 6 * NR};
    const double* s_bb = scal + 3 * NR;
    const double* rz_cur = s_init;
    const double* rr_cur = s_init + NR;
    double host[2 * 64];
    if (!warm_start) FERS_CK(cudaMemsetAsync(x, 0, nv * 8, st));
    const double* ax = nullptr;
    if (warm_start) {
        if (run_spmm<NR>(n_nodes, rowptr, colidx, vals, fixed, kspring, x, Ap, partial, false, st)) return 1;
        ax = Ap;
    }
    k_pcg_init<NR><<<grid, SPMM_THREADS, 0, st>>>(n_nodes, fixed, minv, b, ax, r, p, partial);
    FERS_LAUNCH_CHECK();
    launch_reduce(grid, 3 * NR, partial, s_init, st);
    FERS_LAUNCH_CHECK();
    count_launches(2);
    int it = 0;
    bool converged = false;
    auto check = [&]() -> int {
        FERS_CK(cudaMemcpyAsync(host, rr_cur, NR * 8, cudaMemcpyDeviceToHost, st));
        FERS_CK(cudaMemcpyAsync(host + NR, s_bb, NR * 8, cudaMemcpyDeviceToHost, st));
        FERS_CK(cudaStreamSynchronize(st));
        converged = true;
        for (int c = 0; c < NR; ++c) {
            const double bb = host[NR + c], rr = host[c];
            const double rel = bb > 0.0 ? sqrt(rr / bb) : 0.0;
            relres_out[c] = rel;
            if (!(rel <= tol)) converged = false;
        }
        return 0;
    };
    if (check()) return 1;
    Profile& prof = g_profile;
    if (prof.ensure_events()) return 1;
    int n_samples = 0;
    while (!converged && it < max_iter) {
        const int burst = (max_iter - it) < check_every ? (max_iter - it) : check_every;
        for (int q = 0; q < burst; ++q, ++it) {
            double* pair = s_pair[it & 1];
            // every PROFILE_STRIDE-th iteration is bracketed by CUDA events on this stream
            const bool sample = (it % PROFILE_STRIDE == PROFILE_STRIDE / 2) && n_samples < PROFILE_MAX;
            cudaEvent_t* ev = sample ? prof.ev + 6 * n_samples : nullptr;
            if (sample) cudaEventRecord(ev[0], st);
            const int sgrid = launch_spmm_dot<NR>(n_nodes, rowptr, colidx, vals, fixed, kspring, p, Ap, partial, st);
            if (sample) cudaEventRecord(ev[1], st);
            launch_reduce(sgrid, NR, partial, s_pAp, st);
            if (sample) cudaEventRecord(ev[2], st);
            k_pcg_update<NR><<<grid, SPMM_THREADS, 0, st>>>(n_nodes, minv, rz_cur, s_pAp, p, Ap, x, r, partial);
            if (sample) cudaEventRecord(ev[3], st);
            launch_reduce(grid, 2 * NR, partial, pair, st);   // r.z, r.r
            if (sample) cudaEventRecord(ev[4], st);
            k_pcg_direction<NR><<<grid, SPMM_THREADS, 0, st>>>(n_nodes, minv, pair, rz_cur, r, p);
            if (sample) { cudaEventRecord(ev[5], st); ++n_samples; }
            FERS_LAUNCH_CHECK();
            rz_cur = pair;
            rr_cur = pair + NR;
        }
        count_launches(5LL * burst);
        if (check()) return 1;
    }
    prof.collect(n_samples, it);
    *iters_out = it;
    return 0;
}

}  // namespace fers

using namespace fers;

#define FERS_DISPATCH_NR(nr, CALL)                                                         \
    switch (nr) {                                                                          \
        case 8: { constexpr int NR = 8; CALL; break; }                                     \
        case 16: { constexpr int NR = 16; CALL; break; }                                   \
        case 32: { constexpr int NR = 32; CALL; break; }                                   \
        case 64: { constexpr int NR = 64; CALL; break; }                                   \
        default: fers::set_error("nrhs must be one of 8, 16, 32, 64 (pad the right-hand sides)"); return 2; \
    }

extern "C" {

int fers_b200_spmm(int n_nodes, int nrhs, const int* rowptr, const int* colidx, const double* vals,
                   const uint8_t* fixed, const double* kspring, const double* x, double* y, void* stream) {
    cudaStream_t st = as_stream(stream);
    if (n_nodes <= 0) return 0;
    if (nrhs == 1) {
        const int per_cta = (SPMV_THREADS / 32) * SPMV_NODES_PER_WARP;
        const int grid = (n_nodes + per_cta - 1) / per_cta;
        k_spmv1<<<grid, SPMV_THREADS, 0, st>>>(n_nodes, rowptr, colidx, vals, fixed, kspring, x, y);
        FERS_LAUNCH_CHECK();
        count_launches(1);
        return 0;
    }
    int rc = 0;
    FERS_DISPATCH_NR(nrhs, rc = run_spmm<NR>(n_nodes, rowptr, colidx, vals, fixed, kspring, x, y, nullptr, false, st));
    count_launches(1);
    return rc;
}

int fers_b200_block_jacobi(int n_nodes, const double* vals, const int* diag_slot, const uint8_t* fixed,
                           const double* kspring, double* minv, void* stream) {
    if (n_nodes <= 0) return 0;
    const int T = 128;
    k_block_jacobi<<<(n_nodes + T - 1) / T, T, 0, as_stream(stream)>>>(n_nodes, vals, diag_slot, fixed, kspring, minv);
    FERS_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int fers_b200_set_spmm_variant(int variant) {
    FERS_REQUIRE(variant >= 0 && variant <= 2,
                 "set_spmm_variant: 0 (register/LDG), 1 (TMA-staged) or 2 (TMA, warp-specialised); 1 and 2 need nrhs = 64");
    g_spmm_variant = variant;
    return 0;
}

int fers_b200_pcg_profile(double* out6) {
    const Profile& p = g_profile;
    out6[0] = p.spmm_ms; out6[1] = p.update_ms; out6[2] = p.direction_ms;
    out6[3] = p.samples; out6[4] = p.iterations; out6[5] = p.iter_ms;
    return 0;
}

size_t fers_b200_pcg_workspace(int n_nodes, int nrhs) {
    const size_t nv = (size_t)n_nodes * 6 * nrhs;
    return 3 * align_up(nv * 8) + align_up((size_t)sm_count() * CTAS_PER_SM * 3 * nrhs * 8) + align_up(16 * (size_t)nrhs * 8) + 1024;
}

int fers_b200_pcg_solve(int n_nodes, int nrhs, const int* rowptr, const int* colidx, const double* vals,
                        const uint8_t* fixed, const double* kspring, const double* minv, const double* b, double* x,
                        int warm_start, double tol, int max_iter, int check_every, int* iters_out, double* relres_out,
                        void* workspace, size_t workspace_bytes, void* stream) {
    FERS_REQUIRE(n_nodes > 0, "pcg_solve: empty system");
    FERS_REQUIRE(workspace_bytes >= fers_b200_pcg_workspace(n_nodes, nrhs), "pcg_solve: workspace too small");
    FERS_REQUIRE(fixed != nullptr && minv != nullptr, "pcg_solve: fixed mask and preconditioner are required");
    if (check_every < 1) check_every = 25;
    int rc = 0;
    FERS_DISPATCH_NR(nrhs, rc = run_pcg<NR>(n_nodes, rowptr, colidx, vals, fixed, kspring, minv, b, x, warm_start, tol,
                                            max_iter, check_every, iters_out, relres_out, workspace, as_stream(stream)));
    return rc;
}

}  // extern "C"
