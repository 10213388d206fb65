// ts_probe.cu -- what does tcgen05.cp.128x256b put where, and does tcgen05.mma kind::i8 with the A operand in tensor memory
// read it back as the same matrix the shared-memory descriptor describes?  Then: issue rate of the digit-plane product
// stream of k_oz_syrk (28 MMAs of 128 x 64 x 32 per K chunk) with A from shared memory (SS) against A copied once per
// plane into tensor memory (TS), operands resident in shared memory (no TMA traffic): the ceiling of either form.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/ts_probe tools/micro/ts_probe.cu && /tmp/ts_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int M = 128, N = 64, KC = 32, S = 7;
constexpr uint32_t A_BYTES = M * KC, B_BYTES = N * KC;
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr16, uint32_t lbo, uint32_t sbo) {
    const uint32_t lo = saddr16 | ((lbo >> 4) << 16);
    const uint32_t hi = (sbo >> 4) | (1u << 14);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(IDESC), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(db), "r"(IDESC), "r"(acc) : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t d) { asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(d) : "memory"); }
__device__ __forceinline__ void ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// operand tile in shared memory: [8-row group][16-byte K piece][8 rows][16 B]  (K-major core matrices, no swizzle)
__host__ __device__ inline int tile_off(int r, int k) { return (r >> 3) * (KC * 8) + (k >> 4) * 128 + (r & 7) * 16 + (k & 15); }
__host__ __device__ inline int8_t a_val(int r, int k) { return (int8_t)(((r * 7 + k * 13) % 251) - 125); }
__host__ __device__ inline int8_t b_val(int n, int k) { return (int8_t)(((n * 5 + k * 3 + 1) % 201) - 100); }

// out_a[128][8]: tensor-memory words of the copied A tile; out_ts / out_ss [128][64]: D = A B^T by either form
__global__ void __launch_bounds__(128) k_layout(uint32_t* out_a, int* out_ts, int* out_ss) {
    __shared__ __align__(1024) uint8_t sa[A_BYTES];
    __shared__ __align__(1024) uint8_t sb[B_BYTES];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < M * KC; i += 128) sa[tile_off(i / KC, i % KC)] = (uint8_t)a_val(i / KC, i % KC);
    for (int i = threadIdx.x; i < N * KC; i += 128) sb[tile_off(i / KC, i % KC)] = (uint8_t)b_val(i / KC, i % KC);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the tensor core (async proxy)
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t A_COL = 448;
    if (threadIdx.x == 0) {
        const uint64_t da = desc(smem_u32(sa) >> 4, 128, KC * 8), db = desc(smem_u32(sb) >> 4, 128, KC * 8);
        cp_128x256b(tmem + A_COL, da);
        mma_ts(tmem + 0, tmem + A_COL, db, 0);
        mma_ss(tmem + 64, da, db, 0);
        commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    fence_after();
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * warp) << 16);
    const int r = 32 * warp + lane;
    uint32_t v[8];
    ld8(lane_addr + A_COL, v);
    for (int j = 0; j < 8; ++j) out_a[r * 8 + j] = v[j];
    for (int cb = 0; cb < 64; cb += 8) {
        ld8(lane_addr + cb, v);
        for (int j = 0; j < 8; ++j) out_ts[r * 64 + cb + j] = (int)v[j];
        ld8(lane_addr + 64 + cb, v);
        for (int j = 0; j < 8; ++j) out_ss[r * 64 + cb + j] = (int)v[j];
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// issue-rate loop: `chunks` K chunks of S(S+1)/2 MMAs; mode 0 = SS, 1 = TS (S copies per chunk), 2 = TS with the copies of
// chunk c+1 issued ahead of the MMAs of chunk c into a second tensor-memory buffer (S <= 6 only: 2 * 8 S + 64 S <= 512)
template <int SS_>
__global__ void __launch_bounds__(128) k_rate(int chunks, int mode, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t ring[];       // 4 stages of S (A + B) tiles, contents irrelevant
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    constexpr uint32_t STAGE = SS_ * (A_BYTES + B_BYTES);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (int)(4 * STAGE / 4); i += 128) reinterpret_cast<uint32_t*>(ring)[i] = 0x01020301u * (i & 3);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t A_COL = SS_ * N;
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int c = 0; c < chunks; ++c) {
            const uint32_t sa = (smem_u32(ring) + (c & 3) * STAGE) >> 4, sb = sa + ((SS_ * A_BYTES) >> 4);
            if (mode == 1)
                for (int p = 0; p < SS_; ++p) cp_128x256b(tmem + A_COL + 8 * p, desc(sa + ((p * A_BYTES) >> 4), 128, KC * 8));
#pragma unroll
            for (int p = 0; p < SS_; ++p) {
                const uint64_t da = desc(sa + ((p * A_BYTES) >> 4), 128, KC * 8);
#pragma unroll
                for (int q = 0; q + p < SS_; ++q) {
                    const uint64_t db = desc(sb + ((q * B_BYTES) >> 4), 128, KC * 8);
                    const uint32_t acc = (c > 0 || p > 0) ? 1u : 0u;
                    if (mode == 0) mma_ss(tmem + (uint32_t)(p + q) * N, da, db, acc);
                    else mma_ts(tmem + (uint32_t)(p + q) * N, tmem + A_COL + 8 * p, db, acc);
                }
            }
        }
        commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// bulk-copy streaming rate: one producer lane per CTA keeps `depth` copies of `bytes` each in flight from an L2-resident buffer
// into a shared-memory ring (what the operand stream of k_oz_syrk does, without the MMAs)
__global__ void __launch_bounds__(64) k_stream(const uint8_t* src, size_t span, int bytes, int depth, int iters, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t ring[];
    __shared__ __align__(8) uint64_t bars[64];
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        size_t off = ((size_t)blockIdx.x * 977 * bytes) % span;
        for (int it = 0; it < iters + depth; ++it) {
            const int st = it % depth;
            if (it >= depth) mbar_wait(smem_u32(&bars[st]), ((it / depth) - 1) & 1);
            if (it < iters) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[st])), "r"((uint32_t)bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + (size_t)st * bytes)),
                             "l"(src + off), "r"((uint32_t)bytes), "r"(smem_u32(&bars[st])) : "memory");
                off += (size_t)bytes * 3;
                if (off + bytes > span) off -= span - bytes;
                off &= ~(size_t)127;
            }
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}

int main() {
    uint32_t* d_a; int *d_ts, *d_ss;
    CK(cudaMalloc(&d_a, 128 * 8 * 4)); CK(cudaMalloc(&d_ts, 128 * 64 * 4)); CK(cudaMalloc(&d_ss, 128 * 64 * 4));
    k_layout<<<1, 128>>>(d_a, d_ts, d_ss);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> ha(128 * 8); std::vector<int> hts(128 * 64), hss(128 * 64);
    CK(cudaMemcpy(ha.data(), d_a, ha.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hts.data(), d_ts, hts.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hss.data(), d_ss, hss.size() * 4, cudaMemcpyDeviceToHost));
    int bad_a = 0, bad_ts = 0, bad_ss = 0;
    for (int r = 0; r < 128; ++r)
        for (int j = 0; j < 8; ++j) {
            uint32_t want = 0;
            for (int b = 0; b < 4; ++b) want |= (uint32_t)(uint8_t)a_val(r, 4 * j + b) << (8 * b);
            if (ha[r * 8 + j] != want) { if (bad_a < 4) printf("A word mismatch lane %d col %d: got %08x want %08x\n", r, j, ha[r * 8 + j], want); ++bad_a; }
        }
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < 64; ++n) {
            int want = 0;
            for (int k = 0; k < KC; ++k) want += (int)a_val(r, k) * (int)b_val(n, k);
            if (hts[r * 64 + n] != want) { if (bad_ts < 4) printf("TS mismatch (%d,%d): got %d want %d\n", r, n, hts[r * 64 + n], want); ++bad_ts; }
            if (hss[r * 64 + n] != want) { if (bad_ss < 4) printf("SS mismatch (%d,%d): got %d want %d\n", r, n, hss[r * 64 + n], want); ++bad_ss; }
        }
    printf("layout: copied A words wrong %d / 1024, TS product wrong %d / 8192, SS product wrong %d / 8192\n", bad_a, bad_ts, bad_ss);
    if (bad_a) {            // show where the bytes of lane 0, 1, 8 ended up
        for (int r : {0, 1, 8, 32})
            { printf("lane %3d:", r); for (int j = 0; j < 8; ++j) printf(" %08x", ha[r * 8 + j]); printf("\n"); }
    }
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    long long* d_cyc; CK(cudaMalloc(&d_cyc, sms * 8));
    std::vector<long long> hc(sms);
    const int chunks = 4000;
    auto run = [&](auto kern, int s, int mode, const char* name) {
        const size_t smem = 4 * (size_t)s * (A_BYTES + B_BYTES);
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int grid : {1, sms}) {
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            kern<<<grid, 128, smem>>>(chunks, mode, d_cyc); CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0)); kern<<<grid, 128, smem>>>(chunks, mode, d_cyc); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
            CK(cudaMemcpy(hc.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost));
            const double mmas = (double)chunks * s * (s + 1) / 2;
            const double tops = 2.0 * mmas * M * N * KC * grid / (ms * 1e-3) / 1e12;
            printf("%-28s S=%d grid %3d: %.3f ms, %.1f cycles per MMA (floor 32), %.0f TOP/s\n", name, s, grid, ms, hc[0] / mmas, tops);
        }
    };
    run(k_rate<7>, 7, 0, "SS (A from shared memory)");
    run(k_rate<7>, 7, 1, "TS (A copied to TMEM)");
    run(k_rate<6>, 6, 0, "SS (A from shared memory)");
    run(k_rate<6>, 6, 1, "TS (A copied to TMEM)");
    // operand-stream ceiling: 160 KB in flight per SM, copy sizes 1 KB .. 32 KB, source 96 MB (L2-resident after the warm-up pass)
    {
        const size_t span = 96u << 20;
        uint8_t* d_src; CK(cudaMalloc(&d_src, span)); CK(cudaMemset(d_src, 1, span));
        CK(cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        for (int bytes : {1024, 2048, 4096, 8192, 16384, 32768}) {
            for (int inflight_kb : {40, 160}) {
                const int depth = inflight_kb * 1024 / bytes > 64 ? 64 : (inflight_kb * 1024 / bytes < 1 ? 1 : inflight_kb * 1024 / bytes);
                const int iters = (int)((256u << 20) / bytes / 16);
                cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                k_stream<<<sms, 64, (size_t)depth * bytes>>>(d_src, span, bytes, depth, iters, d_cyc); CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0)); k_stream<<<sms, 64, (size_t)depth * bytes>>>(d_src, span, bytes, depth, iters, d_cyc); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
                float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
                CK(cudaMemcpy(hc.data(), d_cyc, sms * 8, cudaMemcpyDeviceToHost));
                const double tot = (double)iters * bytes * sms;
                printf("bulk copies of %5d B, %2d in flight (%3d KB): %.2f TB/s, %.1f B/clk/SM\n", bytes, depth, depth * bytes / 1024, tot / (ms * 1e-3) / 1e12,
                       (double)iters * bytes / hc[0]);
            }
        }
    }
    return 0;
}
