// mma_rate.cu -- issue rate of tcgen05.mma (cta_group::1, M = 128) with both operands resident in shared memory, by kind
// (i8 / f16), N (64 / 128 / 256), number of accumulators cycled through and number of distinct operand tiles: which of
// operand bytes, accumulator turnover or instruction count sets the 95 cycles per 128 x 64 x 32 int8 MMA seen in k_oz_syrk.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/mma_rate tools/micro/mma_rate.cu && /tmp/mma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ bool elect() {
    uint32_t pred;
    __syncwarp();
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr16, uint32_t lbo, uint32_t sbo) {
    const uint32_t lo = saddr16 | ((lbo >> 4) << 16);
    const uint32_t hi = (sbo >> 4) | (1u << 14);
    return ((uint64_t)hi << 32) | lo;
}
template <int KIND>     // 0 = i8, 1 = f16
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// n_mma MMAs of 128 x N x (32 bytes of K); accumulator = (i % n_acc) * N columns, A tile = i % n_a, B tile = i % n_b
template <int KIND, int N>
__global__ void __launch_bounds__(128) k_rate(int n_mma, int n_acc, int n_a, int n_b, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t tiles[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    constexpr uint32_t A_BYTES = 128 * 32, B_BYTES = N * 32;
    const int warp = threadIdx.x >> 5;
    const int total = n_a * A_BYTES + n_b * B_BYTES;
    for (int i = threadIdx.x; i < total / 4; i += 128) reinterpret_cast<uint32_t*>(tiles)[i] = KIND == 0 ? 0x01020301u * (i & 3) : 0x3c003c00u;
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
    // i8: D = S32 (2), A = B = S8 (1); f16: D = F32 (1), A = B = F16 (0)
    const uint32_t idesc = (KIND == 0 ? ((2u << 4) | (1u << 7) | (1u << 10)) : (1u << 4)) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 0) {               // converged warp, elected lane (a lone thread 0 makes the compiler wrap every MMA in a
                                   // divergence loop that moves its operands to uniform registers: 84 cycles per MMA, seen in r2o)
        const uint32_t sa0 = smem_u32(tiles) >> 4, sb0 = sa0 + ((n_a * A_BYTES) >> 4);
        const long long t0 = clock64();
        int ia = 0, ib = 0, ic = 0;
        for (int i = 0; i < n_mma; ++i) {
            if (elect())
                mma_ss<KIND>(tmem + (uint32_t)ic * N, desc(sa0 + ((ia * A_BYTES) >> 4), 128, 256), desc(sb0 + ((ib * B_BYTES) >> 4), 128, 256), idesc, i >= n_acc ? 1u : 0u);
            __syncwarp();
            if (++ia == n_a) ia = 0;
            if (++ib == n_b) ib = 0;
            if (++ic == n_acc) ic = 0;
        }
        if (elect()) commit(smem_u32(&bar));
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        if ((threadIdx.x & 31) == 0) cycles[blockIdx.x] = clock64() - t0;
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// the same stream issued by `n_iss` warps at once (one lane each, disjoint accumulators): is the per-instruction minimum a
// property of the issuing thread or of the tensor core?
template <int KIND, int N>
__global__ void __launch_bounds__(128) k_rate_multi(int n_mma, int n_iss, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t tiles[];
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t s_tmem;
    __shared__ long long s_cyc[4];
    constexpr uint32_t A_BYTES = 128 * 32, B_BYTES = N * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (int)(4 * (A_BYTES + B_BYTES) / 4); i += 128) reinterpret_cast<uint32_t*>(tiles)[i] = KIND == 0 ? 0x01020301u * (i & 3) : 0x3c003c00u;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = (KIND == 0 ? ((2u << 4) | (1u << 7) | (1u << 10)) : (1u << 4)) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp < n_iss) {            // converged warp, one elected lane issues: UTCIMMA back to back on uniform registers (as k_oz_syrk does)
        const uint32_t sa = (smem_u32(tiles) >> 4) + ((warp * A_BYTES) >> 4), sb = (smem_u32(tiles) >> 4) + ((4 * A_BYTES + warp * B_BYTES) >> 4);
        const uint64_t da = desc(sa, 128, 256), db = desc(sb, 128, 256);
        const long long t0 = clock64();
        for (int i = 0; i < n_mma; i += 16) {
            if (elect()) {
#pragma unroll
                for (int j = 0; j < 16; ++j) mma_ss<KIND>(tmem + (uint32_t)warp * N, da, db, idesc, (i | j) ? 1u : 0u);
            }
            __syncwarp();
        }
        if (elect()) commit(smem_u32(&bar[warp]));
        __syncwarp();
        mbar_wait(smem_u32(&bar[warp]), 0);
        if (lane == 0) s_cyc[warp] = clock64() - t0;
    }
    fence_before();
    __syncthreads();
    if (threadIdx.x == 0) { long long m = 0; for (int i = 0; i < n_iss; ++i) m = s_cyc[i] > m ? s_cyc[i] : m; cycles[blockIdx.x] = m; }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int KIND, int N>
static void run_multi(int sms, long long* d_cyc, int n_iss) {
    const int n_mma = 50000;
    const size_t smem = 4 * (128 * 32 + (size_t)N * 32);
    CK(cudaFuncSetAttribute(k_rate_multi<KIND, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<long long> hc(sms);
    k_rate_multi<KIND, N><<<sms, 128, smem>>>(n_mma, n_iss, d_cyc);
    CK(cudaDeviceSynchronize());
    k_rate_multi<KIND, N><<<sms, 128, smem>>>(n_mma, n_iss, d_cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hc.data(), d_cyc, sms * 8, cudaMemcpyDeviceToHost));
    printf("%s N=%3d, %d issuing warps: %6.1f cycles per MMA overall (floor %3.0f)\n", KIND == 0 ? "i8 " : "f16", N, n_iss, (double)hc[0] / ((double)n_mma * n_iss), 128.0 * N / 256.0);
}

template <int KIND, int N>
static void run(int sms, long long* d_cyc, int n_acc, int n_a, int n_b) {
    const int n_mma = 100000;
    const size_t smem = (size_t)n_a * 128 * 32 + (size_t)n_b * N * 32;
    CK(cudaFuncSetAttribute(k_rate<KIND, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<long long> hc(sms);
    k_rate<KIND, N><<<sms, 128, smem>>>(n_mma, n_acc, n_a, n_b, d_cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0)); k_rate<KIND, N><<<sms, 128, smem>>>(n_mma, n_acc, n_a, n_b, d_cyc); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaMemcpy(hc.data(), d_cyc, sms * 8, cudaMemcpyDeviceToHost));
    const double k_per = KIND == 0 ? 32 : 16;
    const double tops = 2.0 * n_mma * 128.0 * N * k_per * sms / (ms * 1e-3) / 1e12;
    const double floor_cyc = 128.0 * N / 256.0;
    printf("%s N=%3d acc=%d A-tiles=%2d B-tiles=%2d: %6.1f cycles per MMA (floor %3.0f), operand bytes per cycle %5.1f, %6.0f T%s/s on %d SMs\n", KIND == 0 ? "i8 " : "f16", N, n_acc,
           n_a, n_b, (double)hc[0] / n_mma, floor_cyc, (128.0 * 32 + N * 32.0) / ((double)hc[0] / n_mma), tops, KIND == 0 ? "OP" : "FLOP", sms);
}

int main() {
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    long long* d_cyc; CK(cudaMalloc(&d_cyc, sms * 8));
    for (int n_iss : {1, 2, 4}) run_multi<0, 64>(sms, d_cyc, n_iss);
    for (int n_iss : {1, 2, 4}) run_multi<0, 128>(sms, d_cyc, n_iss);
    for (int n_iss : {1, 2}) run_multi<0, 256>(sms, d_cyc, n_iss);
    run_multi<0, 8>(sms, d_cyc, 1);
    run<0, 64>(sms, d_cyc, 7, 7, 7);
    run<0, 64>(sms, d_cyc, 1, 7, 7);
    run<0, 64>(sms, d_cyc, 7, 1, 1);
    run<0, 64>(sms, d_cyc, 1, 1, 1);
    run<0, 64>(sms, d_cyc, 2, 1, 7);
    run<0, 64>(sms, d_cyc, 2, 7, 1);
    run<0, 128>(sms, d_cyc, 1, 1, 1);
    run<0, 128>(sms, d_cyc, 3, 7, 7);
    run<0, 128>(sms, d_cyc, 4, 4, 4);
    run<0, 256>(sms, d_cyc, 1, 1, 1);
    run<0, 256>(sms, d_cyc, 2, 4, 4);
    run<0, 32>(sms, d_cyc, 7, 7, 7);
    run<0, 16>(sms, d_cyc, 7, 7, 7);
    run<1, 64>(sms, d_cyc, 7, 7, 7);
    run<1, 64>(sms, d_cyc, 1, 1, 1);
    run<1, 128>(sms, d_cyc, 3, 7, 7);
    run<1, 256>(sms, d_cyc, 1, 1, 1);
    run<1, 256>(sms, d_cyc, 2, 4, 4);
    return 0;
}
