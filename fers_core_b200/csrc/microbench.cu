// microbench.cu -- live roofline denominators measured by the library itself (bench.py):
// register-resident fp64 throughput of the tensor pipe (DMMA.8x8x4) and of the vector pipe (DFMA).
// MEASURED_PEAKS.json carries HBM and bf16 numbers only; the fp64 peak the Cholesky GEMMs are
// reported against is measured here, on the same device and clocks as the benchmark.
#include "common.cuh"

namespace fers {

__global__ void __launch_bounds__(256) k_peak_dmma(double* out, int iters) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
    const double a = 1.0 + threadIdx.x * 1e-6, b = 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_peak_dfma(double* out, int iters) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double x = 1.0000001, y = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace fers

using namespace fers;

extern "C" int fers_b200_fp64_peak(double min_ms, double* dmma_tflops, double* dfma_tflops, double* scratch, size_t scratch_bytes,
                                   void* stream) {
    FERS_REQUIRE(dmma_tflops && dfma_tflops && scratch, "fp64_peak: bad arguments");
    const int blocks = sm_count() * 4, threads = 256;
    FERS_REQUIRE(scratch_bytes >= (size_t)blocks * threads * sizeof(double), "fp64_peak: scratch too small (need 8 * 1024 * SMs bytes)");
    cudaStream_t st = as_stream(stream);
    cudaEvent_t e0, e1;
    FERS_CK(cudaEventCreate(&e0));
    FERS_CK(cudaEventCreate(&e1));
    for (int which = 0; which < 2; ++which) {
        int iters = 2000;
        double best = 0.0;
        for (int rep = 0; rep < 6; ++rep) {
            FERS_CK(cudaEventRecord(e0, st));
            if (which == 0) k_peak_dmma<<<blocks, threads, 0, st>>>(scratch, iters);
            else k_peak_dfma<<<blocks, threads, 0, st>>>(scratch, iters);
            FERS_CK(cudaEventRecord(e1, st));
            FERS_CK(cudaEventSynchronize(e1));
            float ms = 0;
            FERS_CK(cudaEventElapsedTime(&ms, e0, e1));
            const double per_thread_iter = which == 0 ? 8 * 2.0 * 8 * 8 * 4 / 32.0 : 16 * 2.0;
            const double tf = per_thread_iter * (double)threads * blocks * iters / (ms * 1e-3) / 1e12;
            if (ms >= min_ms && tf > best) best = tf;
            if (ms < min_ms) iters = (int)(iters * (min_ms / (ms > 1e-3 ? ms : 1e-3)) * 1.2) + 1;   // lengthen until the burst lasts min_ms
        }
        *(which == 0 ? dmma_tflops : dfma_tflops) = best;
    }
    count_launches(12);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}
