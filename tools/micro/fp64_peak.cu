// fp64 throughput microbenchmark on sm_100a: vector DFMA vs DMMA (mma.sync f64) shapes.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
    double a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double x = 1.0000001, y = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma884(double* out, int iters) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
    double a = 1.0 + threadIdx.x * 1e-6, b = 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma16816(double* out, int iters) {
    double c[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = i - j;
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-6 + i;
    for (int i = 0; i < 4; ++i) b[i] = 1e-3 * i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma1684(double* out, int iters) {
    double c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = i - j;
    double a[2] = {1.0 + threadIdx.x * 1e-6, 0.5}, b = 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static void run(const char* name, F launch, double flop_per_thread_iter, int threads, int blocks, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(iters / 10);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        launch(iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double flops = flop_per_thread_iter * (double)threads * blocks * iters;
    printf("%-12s threads/CTA %4d  %8.3f ms  %7.2f TFLOP/s  (%s)\n", name, threads, best, flops / best / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sm * 8 * 1024);
    const int iters = 20000;
    for (int threads : {256, 512, 1024}) {
        const int blocks = sm * (1024 / threads);
        run("dfma", [&](int it) { k_dfma<<<blocks, threads>>>(out, it); }, 16 * 2.0, threads, blocks, iters);
        run("dmma884", [&](int it) { k_dmma884<<<blocks, threads>>>(out, it); }, 8 * 2.0 * 8 * 8 * 4 / 32.0, threads, blocks, iters);
        run("dmma1684", [&](int it) { k_dmma1684<<<blocks, threads>>>(out, it); }, 8 * 2.0 * 16 * 8 * 4 / 32.0, threads, blocks, iters);
        run("dmma16816", [&](int it) { k_dmma16816<<<blocks, threads>>>(out, it); }, 4 * 2.0 * 16 * 8 * 16 / 32.0, threads, blocks, iters);
    }
    return 0;
}
