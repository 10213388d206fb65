// common.cuh -- error plumbing and small device helpers shared by the libfers_b200 translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/fers_b200.h"
#include "fers_elem.cuh"

namespace fers {

void set_error(const std::string& msg);

#define FERS_CK(expr)                                                                               \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            fers::set_error(std::string(#expr) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + \
                            ":" + std::to_string(__LINE__) + ")");                                  \
            return 1;                                                                               \
        }                                                                                           \
    } while (0)

#define FERS_LAUNCH_CHECK() FERS_CK(cudaGetLastError())

#define FERS_REQUIRE(cond, msg)          \
    do {                                 \
        if (!(cond)) {                   \
            fers::set_error(msg);        \
            return 2;                    \
        }                                \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }
int sm_count();   // SMs of the current device (cached)
void count_launches(long long n);   // bookkeeping for fers_b200_launch_count()

// One-time set-up that CUDA keeps per device (cudaFuncSetAttribute, streams, events) is keyed by the CURRENT device,
// so engines on several GPUs of one process each get theirs.  Doing the set-up twice is harmless (two host threads
// racing on the same device both set the same attribute).
constexpr int MAX_DEVICES = 64;
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < MAX_DEVICES) ? dev : 0;
}
struct PerDeviceOnce {
    volatile bool flag[MAX_DEVICES] = {};
    bool needed() const { return !flag[current_device()]; }
    void done() { flag[current_device()] = true; }
};

constexpr unsigned FULL = 0xffffffffu;

// 128-bit read-only load of two doubles
__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

}  // namespace fers
