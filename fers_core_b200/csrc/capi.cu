// capi.cu -- library-level entry points of libfers_b200 (include/fers_b200.h).
#include <atomic>

#include "common.cuh"

namespace fers {
static thread_local std::string g_error;
static std::atomic<long long> g_launches{0};
void count_launches(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launches_total() { return g_launches.load(std::memory_order_relaxed); }
void set_error(const std::string& msg) { g_error = msg; }
int sm_count() {
    static int cached[MAX_DEVICES] = {};
    const int dev = current_device();
    if (cached[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = n > 0 ? n : 148;
    }
    return cached[dev];
}
}  // namespace fers

extern "C" {

const char* fers_b200_version(void) { return "fers_b200 0.1.0 (sm_100a, abi 1)"; }
const char* fers_b200_last_error(void) { return fers::g_error.c_str(); }

int fers_b200_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
    // cudaGetDeviceProperties costs milliseconds: query once per device (an Engine is built per analysis)
    static cudaDeviceProp cache[64];
    static bool have[64] = {false};
    FERS_REQUIRE(device >= 0 && device < 64, "device_info: device index out of range");
    if (!have[device]) {
        FERS_CK(cudaGetDeviceProperties(&cache[device], device));
        have[device] = true;
    }
    const cudaDeviceProp& p = cache[device];
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    FERS_REQUIRE(p.major == 10, "libfers_b200 is built for sm_100a only; device is not compute capability 10.x");
    return 0;
}

long long fers_b200_launch_count(void) { return fers::launches_total(); }

int fers_b200_set_device(int device) {
    FERS_CK(cudaSetDevice(device));
    return 0;
}

int fers_b200_zero(void* ptr, size_t bytes, void* stream) {
    if (bytes == 0) return 0;
    FERS_CK(cudaMemsetAsync(ptr, 0, bytes, fers::as_stream(stream)));
    return 0;
}

}  // extern "C"
