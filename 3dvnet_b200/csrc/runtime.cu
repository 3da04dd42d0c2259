// Error reporting, launch accounting and ABI version of lib3dvnet_b200.
#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace dv3d {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("DV3D_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

cudaError_t func_smem_once(std::atomic<unsigned long long>& mask, const void* func, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
    // The bit is published only AFTER the attribute call returned: a second host thread on the same
    // device either sees it set (attribute in place) or takes the mutex and waits for the first one.
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) mask.fetch_or(bit, std::memory_order_release);
    return e;
}

}  // namespace dv3d

extern "C" const char* dv3d_last_error(void) { return dv3d::g_err; }
extern "C" int dv3d_abi_version(void) { return 4; }
extern "C" long long dv3d_launch_count(void) { return dv3d::g_launches.load(std::memory_order_relaxed); }
