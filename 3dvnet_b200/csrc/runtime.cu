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

// ---------------------------------------------------------------- mailbox read-back (common.cuh)
constexpr int kMailItems = 16, kMailItemBytes = 64;
struct MailArgs {
    const unsigned* src[kMailItems];
    int words[kMailItems];
    int n;
    unsigned* box;   // mapped host memory: word 0 = flag, item i at word 16 (i + 1)
    unsigned seq;
};
__global__ void __launch_bounds__(kMailItems * kMailItemBytes / 4)
mailbox_kernel(const __grid_constant__ MailArgs a) {
    pdl_wait();
    const int item = threadIdx.x >> 4, w = threadIdx.x & 15;
    if (item < a.n && w < a.words[item]) {
        volatile unsigned* dst = a.box + 16 * (item + 1) + w;
        *dst = a.src[item][w];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        *(volatile unsigned*)a.box = a.seq;
        __threadfence_system();
    }
}

struct Mailbox {
    unsigned* host = nullptr;
    unsigned seq = 0;
    ~Mailbox() {
        if (host) cudaFreeHost(host);
    }
};

static bool mailbox_enabled() {
    static const bool on = [] {
        const char* e = getenv("DV3D_MAILBOX");
        return !(e && e[0] == '0');
    }();
    return on;
}

int read_back(const ReadItem* items, int n_items, cudaStream_t st) {
    DV3D_REQUIRE(items && n_items >= 1 && n_items <= kMailItems, "read_back: 1..%d items", kMailItems);
    for (int i = 0; i < n_items; ++i)
        DV3D_REQUIRE(items[i].src && items[i].dst && items[i].bytes > 0 && items[i].bytes <= kMailItemBytes &&
                         items[i].bytes % 4 == 0,
                     "read_back: item %d must be 4..%d bytes, a multiple of 4", i, kMailItemBytes);
    static thread_local Mailbox mb;
    if (mailbox_enabled() && !mb.host) {
        void* p = nullptr;
        if (cudaHostAlloc(&p, (kMailItems + 1) * kMailItemBytes, cudaHostAllocPortable | cudaHostAllocMapped) == cudaSuccess) {
            mb.host = (unsigned*)p;
            memset(p, 0, (kMailItems + 1) * kMailItemBytes);
        } else {
            cudaGetLastError();
        }
    }
    if (!mailbox_enabled() || !mb.host) {
        for (int i = 0; i < n_items; ++i)
            DV3D_CUDA(cudaMemcpyAsync(items[i].dst, items[i].src, items[i].bytes, cudaMemcpyDeviceToHost, st));
        DV3D_CUDA(cudaStreamSynchronize(st));
        return DV3D_OK;
    }
    MailArgs a = {};
    a.n = n_items;
    for (int i = 0; i < n_items; ++i) {
        a.src[i] = (const unsigned*)items[i].src;
        a.words[i] = items[i].bytes / 4;
    }
    void* dev_box = nullptr;
    DV3D_CUDA(cudaHostGetDevicePointer(&dev_box, mb.host, 0));
    a.box = (unsigned*)dev_box;
    a.seq = ++mb.seq ? mb.seq : ++mb.seq;  // never 0
    DV3D_LAUNCH((mailbox_kernel), 1, kMailItems * kMailItemBytes / 4, 0, st, a);
    DV3D_LAUNCHED();
    volatile unsigned* flag = mb.host;
    unsigned long long spins = 0;
    while (*flag != a.seq) {
        if ((++spins & 0x3fff) == 0) {
            // a faulted kernel never raises the flag: surface the error instead of spinning forever
            cudaError_t e = cudaStreamQuery(st);
            if (e != cudaSuccess && e != cudaErrorNotReady) {
                set_error("read_back: %s", cudaGetErrorString(e));
                return DV3D_ECUDA;
            }
            if (e == cudaSuccess && *flag != a.seq && spins > (1ull << 24)) {
                set_error("read_back: the stream finished without raising the mailbox flag");
                return DV3D_ECUDA;
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    for (int i = 0; i < n_items; ++i) memcpy(items[i].dst, (const void*)(mb.host + 16 * (i + 1)), items[i].bytes);
    return DV3D_OK;
}

}  // namespace dv3d

extern "C" const char* dv3d_last_error(void) { return dv3d::g_err; }
extern "C" int dv3d_abi_version(void) { return 4; }
extern "C" long long dv3d_launch_count(void) { return dv3d::g_launches.load(std::memory_order_relaxed); }
