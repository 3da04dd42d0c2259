// Shared helpers for lib3dvnet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dv3d.h"

namespace dv3d {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define DV3D_REQUIRE(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            dv3d::set_error(__VA_ARGS__);       \
            return DV3D_EINVAL;                 \
        }                                       \
    } while (0)

#define DV3D_CUDA(expr)                                                                        \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            dv3d::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return DV3D_ECUDA;                                                                 \
        }                                                                                      \
    } while (0)

// after every kernel launch: counts it and surfaces launch-configuration errors
#define DV3D_LAUNCHED()                      \
    do {                                     \
        dv3d::count_launch();                \
        DV3D_CUDA(cudaGetLastError());       \
    } while (0)

// ---------------------------------------------------------------- programmatic dependent launch
// The hot path is a chain of ~250 short dependent kernels per reference view.  Every kernel is
// launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts with
// griddepcontrol.wait: the next launch is set up and its CTAs scheduled while the previous
// kernel drains, and the wait returns once that kernel has completed and flushed its writes.
// DV3D_PDL=0 in the environment launches without the attribute (A/B measurements).
// (An explicit griddepcontrol.launch_dependents at kernel entry was measured and is not used: the
// dependents' CTAs then spin next to the running kernel and the step got 1.5 % slower.)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();
// Function attributes (cudaFuncSetAttribute) are per device: sets the dynamic shared-memory opt-in of `func` the
// first time a call site runs on the current device.  `mask` is a static std::atomic<unsigned long long> of the
// call site (bit = device), published only after the attribute is in place (safe for concurrent host threads).
cudaError_t func_smem_once(std::atomic<unsigned long long>& mask, const void* func, int bytes);
#define DV3D_FUNC_SMEM_ONCE(mask, func, bytes) DV3D_CUDA(dv3d::func_smem_once(mask, (const void*)(func), (int)(bytes)))

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// kernel names with template commas are passed in parentheses
#define DV3D_LAUNCH(kernel, grid, block, smem, stream, ...) \
    (void)dv3d::launch_pdl(kernel, grid, block, smem, stream, __VA_ARGS__)

// ---------------------------------------------------------------- small device -> host read-backs
// The data-dependent sizes of the path (bounding box, voxel / level / tile counts) are a few bytes each.  A
// cudaMemcpyAsync + cudaStreamSynchronize round trip costs 25-40 us of idle GPU per read on this path; read_back
// instead enqueues one tiny kernel that copies the items into a pinned, mapped host "mailbox" and raises a
// sequence flag, and the host spins on that flag (a PCIe write + a cache miss, ~5 us).  Up to 16 items of at
// most 64 bytes per call, bytes multiples of 4; every calling host thread owns its mailbox.  DV3D_MAILBOX=0
// selects the memcpy + synchronize path (A/B measurements).  Returns DV3D_OK / DV3D_ECUDA (error text set).
struct ReadItem {
    const void* src;  // device
    void* dst;        // host
    int bytes;
};
int read_back(const ReadItem* items, int n_items, cudaStream_t st);

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// torch.linspace(start, end, steps) in fp32 (symmetric formula of ATen's linspace kernel)
__device__ __forceinline__ float linspace_torch(float start, float end, int steps, int i) {
    if (steps == 1) return start;
    float step = (end - start) / (float)(steps - 1);
    return (i < steps / 2) ? start + step * (float)i : end - step * (float)(steps - i - 1);
}

// order-preserving float <-> uint mapping for atomicMin/atomicMax on floats
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ---------------------------------------------------------------- sparse coordinate hash
// A level's table maps key(b,x,y,z) -> row.  Open addressing, linear probing, capacity a
// power of two >= 2 n.  Memory: uint64 keys[cap] then int32 rows[cap].
constexpr unsigned long long kEmptyKey = ~0ull;
constexpr int kCoordBias = 1 << 15;

__host__ __device__ __forceinline__ unsigned long long coord_key(int b, int x, int y, int z) {
    // ordered by (b, z, y, x): the order of the reference's voxel ids (utils.py:45-48)
    return ((((unsigned long long)(unsigned)b << 16 | (unsigned)(z + kCoordBias)) << 16 |
             (unsigned)(y + kCoordBias)) << 16) | (unsigned)(x + kCoordBias);
}
__host__ __device__ __forceinline__ bool coord_in_range(int b, int x, int y, int z) {
    return b >= 0 && b < 65535 && x >= -kCoordBias && x < kCoordBias && y >= -kCoordBias && y < kCoordBias &&
           z >= -kCoordBias && z < kCoordBias;
}
__device__ __forceinline__ unsigned hash_mix(unsigned long long k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (unsigned)k;
}
struct HashView {
    unsigned long long* keys;
    int* rows;
    unsigned mask;  // cap - 1
};
static inline size_t hash_capacity_for(long long n) {
    size_t cap = 64;
    while (cap < (size_t)(2 * n)) cap <<= 1;
    return cap;
}
static inline bool hash_view(void* table, size_t bytes, HashView* v) {
    size_t cap = bytes / 12;
    if (cap < 64 || (cap & (cap - 1))) return false;
    v->keys = (unsigned long long*)table;
    v->rows = (int*)((char*)table + cap * 8);
    v->mask = (unsigned)(cap - 1);
    return true;
}
__device__ __forceinline__ int hash_find(const HashView& t, unsigned long long key) {
    unsigned slot = hash_mix(key) & t.mask;
    while (true) {
        unsigned long long k = t.keys[slot];
        if (k == key) return t.rows[slot];
        if (k == kEmptyKey) return -1;
        slot = (slot + 1) & t.mask;
    }
}

}  // namespace dv3d
