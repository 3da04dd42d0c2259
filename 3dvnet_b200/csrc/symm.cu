// Symmetric regions for a scene that spans GPUs (BASELINE config C4, SURVEY.md §8e "shard UNet by voxel-id
// range").  Every rank allocates the same buffer and opens its peers' copies (CUDA IPC, done by the host side:
// 3dvnet_b200/parallel.py); registering the region here makes every GEMM / pair-reduce epilogue whose output lies
// inside it store its finished rows into all copies, so the all-gather of a row-sharded layer costs no extra pass
// over the data and overlaps the tiles still being computed.  dv3d_symm_barrier is the cross-GPU barrier between
// layers: release/acquire flags at system scope in the same region.
#include <string.h>

#include <mutex>

#include "gemm.cuh"

namespace dv3d {

struct SymmRegion {
    const char* base;
    size_t bytes;
    char* peers[kMaxPeers];
    int n_peers;
};
static SymmRegion g_regions[16];
static int g_n_regions = 0;
static std::mutex g_symm_mu;

static thread_local const int* t_halo_rows = nullptr;
static thread_local long long t_halo_base = 0;

void symm_set_halo(const int* peer_rows, long long row_base) {
    t_halo_rows = peer_rows;
    t_halo_base = row_base;
}

void symm_attach(GemmDesc& d) {
    d.n_peers = 0;
    d.peer_rows = nullptr;
    d.row_base = 0;
    if (g_n_regions == 0) return;
    std::lock_guard<std::mutex> lock(g_symm_mu);
    const char* o = reinterpret_cast<const char*>(d.out);
    for (int i = 0; i < g_n_regions; ++i) {
        const SymmRegion& r = g_regions[i];
        if (o >= r.base && o < r.base + r.bytes) {
            for (int p = 0; p < r.n_peers; ++p) d.peer_out[p] = reinterpret_cast<float*>(r.peers[p] + (o - r.base));
            d.n_peers = r.n_peers;
            d.peer_rows = t_halo_rows;
            d.row_base = t_halo_base;
            return;
        }
    }
}

// thread p signals peer p and waits for peer p's signal; flags[j] on a rank is written by rank j only
// wait_mask (device, optional): bit r set = wait for rank r.  Every peer is always signalled; a rank that only gathers
// rows of some peers (halo exchange of the row-sharded U-Net) need not wait for the others.
__global__ void symm_barrier_kernel(unsigned* local_flags, SymmRegion peers, int rank, unsigned epoch, int* err,
                                    const int* __restrict__ wait_mask) {
    pdl_wait();
    const int p = threadIdx.x;
    if (p >= peers.n_peers) return;
    __threadfence_system();
    unsigned* remote = reinterpret_cast<unsigned*>(peers.peers[p]) + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    // peer p's slot in MY flags: the host passes the peers in ascending rank order with this rank left out
    const int src_rank = p < rank ? p : p + 1;
    if (wait_mask && !((__ldg(wait_mask) >> src_rank) & 1)) return;
    const long long t0 = clock64();
    unsigned seen;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(local_flags + src_rank) : "memory");
        if (clock64() - t0 > 40000000000ll) {  // ~20 s: a peer died. Fail loudly instead of hanging the GPU or
            *err = 1;                           // letting later layers read rows that never arrived
            __threadfence_system();
            __trap();
        }
    } while ((int)(seen - epoch) < 0);
    __threadfence_system();
}

}  // namespace dv3d

using namespace dv3d;

// One zeroed allocation of its own (an IPC handle exports a whole cudaMalloc block) and its 64-byte handle.
extern "C" int dv3d_symm_alloc(size_t bytes, void** ptr, void* handle64) {
    DV3D_REQUIRE(bytes > 0 && ptr && handle64, "symm_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    void* p = nullptr;
    DV3D_CUDA(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(p);
        DV3D_CUDA(e);
    }
    memcpy(handle64, &h, 64);
    *ptr = p;
    return DV3D_OK;
}

// Maps a peer process's allocation into the CURRENT device's address space (peer access is enabled on demand).
extern "C" int dv3d_symm_open(const void* handle64, void** ptr) {
    DV3D_REQUIRE(handle64 && ptr, "symm_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    DV3D_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DV3D_OK;
}

extern "C" int dv3d_symm_close(void* ptr) {
    DV3D_REQUIRE(ptr, "symm_close: null pointer");
    DV3D_CUDA(cudaIpcCloseMemHandle(ptr));
    return DV3D_OK;
}

extern "C" int dv3d_symm_free(void* ptr) {
    DV3D_REQUIRE(ptr, "symm_free: null pointer");
    DV3D_CUDA(cudaFree(ptr));
    return DV3D_OK;
}

extern "C" int dv3d_symm_register(const void* base, size_t bytes, void* const* peer_bases, int n_peers) {
    DV3D_REQUIRE(base && bytes > 0 && peer_bases && n_peers >= 1 && n_peers <= kMaxPeers,
                 "symm_register: bad arguments (1..%d peers)", kMaxPeers);
    DV3D_REQUIRE(((uintptr_t)base & 15) == 0, "symm_register: base must be 16-byte aligned");
    std::lock_guard<std::mutex> lock(g_symm_mu);
    DV3D_REQUIRE(g_n_regions < 16, "symm_register: too many regions");
    SymmRegion& r = g_regions[g_n_regions];
    r.base = reinterpret_cast<const char*>(base);
    r.bytes = bytes;
    r.n_peers = n_peers;
    for (int p = 0; p < n_peers; ++p) {
        DV3D_REQUIRE(peer_bases[p] && ((uintptr_t)peer_bases[p] & 15) == 0, "symm_register: bad peer base %d", p);
        r.peers[p] = reinterpret_cast<char*>(peer_bases[p]);
    }
    ++g_n_regions;
    return DV3D_OK;
}

extern "C" int dv3d_symm_unregister(const void* base) {
    std::lock_guard<std::mutex> lock(g_symm_mu);
    for (int i = 0; i < g_n_regions; ++i)
        if (g_regions[i].base == base) {
            g_regions[i] = g_regions[--g_n_regions];
            return DV3D_OK;
        }
    set_error("symm_unregister: region not registered");
    return DV3D_EINVAL;
}

extern "C" int dv3d_symm_barrier_masked(void* local_flags, void* const* peer_flags, int n_peers, int rank, int epoch,
                                        int* err_flag, const int* wait_mask_dev, void* stream) {
    DV3D_REQUIRE(local_flags && peer_flags && err_flag && n_peers >= 1 && n_peers <= kMaxPeers && rank >= 0 && rank <= n_peers,
                 "symm_barrier: bad arguments");
    SymmRegion peers = {};
    peers.n_peers = n_peers;
    for (int p = 0; p < n_peers; ++p) peers.peers[p] = reinterpret_cast<char*>(peer_flags[p]);
    DV3D_LAUNCH((symm_barrier_kernel), 1, 32, 0, (cudaStream_t)stream, reinterpret_cast<unsigned*>(local_flags), peers, rank,
                (unsigned)epoch, err_flag, wait_mask_dev);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_symm_barrier(void* local_flags, void* const* peer_flags, int n_peers, int rank, int epoch,
                                 int* err_flag, void* stream) {
    return dv3d_symm_barrier_masked(local_flags, peer_flags, n_peers, rank, epoch, err_flag, nullptr, stream);
}
