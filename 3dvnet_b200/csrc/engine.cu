// Engine: the whole hot path enqueued from native code by one call (dv3d_hot_path).
//
// The reference drives this path from Python, one library call per tensor op
// (eval-3dvnet.py:58-99 -> lightningmodel.py:124-242).  On a B200 the kernels of the path run
// 5-50 us each, so ~220 interpreter round trips per reference view would bound the step; here
// the same per-op entry points of this library are called back to back from C++ with a bump
// allocator over a caller-provided arena.  The sequence below mirrors, call for call, the
// Python-composed path in 3dvnet_b200/mv3d (which stays as the reference-shaped interface and
// as the parity cross-check: tests assert bit-identical depth).
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"
#include "gemm.cuh"

namespace dv3d {

// ------------------------------------------------------------------ stage timing (optional)
// CUDA events recorded on the launching stream around the stages of dv3d_hot_path, so that a
// benchmark can attribute device time inside the one native call (bench.py's roofline block).
struct ProfRec {
    int id;
    cudaEvent_t a, b;
};
static bool g_prof = false;
static std::mutex g_prof_mu;  // hot-path calls may come from several host threads (one stream each)
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_pool;
static size_t g_pool_used = 0;

static cudaEvent_t prof_event() {  // g_prof_mu held
    if (g_pool_used == g_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        g_pool.push_back(e);
    }
    return g_pool[g_pool_used++];
}
struct Prof {
    cudaStream_t st;
    cudaEvent_t b;
    bool on;
    Prof(int id, cudaStream_t s) : st(s), b(nullptr), on(g_prof) {
        if (!on) return;
        cudaEvent_t a;
        {
            std::lock_guard<std::mutex> lock(g_prof_mu);
            a = prof_event();
            b = prof_event();
            g_recs.push_back(ProfRec{id, a, b});
        }
        cudaEventRecord(a, st);
    }
    ~Prof() {
        if (on) cudaEventRecord(b, st);
    }
};

// ------------------------------------------------------------------ side stream
// PointNet (per point) and the coordinate levels / kernel maps / pair plans (per voxel) both
// depend only on the voxelisation: they run concurrently, the latter on a per-device side stream
// (it contains the host read-backs of the level sizes, which then overlap PointNet's kernels).
struct SideStream {
    cudaStream_t main;  // the caller's stream this side stream belongs to
    int device;
    cudaStream_t stream;
    cudaEvent_t fork, join;
};
// one side stream per (device, caller stream): concurrent dv3d_hot_path calls on different
// streams (one host thread each) must not share fork / join events
static SideStream* side_stream(cudaStream_t main) {
    static SideStream table[32] = {};
    static int n_used = 0;
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < n_used; ++i)
        if (table[i].main == main && table[i].device == dev) return &table[i];
    if (n_used == 32) return nullptr;
    SideStream& s = table[n_used];
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming);
    s.main = main;
    s.device = dev;
    ++n_used;
    return &s;
}

// ------------------------------------------------------------------ arena
struct Arena {
    char* base;
    size_t cap, off;
    bool failed;
    template <typename T>
    T* get(size_t count) {
        size_t bytes = align_up(count * sizeof(T), 256);
        if (off + bytes > cap) {
            failed = true;
            return nullptr;
        }
        T* p = reinterpret_cast<T*>(base + off);
        off += bytes;
        return p;
    }
};

#define ARENA_CHECK(a)                                                                                        \
    do {                                                                                                      \
        if ((a).failed) {                                                                                     \
            set_error("hot_path: workspace too small (%zu bytes given, more than %zu needed)", (a).cap, (a).off); \
            return DV3D_ENOSPC;                                                                               \
        }                                                                                                     \
    } while (0)
#define TRY(expr)                 \
    do {                          \
        int _rc = (expr);         \
        if (_rc != DV3D_OK) return _rc; \
    } while (0)

constexpr size_t kBitmapShare = 64ull << 20;  // voxelize / coarsen bitmaps and scans

// pts_batch[r*P + p] = depth_batch[r]  (lightningmodel.py:171-172)
__global__ void expand_batch_kernel(const long long* __restrict__ depth_batch, int P, long long n, long long* __restrict__ out) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = depth_batch[i / P];
}
// depth += offset  (eval-3dvnet.py:99)
__global__ void add_inplace_kernel(float* __restrict__ x, const float* __restrict__ y, long long n) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = x[i] + y[i];
}


// Halo exchange of the row-sharded U-Net.  Rows are sorted by voxel id, so the rows of level l a rank gathers through
// its kernel maps (its own rows' 3x3x3 neighbours, parents and children) form a band around its own range: each rank
// takes the min / max input row over its maps per level ...
struct HaloMaps {
    const int* nbr[3 * DV3D_MAX_LEVELS];
    long long n[3 * DV3D_MAX_LEVELS];   // entries (rows x 27)
    int level[3 * DV3D_MAX_LEVELS];     // level of the map's INPUT rows
};
__global__ void __launch_bounds__(256)
halo_needs_kernel(HaloMaps hm, int* __restrict__ lo, int* __restrict__ hi) {
    pdl_wait();
    const int* nbr = hm.nbr[blockIdx.y];
    const long long n = hm.n[blockIdx.y];
    int mn = 0x7fffffff, mx = -1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int v = __ldg(nbr + i);
        if (v >= 0) mn = min(mn, v), mx = max(mx, v);
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx >= 0) {
        atomicMin(lo + hm.level[blockIdx.y], mn);
        atomicMax(hi + hm.level[blockIdx.y], mx + 1);
    }
}
// ... and writes them into every peer's table (first 256 bytes of the heap, behind the barrier flags): entry
// [level][index of this rank among the destination's peers] = (lo, hi).  The barrier in front of the first layer
// publishes them.  A producer then stores row g of a level-l output into peer p only when lo <= g < hi.
// which ranks send this rank rows of a level: those whose row range meets [lo, hi).  The barrier behind a layer then
// waits for them only (dv3d_symm_barrier_masked): a slow layer on one rank delays its neighbours, not everybody.
struct RankRows {
    int b[DV3D_MAX_LEVELS][kMaxPeers + 2];   // rank q owns rows [b[l][q], b[l][q + 1]) of level l
};
__global__ void halo_wait_mask_kernel(const int* __restrict__ lo, const int* __restrict__ hi, RankRows rr, int n_levels, int world,
                                      int rank, int* __restrict__ mask) {
    pdl_wait();
    const int l = threadIdx.x;
    if (l >= n_levels) return;
    int m = 0;
    for (int q = 0; q < world; ++q)
        if (q != rank && lo[l] < hi[l] && lo[l] < rr.b[l][q + 1] && hi[l] > rr.b[l][q]) m |= 1 << q;
    mask[l] = m;
}
constexpr int kHaloTabOff = 64;   // bytes; 3 levels x 7 peers x 2 ints = 168 bytes
struct PeerPtrs {
    char* p[kMaxPeers];
};
// the rank's point rows into every peer's heap in one launch: all NVLink ports at once (seven peer memcpys in a row took
// ~100 us of copy-engine time per exchange at 8 ranks; the stores of one kernel overlap)
struct PushSeg {
    const float4* src;   // local heap
    size_t off16;        // offset inside the heap, in 16-byte units
    size_t n16;
};
__global__ void __launch_bounds__(256)
push_rows_kernel(PushSeg a, PushSeg b, PeerPtrs peers, int n_peers) {
    pdl_wait();
    const size_t total = a.n16 + b.n16;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const bool first = i < a.n16;
        const size_t k = first ? i : i - a.n16;
        const float4 v = first ? a.src[k] : b.src[k];
        const size_t at = (first ? a.off16 : b.off16) + k;
#pragma unroll
        for (int p = 0; p < kMaxPeers; ++p)
            if (p < n_peers) reinterpret_cast<float4*>(peers.p[p])[at] = v;
    }
}
__global__ void halo_publish_kernel(const int* __restrict__ lo, const int* __restrict__ hi, int n_levels, int rank, int n_peers,
                                    PeerPtrs peers) {
    pdl_wait();
    const int p = threadIdx.x;
    if (p >= n_peers) return;
    const int dest_rank = p < rank ? p : p + 1;
    const int my_idx = rank < dest_rank ? rank : rank - 1;
    int* tab = reinterpret_cast<int*>(peers.p[p] + kHaloTabOff);
    for (int l = 0; l < n_levels; ++l) {
        tab[(l * kMaxPeers + my_idx) * 2 + 0] = lo[l];
        tab[(l * kMaxPeers + my_idx) * 2 + 1] = hi[l];
    }
}

// Sharded PointNet: a rank runs the per-point layers for the points that fall into ITS voxel rows [r0, r1) - whichever
// rank back-projected them; every rank holds the whole cloud - so the per-voxel max pools (scenemodeling.py:129) are
// complete locally and nothing is exchanged.  The order of the selected points is arbitrary (atomic append): it only
// permutes rows of private activations, and max pooling is order independent.
__global__ void __launch_bounds__(256)
select_points_kernel(const int* __restrict__ seg, long long N, int r0, int r1, int* __restrict__ sel, int* __restrict__ count) {
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool mine = i < N && __ldg(seg + i) >= r0 && __ldg(seg + i) < r1;
    const unsigned m = __ballot_sync(0xffffffffu, mine);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (mine) sel[base + __popc(m & ((1u << lane) - 1u))] = (int)i;
}
// rows [pts - anchor | feat | 0-padding] of the selected points (as pointnet_input_kernel) and their LOCAL voxel row
__global__ void __launch_bounds__(256)
pointnet_input_sel_kernel(const float* __restrict__ pts, const float* __restrict__ pts_feat, int feat_ld,
                          const float* __restrict__ anchor_pts, const int* __restrict__ seg, const int* __restrict__ sel,
                          long long M, int C, int ld, int r0, float* __restrict__ out, int* __restrict__ seg_sel) {
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * ld) return;
    const long long j = i / ld;
    const int c = (int)(i - j * ld);
    const long long p = __ldg(sel + j);
    const int v = __ldg(seg + p);
    float x = 0.f;
    if (c < 3)
        x = __fsub_rn(__ldg(pts + 3 * p + c), __ldg(anchor_pts + 3 * (long long)v + c));
    else if (c < 3 + C)
        x = __ldg(pts_feat + p * feat_ld + (c - 3));
    out[i] = x;
    if (c == 0) seg_sel[j] = v - r0;
}

struct Level {
    int* coords;
    long long n;
    int stride;
    void* table;
    size_t table_bytes;
};

// kernel map + its pair-major plan (csrc/sparse_pairs.cu)
struct KMap {
    int* nbr;
    long long n_out;
    void* plan;
    size_t plan_bytes;
    long long n_tiles;
    bool use_pairs;
    long long row0;              // first output row of this map (sharded scene: the rank's range), n_out rows from there
    const struct Level* out_lv;
    const struct Level* in_lv;
    int step;
};

// A scene that spans GPUs (BASELINE config C4): this rank's share of the sparse U-Net.  Rank r owns rows
// [r m, min((r+1) m, n)) of every level, m = ceil(n / world).  Layer outputs live in the symmetric heap (same offset
// on every rank; the convolution epilogues store finished rows into every peer's copy, csrc/symm.cu) and one flag
// barrier follows every layer whose output the next layer gathers through a kernel map.
struct Shard {
    int rank, world;
    char* heap;                 // local symmetric block, registered with dv3d_symm_register
    size_t heap_bytes, heap_off;
    void* const* peer_heaps;    // ascending rank order, this rank left out
    int n_peers;
    int* err_flag;
    int epoch;                  // last barrier epoch used (strictly increasing over the life of the heap)
};

struct Scene {
    Shard* shard;                    // nullptr: single GPU
    const int* halo_tab;             // sharded: [level][peer][lo, hi) rows each peer needs (device, in the heap header)
    const int* wait_mask;            // sharded: per level, the ranks that send this rank rows (device)
    long long r0[DV3D_MAX_LEVELS], r1[DV3D_MAX_LEVELS];  // this rank's row range per level (whole level when !shard)
    Level lv[DV3D_MAX_LEVELS];
    int n_levels;
    KMap same[DV3D_MAX_LEVELS];      // k3 s1 kernel map of level l
    KMap down[DV3D_MAX_LEVELS - 1];  // rows of level l+1 <- level l
    KMap up[DV3D_MAX_LEVELS - 1];    // rows of level l   <- level l+1
    void* pair_ws;                   // P buffer of the pair-major convolutions
    size_t pair_ws_bytes;
    float* feats[DV3D_MAX_LEVELS];   // decoder inputs (output of the U-Net per level)
    int dims[DV3D_MAX_LEVELS];
    float* origin;                   // [n_batch,3]
    void* split_ws;
    size_t split_ws_bytes;
};

// DV3D_SHARD_BALANCE=1: coarse levels cut by work instead of equal row counts.  Off by default: measured at 8 ranks it
// costs one more read-back per scene model and gains nothing (10.60 vs 10.50 ms per step) - what is left in the barriers
// is latency, not imbalance.
static bool shard_balance() {   // read per call: every rank of a job must see the same value
    const char* e = getenv("DV3D_SHARD_BALANCE");
    return e && e[0] == '1';
}

// DV3D_SHARD_NEIGHBOUR_WAIT=0: every barrier waits for every rank (A/B measurements)
static bool shard_neighbour_wait() {
    const char* e = getenv("DV3D_SHARD_NEIGHBOUR_WAIT");
    return !(e && e[0] == '0');
}

// rows of a layer output: the arena, or - sharded scene - the symmetric heap (same offset on every rank, because every
// rank allocates the same sizes in the same order: the coordinate levels are identical everywhere)
static float* layer_rows(Scene& sc, Arena& ar, long long n, int C) {
    if (!sc.shard) return ar.get<float>((size_t)n * C);
    Shard& sh = *sc.shard;
    const size_t bytes = align_up((size_t)n * C * sizeof(float), 256);
    if (sh.heap_off + bytes > sh.heap_bytes) {
        ar.failed = true;  // reported as ENOSPC by ARENA_CHECK
        return nullptr;
    }
    float* p = reinterpret_cast<float*>(sh.heap + sh.heap_off);
    sh.heap_off += bytes;
    return p;
}

// cross-GPU barrier behind a layer whose output the next layer gathers through a kernel map
// level >= 0: the layer's rows travel by the halo table, wait only for the ranks that send this rank rows of that level
static int layer_barrier(Scene& sc, void* st, int level = -1) {
    if (!sc.shard) return DV3D_OK;
    Shard& sh = *sc.shard;
    Prof pr(DV3D_STAGE_BARRIER, (cudaStream_t)st);
    const int* mask = (level >= 0 && sc.wait_mask && shard_neighbour_wait()) ? sc.wait_mask + level : nullptr;
    return dv3d_symm_barrier_masked(sh.heap, sh.peer_heaps, sh.n_peers, sh.rank, ++sh.epoch, sh.err_flag, mask, st);
}

// One sparse convolution + GroupNorm + ReLU for the rows the kernel map covers (all rows, or this rank's range):
// `feat` and `out` are full [n, C] arrays, `residual` too.  Sharded: the epilogue stores the rows into every rank's
// copy of `out` (csrc/symm.cu) and the barrier follows.
static void set_halo(const Scene& sc, int out_level, bool to_all, long long row0) {
    if (sc.shard) symm_set_halo(to_all ? nullptr : sc.halo_tab + out_level * kMaxPeers * 2, row0);
}

static int sparse_conv(const dv3d_dense_params_t& p, const float* feat, long long n_in, const KMap& km, const float* residual,
                       Scene& sc, float* out, int out_level, bool to_all, void* st) {
    if (km.n_out > 0) {
        // to_all: the level's final features, which every rank samples in its PointFlow passes; otherwise only the
        // ranks whose kernel maps reach a row receive it
        set_halo(sc, out_level, to_all, km.row0);
        const float* res = residual ? residual + km.row0 * p.N : nullptr;
        float* o = out + km.row0 * p.N;
        if (km.use_pairs && p.Wp)
            TRY(dv3d_sparse_conv_pairs(feat, n_in, p.K / 27, km.plan, km.n_tiles, km.n_out, p.Wp, p.N, p.a, p.b, res, 1,
                                       sc.pair_ws, sc.pair_ws_bytes, o, st));
        else
            TRY(dv3d_sparse_conv(feat, n_in, p.K / 27, km.nbr, km.n_out, p.W, p.Wp, p.N, p.a, p.b, res, 1, sc.split_ws,
                                 sc.split_ws_bytes, o, st));
        if (sc.shard) symm_set_halo(nullptr, 0);
    }
    return layer_barrier(sc, st, to_all ? -1 : out_level);
}

// relu(x + GN2(conv2(relu(GN1(conv1(x))))))  (scenemodeling.py:16-44)
static int res_block(const dv3d_dense_params_t (&p)[2], const float* x, long long n, const KMap& nbr, Scene& sc, Arena& ar,
                     float** out, int level, bool to_all, void* st) {
    const int C = p[0].N;
    float* h = layer_rows(sc, ar, n, C);
    float* y = layer_rows(sc, ar, n, C);
    ARENA_CHECK(ar);
    TRY(sparse_conv(p[0], x, n, nbr, nullptr, sc, h, level, false, st));
    TRY(sparse_conv(p[1], h, n, nbr, x, sc, y, level, to_all, st));
    *out = y;
    return DV3D_OK;
}

static int build_level(Level& L, int* err_flag, Arena& ar, void* st) {
    L.table_bytes = dv3d_hash_bytes(L.n);
    L.table = ar.get<char>(L.table_bytes);
    ARENA_CHECK(ar);
    return dv3d_hash_build(L.coords, L.n, L.table, L.table_bytes, err_flag, st);
}

// kernel map and, on the tensor-core path, its pair-major plan (counts are read back later, once)
static int kernel_map(const Level& out_lv, long long row0, long long row1, const Level& in_lv, int step, bool want_plan,
                      Arena& ar, KMap* km, void* st) {
    km->row0 = row0;
    km->n_out = row1 - row0;
    km->nbr = ar.get<int>((size_t)(km->n_out > 0 ? km->n_out : 1) * 27);
    km->plan = nullptr;
    km->n_tiles = 0;
    km->use_pairs = false;
    const size_t pb = want_plan ? dv3d_pair_plan_bytes(km->n_out) : 0;
    if (want_plan) km->plan = ar.get<char>(pb);
    ARENA_CHECK(ar);
    km->plan_bytes = pb;
    km->out_lv = &out_lv;
    km->in_lv = &in_lv;
    km->step = step;
    (void)st;  // the maps of a scene are filled by one batched launch (model_scene)
    return DV3D_OK;
}

// voxelise -> PointNet -> sparse 3D-UNet on the feature-rich point cloud (lightningmodel.py:176-185)
static int model_scene(const dv3d_net_params_t& net, const float* pts, const float* pts_feat, int feat_ld,
                       const long long* pts_batch, long long N, double edge_len, Scene& sc, Arena& ar, void* st) {
    cudaStream_t cs = (cudaStream_t)st;
    SideStream* side = side_stream(cs);
    DV3D_REQUIRE(side, "hot_path: cannot create the side stream (more than 32 caller streams?)");
    // ---- voxelise (utils.py:38-64)
    Prof* pr = new Prof(DV3D_STAGE_VOXELIZE, cs);
    struct ProfGuard {  // closes the open stage on every return path
        Prof*& p;
        ~ProfGuard() { delete p; }
    } guard{pr};
    auto next_stage = [&](int id, cudaStream_t on) {
        delete pr;
        pr = new Prof(id, on);
    };
    dv3d_voxel_grid_t grid;
    void* scratch = ar.get<char>(256);
    ARENA_CHECK(ar);
    TRY(dv3d_voxel_grid(pts, pts_batch, N, (float)edge_len, &grid, scratch, st));
    const size_t vws_bytes = dv3d_voxelize_workspace_bytes(&grid, N);
    void* vws = ar.get<char>(vws_bytes);
    float* a_pts = ar.get<float>((size_t)N * 3);
    int* a_idx = ar.get<int>((size_t)N * 3);
    long long* a_batch = ar.get<long long>((size_t)N);
    int* seg = ar.get<int>((size_t)N);
    ARENA_CHECK(ar);
    long long nv = 0;
    TRY(dv3d_voxelize(pts, pts_batch, N, &grid, vws, vws_bytes, N, &nv, a_pts, a_idx, a_batch, seg, st));

    // fork: the side stream sees the voxelisation
    DV3D_CUDA(cudaEventRecord(side->fork, cs));
    DV3D_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));

    // ---- PointNet (scenemodeling.py:127-144), main stream
    next_stage(DV3D_STAGE_POINTNET, cs);
    const int in_pad = net.pointnet_in_pad, Hd = net.pointnet[0].N;
    // single GPU: all points and voxels; sharded: this rank's voxel rows [v0, v1) and the points inside them
    long long v0 = 0, v1 = nv, Npl = N;
    const int* seg_l = seg;
    int* sel = nullptr;
    if (sc.shard) {
        const long long m = (nv + sc.shard->world - 1) / sc.shard->world;   // as sc.r0 / sc.r1 of level 0 below
        v0 = std::min((long long)sc.shard->rank * m, nv);
        v1 = std::min((long long)(sc.shard->rank + 1) * m, nv);
        sel = ar.get<int>((size_t)N);
        int* count = ar.get<int>(1);
        ARENA_CHECK(ar);
        DV3D_CUDA(cudaMemsetAsync(count, 0, sizeof(int), cs));
        DV3D_LAUNCH((select_points_kernel), cdiv(N, 256), 256, 0, cs, (const int*)seg, N, (int)v0, (int)v1, sel, count);
        DV3D_LAUNCHED();
        int host_count = 0;
        ReadItem it = {count, &host_count, 4};
        TRY(read_back(&it, 1, cs));
        Npl = host_count;
    }
    const long long nvl = v1 - v0;
    float* x0 = ar.get<float>((size_t)Npl * in_pad);
    float* xa = ar.get<float>((size_t)Npl * Hd);
    float* xb = ar.get<float>((size_t)Npl * Hd);
    // the four per-voxel max pools are produced by the epilogues of fc2..fc5 (dv3d_linear_pool): one 0xFF fill for all
    float* pools = ar.get<float>((size_t)4 * nvl * Hd);
    float* F = sc.shard ? nullptr : ar.get<float>((size_t)nv * net.pointnet[5].N);
    ARENA_CHECK(ar);
    if (nvl > 0) DV3D_CUDA(cudaMemsetAsync(pools, 0xFF, sizeof(float) * 4 * (size_t)nvl * Hd, cs));
    if (sc.shard && Npl > 0) {
        int* seg_sel = ar.get<int>((size_t)Npl);
        ARENA_CHECK(ar);
        DV3D_LAUNCH((pointnet_input_sel_kernel), cdiv(Npl * in_pad, 256), 256, 0, cs, pts, pts_feat, feat_ld, (const float*)a_pts,
                    (const int*)seg, (const int*)sel, Npl, 32, in_pad, (int)v0, x0, seg_sel);
        DV3D_LAUNCHED();
        seg_l = seg_sel;
    } else if (Npl > 0) {
        TRY(dv3d_pointnet_input(pts, pts_feat, feat_ld, a_pts, seg, N, 32, in_pad, x0, st));
    }
    if (Npl > 0) {
        TRY(dv3d_linear(x0, in_pad, in_pad, nullptr, nullptr, 0, Npl, net.pointnet[0].W, net.pointnet[0].Wp, net.pointnet[0].b,
                        Hd, 0, xa, st));
        TRY(dv3d_linear_pool(xa, Hd, Hd, nullptr, nullptr, 0, Npl, net.pointnet[1].W, net.pointnet[1].Wp, net.pointnet[1].b, Hd,
                             1, xb, pools, seg_l, st));
        float *cur = xb, *nxt = xa;
        for (int i = 2; i <= 4; ++i) {
            float* pool_in = pools + (size_t)(i - 2) * nvl * Hd;
            float* pool_out = pools + (size_t)(i - 1) * nvl * Hd;
            TRY(dv3d_linear_pool(cur, Hd, Hd, pool_in, seg_l, Hd, Npl, net.pointnet[i].W, net.pointnet[i].Wp, net.pointnet[i].b,
                                 Hd, 1, nxt, pool_out, seg_l, st));
            float* t = cur;
            cur = nxt;
            nxt = t;
        }
    }
    if (!sc.shard)
        TRY(dv3d_linear(pools + (size_t)3 * nv * Hd, Hd, Hd, nullptr, nullptr, 0, nv, net.pointnet[5].W, net.pointnet[5].Wp,
                        net.pointnet[5].b, net.pointnet[5].N, 1, F, st));
    // sharded: the output layer runs behind the join (it stores its rows through the halo table)

    // ---- coordinate levels, hash tables, kernel maps (what ME keeps in its coordinate manager)
    next_stage(DV3D_STAGE_LEVELS, side->stream);
    void* const sst = (void*)side->stream;
    const int nl = net.n_levels;
    sc.n_levels = nl;
    int* err_flag = ar.get<int>(1);
    ARENA_CHECK(ar);
    DV3D_CUDA(cudaMemsetAsync(err_flag, 0, sizeof(int), side->stream));
    sc.lv[0].n = nv;
    sc.lv[0].stride = 1;
    sc.lv[0].coords = ar.get<int>((size_t)nv * 4);
    ARENA_CHECK(ar);
    TRY(dv3d_make_coords(a_idx, a_batch, nv, sc.lv[0].coords, sst));
    TRY(build_level(sc.lv[0], err_flag, ar, sst));
    // every coarser level straight from the finest coordinates (floor(c / 2^l) * 2^l), one sync for all counts
    {
        void* cws[DV3D_MAX_LEVELS] = {};
        const int gx = (int)grid.n_cells[0], gy = (int)grid.n_cells[1], gz = (int)grid.n_cells[2], gb = (int)grid.n_batch;
        size_t cws_bytes[DV3D_MAX_LEVELS] = {};
        int strides[DV3D_MAX_LEVELS] = {};
        int* ccoords[DV3D_MAX_LEVELS] = {};
        size_t cws_total = 0;
        for (int l = 1; l < nl; ++l) {
            strides[l] = 1 << l;
            cws_bytes[l] = (dv3d_coarsen_workspace_bytes(gx, gy, gz, gb, strides[l]) + 255) & ~(size_t)255;
            cws_total += cws_bytes[l];
        }
        if (nl > 1) {
            // one contiguous span, so that a single memset clears every level's bitmap and counters
            char* span = ar.get<char>(cws_total);
            for (int l = 1; l < nl; ++l) {
                cws[l] = span;
                span += cws_bytes[l];
                ccoords[l] = sc.lv[l].coords = ar.get<int>((size_t)nv * 4);
            }
            ARENA_CHECK(ar);
            TRY(dv3d_coarsen_enqueue_batch(sc.lv[0].coords, nv, strides + 1, nl - 1, gx, gy, gz, gb, cws + 1, cws_bytes + 1, nv,
                                           ccoords + 1, sst));
        }
        if (nl > 1) {
            long long n_out[DV3D_MAX_LEVELS] = {};
            TRY(dv3d_coarsen_finish_batch(cws + 1, strides + 1, nl - 1, gx, gy, gz, gb, nv, n_out, sst));
            for (int l = 1; l < nl; ++l) {
                sc.lv[l].n = n_out[l - 1];
                sc.lv[l].stride = 1 << l;
            }
        }
        if (nl > 1) {
            const int* hc[DV3D_MAX_LEVELS];
            long long hn[DV3D_MAX_LEVELS];
            void* ht[DV3D_MAX_LEVELS];
            size_t hb[DV3D_MAX_LEVELS];
            for (int l = 1; l < nl; ++l) {
                Level& L = sc.lv[l];
                L.table_bytes = dv3d_hash_bytes(L.n);
                L.table = ar.get<char>(L.table_bytes);
                hc[l] = L.coords;
                hn[l] = L.n;
                ht[l] = L.table;
                hb[l] = L.table_bytes;
            }
            ARENA_CHECK(ar);
            TRY(dv3d_hash_build_batch(hc + 1, hn + 1, ht + 1, hb + 1, nl - 1, err_flag, sst));
        }
    }

    // every kernel map of the U-Net and its pair-major plan, then ONE sync for the counts
    {
        const bool want_plan = net.res_down[0][0][0].Wp != nullptr;
        DV3D_REQUIRE(!sc.shard || want_plan, "hot_path: the row-sharded U-Net needs a tensor-core GEMM mode (packed weights)");
        RankRows rank_rows = {};         // sharded: every rank's rows of every level (identical on all ranks)
        for (int l = 0; l < nl; ++l) {   // this rank's rows of every level
            const long long n = sc.lv[l].n;
            if (sc.shard) {
                const long long m = (n + sc.shard->world - 1) / sc.shard->world;
                sc.r0[l] = std::min((long long)sc.shard->rank * m, n);
                sc.r1[l] = std::min((long long)(sc.shard->rank + 1) * m, n);
                for (int q = 0; q <= sc.shard->world; ++q) rank_rows.b[l][q] = (int)std::min((long long)q * m, n);
            } else {
                sc.r0[l] = 0;
                sc.r1[l] = n;
            }
        }
        if (sc.shard && nl > 1 && shard_balance()) {
            // coarser levels: equal WORK per rank instead (level 0 keeps equal rows: PointNet, which runs beside this
            // on the main stream, already works on that range, and its finest-level layers are balanced as they are)
            const int world = sc.shard->world;
            int* bounds = ar.get<int>((size_t)(nl - 1) * 16);
            char* scratch[DV3D_MAX_LEVELS] = {};
            for (int l = 1; l < nl; ++l) scratch[l] = ar.get<char>(dv3d_balanced_row_bounds_scratch_bytes(sc.lv[l].n, 8));
            ARENA_CHECK(ar);
            for (int l = 1; l < nl; ++l)
                TRY(dv3d_balanced_row_bounds(sc.lv[l].coords, sc.lv[l].n, sc.lv[l].table, sc.lv[l].table_bytes, sc.lv[l].stride,
                                             world, 8, scratch[l], bounds + (l - 1) * 16, sst));
            int host_bounds[DV3D_MAX_LEVELS][16];
            ReadItem items[DV3D_MAX_LEVELS];
            for (int l = 1; l < nl; ++l) items[l - 1] = ReadItem{bounds + (l - 1) * 16, host_bounds[l], (int)sizeof(int) * (world + 1)};
            TRY(read_back(items, nl - 1, side->stream));
            for (int l = 1; l < nl; ++l) {
                sc.r0[l] = host_bounds[l][sc.shard->rank];
                sc.r1[l] = host_bounds[l][sc.shard->rank + 1];
                for (int q = 0; q <= world; ++q) rank_rows.b[l][q] = host_bounds[l][q];
                DV3D_REQUIRE(sc.r0[l] >= 0 && sc.r0[l] <= sc.r1[l] && sc.r1[l] <= sc.lv[l].n, "hot_path_sharded: bad row bounds");
            }
        }
        KMap* maps[3 * DV3D_MAX_LEVELS];
        int n_maps = 0;
        for (int l = 0; l < nl; ++l) {
            TRY(kernel_map(sc.lv[l], sc.r0[l], sc.r1[l], sc.lv[l], sc.lv[l].stride, want_plan, ar, &sc.same[l], sst));
            if (sc.same[l].n_out > 0) maps[n_maps++] = &sc.same[l];  // a tail rank may own no row of a small level
        }
        for (int l = 0; l + 1 < nl; ++l) {
            TRY(kernel_map(sc.lv[l + 1], sc.r0[l + 1], sc.r1[l + 1], sc.lv[l], sc.lv[l].stride, want_plan, ar, &sc.down[l], sst));
            if (sc.down[l].n_out > 0) maps[n_maps++] = &sc.down[l];
        }
        for (int l = 0; l + 1 < nl; ++l) {
            TRY(kernel_map(sc.lv[l], sc.r0[l], sc.r1[l], sc.lv[l + 1], -sc.lv[l].stride, want_plan, ar, &sc.up[l], sst));
            if (sc.up[l].n_out > 0) maps[n_maps++] = &sc.up[l];
        }
        {
            const int* co[3 * DV3D_MAX_LEVELS];
            long long no[3 * DV3D_MAX_LEVELS];
            const void* tb[3 * DV3D_MAX_LEVELS];
            size_t tbb[3 * DV3D_MAX_LEVELS];
            int stp[3 * DV3D_MAX_LEVELS];
            int* nb[3 * DV3D_MAX_LEVELS];
            for (int i = 0; i < n_maps; ++i) {
                co[i] = maps[i]->out_lv->coords + 4 * maps[i]->row0;
                no[i] = maps[i]->n_out;
                tb[i] = maps[i]->in_lv->table;
                tbb[i] = maps[i]->in_lv->table_bytes;
                stp[i] = maps[i]->step;
                nb[i] = maps[i]->nbr;
            }
            TRY(dv3d_kernel_map_batch(co, no, tb, tbb, stp, nb, n_maps, sst));
        }
        if (sc.shard) {
            // the rows of every level this rank gathers, into every peer's halo table (published by the barrier in
            // front of the first layer)
            Shard& sh = *sc.shard;
            int* lohi = ar.get<int>(2 * DV3D_MAX_LEVELS);
            ARENA_CHECK(ar);
            int *lo = lohi, *hi = lohi + DV3D_MAX_LEVELS;
            DV3D_CUDA(cudaMemsetAsync(lo, 0x7f, sizeof(int) * DV3D_MAX_LEVELS, side->stream));
            DV3D_CUDA(cudaMemsetAsync(hi, 0, sizeof(int) * DV3D_MAX_LEVELS, side->stream));
            if (n_maps > 0) {
                HaloMaps hm = {};
                long long max_n = 0;
                for (int i = 0; i < n_maps; ++i) {
                    hm.nbr[i] = maps[i]->nbr;
                    hm.n[i] = maps[i]->n_out * 27;
                    hm.level[i] = (int)(maps[i]->in_lv - sc.lv);
                    max_n = std::max(max_n, hm.n[i]);
                }
                const unsigned gx = (unsigned)std::min<long long>(cdiv(max_n, 256 * 8), 148 * 4);
                DV3D_LAUNCH((halo_needs_kernel), dim3(gx, n_maps), 256, 0, side->stream, hm, lo, hi);
                DV3D_LAUNCHED();
            }
            PeerPtrs pp = {};
            for (int p = 0; p < sh.n_peers; ++p) pp.p[p] = reinterpret_cast<char*>(sh.peer_heaps[p]);
            DV3D_LAUNCH((halo_publish_kernel), 1, 32, 0, side->stream, (const int*)lo, (const int*)hi, nl, sh.rank, sh.n_peers, pp);
            DV3D_LAUNCHED();
            sc.halo_tab = reinterpret_cast<const int*>(sh.heap + kHaloTabOff);
            int* mask = ar.get<int>(DV3D_MAX_LEVELS);
            ARENA_CHECK(ar);
            DV3D_LAUNCH((halo_wait_mask_kernel), 1, 32, 0, side->stream, (const int*)lo, (const int*)hi, rank_rows, nl, sh.world, sh.rank,
                        mask);
            DV3D_LAUNCHED();
            sc.wait_mask = mask;
        }
        sc.pair_ws = nullptr;
        sc.pair_ws_bytes = 0;
        if (want_plan) {
            const void* plans[3 * DV3D_MAX_LEVELS];
            void* plans_w[3 * DV3D_MAX_LEVELS];
            const int* nbrs[3 * DV3D_MAX_LEVELS];
            long long n_outs[3 * DV3D_MAX_LEVELS], tiles[3 * DV3D_MAX_LEVELS];
            size_t pbytes[3 * DV3D_MAX_LEVELS];
            for (int i = 0; i < n_maps; ++i) {
                plans[i] = plans_w[i] = maps[i]->plan;
                nbrs[i] = maps[i]->nbr;
                n_outs[i] = maps[i]->n_out;
                pbytes[i] = maps[i]->plan_bytes;
            }
            TRY(dv3d_pair_plan_build(nbrs, n_outs, plans_w, pbytes, n_maps, sst));
            TRY(dv3d_pair_plan_counts(plans, n_maps, tiles, nullptr, sst));
            long long max_tiles = 0;
            for (int i = 0; i < n_maps; ++i) {
                maps[i]->n_tiles = tiles[i];
                maps[i]->use_pairs = dv3d_sparse_conv_prefers_pairs(maps[i]->n_out, tiles[i]) != 0;
                if (maps[i]->use_pairs && tiles[i] > max_tiles) max_tiles = tiles[i];
            }
            sc.pair_ws_bytes = dv3d_sparse_conv_pairs_workspace_bytes(max_tiles, 128);
            sc.pair_ws = ar.get<char>(sc.pair_ws_bytes + 256);
            ARENA_CHECK(ar);
        }
    }

    // join: the U-Net needs PointNet's features (main stream) and the maps / plans (side stream)
    DV3D_CUDA(cudaEventRecord(side->join, side->stream));
    DV3D_CUDA(cudaStreamWaitEvent(cs, side->join, 0));

    // ---- sparse U-Net (scenemodeling.py:191-237)
    next_stage(DV3D_STAGE_UNET, cs);
    float* xs[DV3D_MAX_LEVELS];
    float* x = F;
    if (sc.shard) {
        // this barrier publishes the halo tables (and no rank stores into a heap its owner still reads); then PointNet's
        // output layer (scenemodeling.py:141-144) for this rank's voxel rows, delivered like any level-0 layer
        TRY(layer_barrier(sc, st));
        const int Cf = net.pointnet[5].N;
        F = layer_rows(sc, ar, nv, Cf);
        ARENA_CHECK(ar);
        DV3D_REQUIRE(sc.r0[0] == v0 && sc.r1[0] == v1, "hot_path_sharded: PointNet and U-Net disagree on the rank's voxel rows");
        if (nvl > 0) {
            set_halo(sc, 0, nl == 1 && net.n_res[0] == 0, v0);
            TRY(dv3d_linear(pools + (size_t)3 * nvl * Hd, Hd, Hd, nullptr, nullptr, 0, nvl, net.pointnet[5].W, net.pointnet[5].Wp,
                            net.pointnet[5].b, Cf, 1, F + v0 * Cf, st));
            symm_set_halo(nullptr, 0);
        }
        TRY(layer_barrier(sc, st, (nl == 1 && net.n_res[0] == 0) ? -1 : 0));
        x = F;
    }
    for (int b = 0; b < net.n_res[0]; ++b)
        TRY(res_block(net.res_down[0][b], x, sc.lv[0].n, sc.same[0], sc, ar, &x, 0, nl == 1 && b == net.n_res[0] - 1, st));
    xs[0] = x;
    for (int i = 1; i < nl; ++i) {
        float* y = layer_rows(sc, ar, sc.lv[i].n, net.down[i - 1].N);
        ARENA_CHECK(ar);
        const bool deepest = i == nl - 1;
        TRY(sparse_conv(net.down[i - 1], x, sc.lv[i - 1].n, sc.down[i - 1], nullptr, sc, y, i, deepest && net.n_res[i] == 0, st));
        x = y;
        for (int b = 0; b < net.n_res[i]; ++b)
            TRY(res_block(net.res_down[i][b], x, sc.lv[i].n, sc.same[i], sc, ar, &x, i, deepest && b == net.n_res[i] - 1, st));
        xs[i] = x;
    }
    sc.feats[nl - 1] = xs[nl - 1];
    sc.dims[nl - 1] = net.res_down[nl - 1][0][0].N;
    for (int i = 0; i < nl - 1; ++i) {
        const int l = nl - 2 - i;  // target (finer) level
        const int Cu = net.up[i].N, Cx = net.feat_adj[i].K - Cu, Ca = net.feat_adj[i].N;
        const long long r0 = sc.r0[l], n_loc = sc.r1[l] - sc.r0[l];
        // the transposed convolution feeds the 1x1 'feature adjust' row by row: its output stays local to the rank
        float* up = ar.get<float>((size_t)(n_loc > 0 ? n_loc : 1) * Cu);
        float* adj = layer_rows(sc, ar, sc.lv[l].n, Ca);
        ARENA_CHECK(ar);
        if (n_loc > 0) {
            const KMap& km = sc.up[l];
            const dv3d_dense_params_t& p = net.up[i];
            if (km.use_pairs && p.Wp)
                TRY(dv3d_sparse_conv_pairs(x, sc.lv[l + 1].n, p.K / 27, km.plan, km.n_tiles, n_loc, p.Wp, p.N, p.a, p.b, nullptr, 1,
                                           sc.pair_ws, sc.pair_ws_bytes, up, st));
            else
                TRY(dv3d_sparse_conv(x, sc.lv[l + 1].n, p.K / 27, km.nbr, n_loc, p.W, p.Wp, p.N, p.a, p.b, nullptr, 1, sc.split_ws,
                                     sc.split_ws_bytes, up, st));
            set_halo(sc, l, net.n_res[l] == 0, r0);
            TRY(dv3d_concat_linear_gn_relu(up, Cu, xs[l] + r0 * Cx, Cx, n_loc, net.feat_adj[i].W, net.feat_adj[i].Wp, Ca,
                                           net.feat_adj[i].a, net.feat_adj[i].b, adj + r0 * Ca, st));
            if (sc.shard) symm_set_halo(nullptr, 0);
        }
        TRY(layer_barrier(sc, st, net.n_res[l] == 0 ? -1 : l));
        x = adj;
        // the reversed n_res list: res_up[i] has n_res[nl-2-i] blocks (scenemodeling.py:168-175)
        for (int b = 0; b < net.n_res[l]; ++b)
            TRY(res_block(net.res_up[i][b], x, sc.lv[l].n, sc.same[l], sc, ar, &x, l, b == net.n_res[l] - 1, st));
        sc.feats[l] = x;
        sc.dims[l] = Ca;
    }
    // position of index (0,0,0) of every batch (scenemodeling.py:211-226 / refinement.py:33)
    sc.origin = ar.get<float>((size_t)grid.n_batch * 3);
    ARENA_CHECK(ar);
    DV3D_CUDA(cudaMemsetAsync(sc.origin, 0, sizeof(float) * grid.n_batch * 3, cs));
    TRY(dv3d_batch_origin(a_pts, a_idx, a_batch, nv, (float)edge_len, sc.origin, st));
    return DV3D_OK;
}


// everything a PointFlow pass needs besides the scene model (lightningmodel.py:187-242)
struct PassCtx {
    const float* feats_nhwc;
    int n_imgs, Hf, Wf;
    const float* cams;
    const int *ref_img, *edge_rowptr, *edge_src;
    float* depth;            // [n_ref, h, w], updated in place
    int n_ref, h, w, H, W;
    long long Np;
    int in_dim, var_off;
    float *operand, *dec_a, *dec_b, *pts_hyp, *offs;
    const long long* pts_batch;
    void* split_ws;
    size_t split_ws_bytes;
    double edge_len;
    void* stream;
};

static int pointflow_pass(const dv3d_net_params_t& net, const PassCtx& c, const Scene& sc, double offset) {
    cudaStream_t cs = (cudaStream_t)c.stream;
    const float* feats_nhwc = c.feats_nhwc;
    const int n_imgs = c.n_imgs, Hf = c.Hf, Wf = c.Wf, n_ref = c.n_ref, h = c.h, w = c.w, H = c.H, W = c.W;
    const float* cams = c.cams;
    const int *ref_img = c.ref_img, *edge_rowptr = c.edge_rowptr, *edge_src = c.edge_src;
    float *depth = c.depth, *operand = c.operand, *dec_a = c.dec_a, *dec_b = c.dec_b, *pts_hyp = c.pts_hyp, *offs = c.offs;
    const long long* pts_batch = c.pts_batch;
    const long long Np = c.Np;
    const int in_dim = c.in_dim, var_off = c.var_off;
    void* split_ws = c.split_ws;
    const size_t split_ws_bytes = c.split_ws_bytes;
    const double edge_len = c.edge_len;
    void* stream = c.stream;
    // PointFlow pass (lightningmodel.py:187-242)
    {
        Prof pr(DV3D_STAGE_FLOW_WARP, cs);
        TRY(dv3d_points_var(feats_nhwc, n_imgs, 32, Hf, Wf, cams, ref_img, edge_rowptr, edge_src, depth, n_ref, h, w,
                            H, W, 3, offset, pts_hyp, operand, 8, in_dim, var_off, stream));
    }
    int off_c = 0;
    {
        Prof pr(DV3D_STAGE_FLOW_INTERP, cs);
        // all levels with one launch, finest level first in the operand (refinement.py:41 prepends)
        float res_l[DV3D_MAX_LEVELS];
        int stride_l[DV3D_MAX_LEVELS], C_l[DV3D_MAX_LEVELS], off_l[DV3D_MAX_LEVELS];
        const void* tab_l[DV3D_MAX_LEVELS];
        size_t tabb_l[DV3D_MAX_LEVELS];
        const float* feat_l[DV3D_MAX_LEVELS];
        for (int l = 0; l < sc.n_levels; ++l) {
            res_l[l] = (float)(sc.lv[l].stride * edge_len);
            stride_l[l] = sc.lv[l].stride;
            tab_l[l] = sc.lv[l].table;
            tabb_l[l] = sc.lv[l].table_bytes;
            feat_l[l] = sc.feats[l];
            C_l[l] = sc.dims[l];
            off_l[l] = off_c;
            off_c += sc.dims[l];
        }
        TRY(dv3d_sparse_interp_batch(pts_hyp, pts_batch, Np, 7, 8, sc.origin, sc.n_levels, res_l, stride_l, tab_l,
                                     tabb_l, feat_l, C_l, off_l, operand, in_dim, stream));
    }
    DV3D_REQUIRE(off_c == var_off, "hot_path: decoder input width %d != level widths %d + 32", in_dim, off_c);
    if (net.dec_fused[0] && net.dec_fused[1] && net.dec_fused[2]) {
        // the whole decoder + `depth += offset` as one tcgen05 kernel (csrc/decoder_fused.cu)
        Prof pr(DV3D_STAGE_DEC_GEMM0, cs);
        const void* wp[3] = {net.dec_fused[0], net.dec_fused[1], net.dec_fused[2]};
        const float* sc[3] = {net.dec[0].a, net.dec[1].a, net.dec[2].a};
        const float* sh[3] = {net.dec[0].b, net.dec[1].b, net.dec[2].b};
        TRY(dv3d_decoder_fused(operand, Np, 8, in_dim, in_dim, wp, sc, sh, net.dec[0].N, net.dec_head_weight,
                               net.dec_head_bias, offset, dv3d_get_gemm_precision() & 0xff, nullptr, nullptr, depth,
                               stream));
    } else {
        {
            Prof pr(DV3D_STAGE_DEC_GEMM0, cs);
            TRY(dv3d_conv1d_bn_relu(operand, Np, 8, in_dim, in_dim, net.dec[0].W, net.dec[0].Wp, net.dec[0].a,
                                    net.dec[0].b, net.dec[0].N, dec_a, net.dec[0].N, split_ws, split_ws_bytes, stream));
        }
        {
            Prof pr(DV3D_STAGE_DEC_REST, cs);
            TRY(dv3d_conv1d_bn_relu(dec_a, Np, 8, net.dec[1].K / 3, net.dec[0].N, net.dec[1].W, net.dec[1].Wp,
                                    net.dec[1].a, net.dec[1].b, net.dec[1].N, dec_b, net.dec[1].N, split_ws, split_ws_bytes, stream));
            TRY(dv3d_conv1d_bn_relu(dec_b, Np, 8, net.dec[2].K / 3, net.dec[1].N, net.dec[2].W, net.dec[2].Wp,
                                    net.dec[2].a, net.dec[2].b, net.dec[2].N, dec_a, net.dec[2].N, split_ws, split_ws_bytes, stream));
            TRY(dv3d_decoder_head(dec_a, Np, 7, 8, net.dec[2].N, net.dec[2].N, net.dec_head_weight, net.dec_head_bias,
                                  offset, nullptr, offs, stream));
            DV3D_LAUNCH((add_inplace_kernel), cdiv(Np, 256), 256, 0, cs, depth, offs, Np);
            DV3D_LAUNCHED();
        }
    }
    return DV3D_OK;
}

}  // namespace dv3d

using namespace dv3d;

extern "C" int dv3d_engine_profile(int enable) {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_prof = enable != 0;
    g_recs.clear();
    g_pool_used = 0;
    return DV3D_OK;
}

extern "C" int dv3d_engine_profile_read(int* ids, float* ms, int cap) {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    int n = 0;
    for (const ProfRec& r : g_recs) {
        if (n >= cap) break;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) {
            cudaGetLastError();
            set_error("engine_profile_read: events not complete - synchronise the stream first");
            return DV3D_ECUDA;
        }
        ids[n] = r.id;
        ms[n] = t;
        ++n;
    }
    return n;
}

extern "C" size_t dv3d_hot_path_workspace_bytes(const dv3d_net_params_t* net, int n_imgs, int n_ref, int D, int h, int w) {
    if (!net || n_imgs <= 0 || n_ref <= 0 || D <= 0 || h <= 0 || w <= 0) return 0;
    const size_t P = (size_t)h * w, Np = (size_t)n_ref * P, vol = (size_t)D * P;
    size_t floats = 0;
    floats += (size_t)n_imgs * 36;
    floats += (size_t)n_ref * 32 * vol;  // x_var
    // CostRegNet activations: 8 + 8 (full res), 16 x3 (1/8 voxels), 32 x3 (1/64), 64 x2 (1/512)
    floats += (size_t)n_ref * vol * (8 + 8) + (size_t)n_ref * vol / 8 * 48 + (size_t)n_ref * vol / 64 * 96 +
              (size_t)n_ref * vol / 512 * 128 + 4096;
    floats += 2 * Np;  // depth, offsets
    size_t per_outer = 0;
    per_outer += Np * (3 + 32 + 2);                               // pts, feat, pts_batch (int64)
    per_outer += Np * (3 + 3 + 2 + 1);                            // anchors
    per_outer += Np * (net->pointnet_in_pad + 6 * 128 + 64);      // PointNet (x0, xa, xb, 4 pools, F)
    // levels: coords 4, table <= 12 * 4 n + 768, kernel maps 27 x (3 same + 2 down + 2 up) per voxel
    per_outer += Np * 3 * (4 + 12) + Np * 27 * 7 + 3 * 1024;
    // pair-major plans (7 maps: pair_in <= 27 n + 27*128, pair_slot 27 n, tile ids) and the P buffer of
    // the largest map that takes the pair path (at most 0.6 * 27 * n rows of 128 channels)
    per_outer += 7 * (Np * 27 * 2 + Np * 27 / 128 + 27 * 130 + 256) + (Np * 27 * 6 / 10 + 27 * 128) * 128;
    // U-Net features: 2 outputs per residual block + down / up / adj, at most 128 channels each
    int blocks = 0;
    for (int l = 0; l < net->n_levels; ++l) blocks += net->n_res[l] * (l == net->n_levels - 1 ? 1 : 2);
    per_outer += Np * 128 * (size_t)(2 * blocks + 3 * (net->n_levels - 1));
    floats += per_outer;  // the arena is rewound after every outer iteration
    floats += Np * 8 * (size_t)net->dec[0].K / 3 + 2 * Np * 8 * 128 + Np * 7 * 3;  // decoder operand, activations, points
    return floats * sizeof(float) + kBitmapShare + dv3d_sparse_conv_workspace_bytes(128) + (64 << 10) /* alignment slack */;
}

// The whole pass for `n_ref` reference views.  With a Shard these are the rank's contiguous range
// [ref_start, ref_start + n_ref) of the scene's `n_ref_total` references (depth_batch then holds all n_ref_total
// entries): the rank's point rows go into the symmetric heap of every rank (peer copies over NVLink + one flag
// barrier - the exchange step of SURVEY.md section 8e), every rank voxelises the whole cloud, and the sparse U-Net runs
// row-sharded (model_scene).
static int hot_path_impl(const dv3d_net_params_t& net, const float* feats_nhwc, int n_imgs, int Hf, int Wf,
                         const float* rotmats, const float* tvecs, const float* K, const int* ref_img,
                         const int* edge_rowptr, const int* edge_src, int n_ref, const long long* depth_batch,
                         double depth_start, double depth_interval, int D, int h, int w, int H, int W, double edge_len,
                         const double* offsets_host, int n_outer, int n_inner, void* workspace, size_t workspace_bytes,
                         Shard* sh, int n_ref_total, int ref_start, float* depth_init_out, float* depth_out, void* stream) {
    cudaStream_t cs = (cudaStream_t)stream;
    Arena ar{(char*)workspace, workspace_bytes, 0, false};
    const long long P = (long long)h * w, Np = (long long)n_ref * P;
    const long long Ng = (long long)n_ref_total * P, row0 = (long long)ref_start * P;  // the whole cloud, this rank's first row

    // ================= path A: cost volume -> CostRegNet -> soft-argmin (mvsnet.py:176-229)
    float* cams = ar.get<float>((size_t)n_imgs * 36);
    ARENA_CHECK(ar);
    TRY(dv3d_camera_tables(rotmats, tvecs, K, n_imgs, cams, stream));
    float* depth = nullptr;
    float* offs = nullptr;
    if (n_ref > 0) {
        float* x_var = ar.get<float>((size_t)n_ref * 32 * D * P);
        ARENA_CHECK(ar);
        {
            Prof pr(DV3D_STAGE_PLANESWEEP, cs);
            TRY(dv3d_planesweep_var(feats_nhwc, n_imgs, 32, Hf, Wf, cams, ref_img, edge_rowptr, edge_src, n_ref, depth_start,
                                    depth_interval, D, h, w, H, W, x_var, stream));
        }
        float* act[10];
        int aD[10], aH[10], aW[10];
        {
            Prof pr(DV3D_STAGE_COSTREG, cs);
            const float* in = x_var;
            int cd = D, ch = h, cw = w;
            for (int i = 0; i < 10; ++i) {
                const dv3d_conv3d_params_t& c = net.costreg[i];
                int od = cd, oh = ch, ow = cw;
                if (c.kind == 1) od = (cd + 1) / 2, oh = (ch + 1) / 2, ow = (cw + 1) / 2;
                if (c.kind == 2) od = 2 * cd, oh = 2 * ch, ow = 2 * cw;
                act[i] = ar.get<float>((size_t)n_ref * c.Cout * od * oh * ow);
                ARENA_CHECK(ar);
                // skips: x = conv4 + conv7(x); x = conv2 + conv8(x); x = conv0 + conv9(x)  (mvsnet.py:159-161)
                const float* skip = i == 7 ? act[4] : i == 8 ? act[2] : i == 9 ? act[0] : nullptr;
                if (c.kind == 2)
                    TRY(dv3d_deconv3d_bn_relu(in, n_ref, c.Cin, cd, ch, cw, c.weight, c.scale, c.shift, c.Cout, skip, act[i], stream));
                else
                    TRY(dv3d_conv3d_bn_relu(in, n_ref, c.Cin, cd, ch, cw, c.weight, c.scale, c.shift, c.Cout, c.kind == 1 ? 2 : 1,
                                            skip, act[i], stream));
                aD[i] = od, aH[i] = oh, aW[i] = ow;
                in = act[i];
                cd = od, ch = oh, cw = ow;
            }
            DV3D_REQUIRE(aD[9] == D && aH[9] == h && aW[9] == w, "hot_path: CostRegNet does not return to full resolution");
        }
        depth = ar.get<float>((size_t)Np);
        offs = ar.get<float>((size_t)Np);
        ARENA_CHECK(ar);
        const double depth_end = depth_start + depth_interval * (D - 1);
        {
            Prof pr(DV3D_STAGE_SOFTARGMIN, cs);
            TRY(dv3d_prob_softargmin(act[9], n_ref, net.costreg[9].Cout, D, h, w, net.prob_weight, net.prob_bias,
                                     (float)depth_start, (float)depth_end, nullptr, depth, stream));
        }
        if (depth_init_out) DV3D_CUDA(cudaMemcpyAsync(depth_init_out, depth, sizeof(float) * Np, cudaMemcpyDeviceToDevice, cs));
    }

    // ================= path B: volumetric refinement (eval-3dvnet.py:73-99)
    if (n_outer * n_inner > 0) {
        ar.off = 0;  // path A's volumes are dead; keep only cams / depth / offs by re-reserving them first
        float* cams2 = ar.get<float>((size_t)n_imgs * 36);
        (void)cams2;  // same address as cams
        // depth / offs live behind the volumes: move them to the front of the arena
        float* depth_f = ar.get<float>((size_t)Np);
        float* offs_f = ar.get<float>((size_t)Np);
        if (Np > 0) DV3D_CUDA(cudaMemcpyAsync(depth_f, depth, sizeof(float) * Np, cudaMemcpyDeviceToDevice, cs));
        depth = depth_f;
        offs = offs_f;
        const int in_dim = net.dec[0].K / 3, var_off = in_dim - 32;
        float* operand = ar.get<float>((size_t)Np * 8 * in_dim);  // [n_pts, 8, in_dim], padding row stays zero
        const bool fused_dec = net.dec_fused[0] && net.dec_fused[1] && net.dec_fused[2];
        float* dec_a = fused_dec ? nullptr : ar.get<float>((size_t)Np * 8 * net.dec[0].N);
        float* dec_b = fused_dec ? nullptr : ar.get<float>((size_t)Np * 8 * net.dec[1].N);
        float* pts_hyp = ar.get<float>((size_t)Np * 7 * 3);
        long long* pts_batch_all = ar.get<long long>((size_t)Ng);  // batch index of every point of the cloud
        const size_t split_ws_bytes = dv3d_sparse_conv_workspace_bytes(128);
        void* split_ws = ar.get<char>(split_ws_bytes);
        ARENA_CHECK(ar);
        const PassCtx pc = {feats_nhwc, n_imgs, Hf, Wf, cams, ref_img, edge_rowptr, edge_src, depth, n_ref, h, w, H, W, Np, in_dim,
                            var_off, operand, dec_a, dec_b, pts_hyp, offs, pts_batch_all + row0, split_ws, split_ws_bytes, edge_len,
                            stream};
        if (Np > 0) DV3D_CUDA(cudaMemsetAsync(operand, 0, sizeof(float) * Np * 8 * in_dim, cs));
        DV3D_CUDA(cudaMemsetAsync(split_ws, 0, dv3d_sparse_conv_workspace_bytes(128), cs));
        DV3D_LAUNCH((expand_batch_kernel), cdiv(Ng, 256), 256, 0, cs, depth_batch, (int)P, Ng, pts_batch_all);
        DV3D_LAUNCHED();
        // sharded: the cloud lives at fixed offsets of the symmetric heap, the layer outputs behind it
        float *pts_sym = nullptr, *feat_sym = nullptr;
        size_t layers_off = 0;
        if (sh) {
            const size_t pts_bytes = align_up((size_t)Ng * 3 * sizeof(float), 256), feat_bytes = align_up((size_t)Ng * 32 * sizeof(float), 256);
            DV3D_REQUIRE(256 + pts_bytes + feat_bytes <= sh->heap_bytes, "hot_path_sharded: the symmetric heap is too small for the point cloud");
            pts_sym = reinterpret_cast<float*>(sh->heap + 256);
            feat_sym = reinterpret_cast<float*>(sh->heap + 256 + pts_bytes);
            layers_off = 256 + pts_bytes + feat_bytes;
        }
        const size_t mark = ar.off;
        for (int o = 0; o < n_outer; ++o) {
            ar.off = mark;
            // feature-rich point cloud (lightningmodel.py:132-174)
            float* pts = sh ? pts_sym : ar.get<float>((size_t)Np * 3);
            float* pfeat = sh ? feat_sym : ar.get<float>((size_t)Np * 32);
            ARENA_CHECK(ar);
            if (Np > 0) {
                Prof pr(DV3D_STAGE_POINTCLOUD, cs);
                TRY(dv3d_points_var(feats_nhwc, n_imgs, 32, Hf, Wf, cams, ref_img, edge_rowptr, edge_src, depth, n_ref, h, w, H,
                                    W, 0, 0.0, pts + 3 * row0, pfeat + 32 * row0, 1, 32, 0, stream));
            }
            if (sh) {
                // all-gather of the point rows: this rank's rows into every peer's copy, then the barrier.  A peer is
                // past every reader of its copy of the cloud: it has arrived at the last barrier of the previous
                // scene model, which follows its voxelisation and PointNet in stream order.
                Prof pr(DV3D_STAGE_EXCHANGE, cs);
                if (Np > 0) {
                    const float* p_src = pts + 3 * row0;
                    const float* f_src = pfeat + 32 * row0;
                    DV3D_REQUIRE((((uintptr_t)p_src | (uintptr_t)f_src) & 15) == 0 && (3 * Np) % 4 == 0,
                                 "hot_path_sharded: point rows must be 16-byte aligned (h * w a multiple of 4)");
                    PushSeg a = {reinterpret_cast<const float4*>(p_src), (size_t)((const char*)p_src - sh->heap) / 16, (size_t)(3 * Np) / 4};
                    PushSeg b = {reinterpret_cast<const float4*>(f_src), (size_t)((const char*)f_src - sh->heap) / 16, (size_t)(32 * Np) / 4};
                    PeerPtrs pp = {};
                    for (int p = 0; p < sh->n_peers; ++p) pp.p[p] = reinterpret_cast<char*>(sh->peer_heaps[p]);
                    DV3D_LAUNCH((push_rows_kernel), 2 * kNumSMs, 256, 0, cs, a, b, pp, sh->n_peers);
                    DV3D_LAUNCHED();
                }
                TRY(dv3d_symm_barrier(sh->heap, sh->peer_heaps, sh->n_peers, sh->rank, ++sh->epoch, sh->err_flag, stream));
                sh->heap_off = layers_off;
            }
            Scene sc;
            memset(&sc, 0, sizeof(sc));
            sc.shard = sh;
            sc.split_ws = split_ws;
            sc.split_ws_bytes = dv3d_sparse_conv_workspace_bytes(128);
            TRY(model_scene(net, pts, pfeat, 32, pts_batch_all, Ng, edge_len, sc, ar, stream));
            for (int it = 0; it < n_inner && Np > 0; ++it) {
                TRY(pointflow_pass(net, pc, sc, offsets_host[o * n_inner + it]));
            }
        }
    }
    if (Np > 0) DV3D_CUDA(cudaMemcpyAsync(depth_out, depth, sizeof(float) * Np, cudaMemcpyDeviceToDevice, cs));
    return DV3D_OK;
}

extern "C" int dv3d_hot_path(const dv3d_net_params_t* netp, const float* feats_nhwc, int n_imgs, int Hf, int Wf,
                             const float* rotmats, const float* tvecs, const float* K, const int* ref_img,
                             const int* edge_rowptr, const int* edge_src, int n_ref, const long long* depth_batch,
                             double depth_start, double depth_interval, int D, int h, int w, int H, int W, double edge_len,
                             const double* offsets_host, int n_outer, int n_inner, void* workspace, size_t workspace_bytes,
                             float* depth_init_out, float* depth_out, void* stream) {
    DV3D_REQUIRE(netp && feats_nhwc && rotmats && tvecs && K && ref_img && edge_rowptr && edge_src && depth_batch && depth_out,
                 "hot_path: null pointer");
    DV3D_REQUIRE(workspace && ((uintptr_t)workspace & 255) == 0, "hot_path: workspace must be a 256-byte aligned device arena");
    DV3D_REQUIRE(n_outer >= 0 && n_inner >= 0 && (n_outer * n_inner == 0 || offsets_host), "hot_path: bad refinement schedule");
    DV3D_REQUIRE(D % 8 == 0 && h % 8 == 0 && w % 8 == 0,
                 "hot_path: D, h, w must be multiples of 8 (three stride-2 levels of CostRegNet), got %d %d %d", D, h, w);
    DV3D_REQUIRE(netp->n_levels >= 1 && netp->n_levels <= DV3D_MAX_LEVELS, "hot_path: n_levels out of range");
    if (n_ref == 0) return DV3D_OK;
    return hot_path_impl(*netp, feats_nhwc, n_imgs, Hf, Wf, rotmats, tvecs, K, ref_img, edge_rowptr, edge_src, n_ref, depth_batch,
                         depth_start, depth_interval, D, h, w, H, W, edge_len, offsets_host, n_outer, n_inner, workspace,
                         workspace_bytes, nullptr, n_ref, 0, depth_init_out, depth_out, stream);
}

extern "C" size_t dv3d_hot_path_sharded_heap_bytes(const dv3d_net_params_t* net, int n_ref_total, int h, int w) {
    if (!net || n_ref_total <= 0 || h <= 0 || w <= 0) return 0;
    const size_t Ng = (size_t)n_ref_total * h * w;
    int blocks = 0;
    for (int l = 0; l < net->n_levels; ++l) blocks += net->n_res[l] * (l == net->n_levels - 1 ? 1 : 2);
    // flags, the cloud (3 + 32 floats per point), every layer output of the U-Net (at most one voxel per point and
    // 128 channels: two per residual block, one per down / feature-adjust layer), 256 bytes of alignment each
    const size_t layers = (size_t)(2 * blocks + 2 * (net->n_levels - 1));
    // + PointNet's output layer
    return 256 + Ng * 35 * 4 + 512 + (Ng * 128 * 4 + 256) * (layers + 1);
}

extern "C" int dv3d_hot_path_sharded(const dv3d_net_params_t* netp, const float* feats_nhwc, int n_imgs, int Hf, int Wf,
                                     const float* rotmats, const float* tvecs, const float* K, const int* ref_img,
                                     const int* edge_rowptr, const int* edge_src, int n_ref_local, int ref_start,
                                     int n_ref_total, const long long* depth_batch_all, double depth_start,
                                     double depth_interval, int D, int h, int w, int H, int W, double edge_len,
                                     const double* offsets_host, int n_outer, int n_inner, void* workspace,
                                     size_t workspace_bytes, int rank, int world, void* heap, size_t heap_bytes,
                                     void* const* peer_heaps, int* epoch, int* err_flag, float* depth_init_out,
                                     float* depth_out, void* stream) {
    DV3D_REQUIRE(netp && feats_nhwc && rotmats && tvecs && K && depth_batch_all, "hot_path_sharded: null pointer");
    DV3D_REQUIRE(n_ref_local >= 0 && ref_start >= 0 && ref_start + n_ref_local <= n_ref_total && n_ref_total > 0,
                 "hot_path_sharded: reference range [%d, %d) outside 0..%d", ref_start, ref_start + n_ref_local, n_ref_total);
    DV3D_REQUIRE(n_ref_local == 0 || (ref_img && edge_rowptr && edge_src && depth_out), "hot_path_sharded: null pointer");
    DV3D_REQUIRE(workspace && ((uintptr_t)workspace & 255) == 0, "hot_path_sharded: workspace must be a 256-byte aligned device arena");
    DV3D_REQUIRE(n_outer >= 0 && n_inner >= 0 && (n_outer * n_inner == 0 || offsets_host), "hot_path_sharded: bad refinement schedule");
    DV3D_REQUIRE(D % 8 == 0 && h % 8 == 0 && w % 8 == 0,
                 "hot_path_sharded: D, h, w must be multiples of 8 (three stride-2 levels of CostRegNet), got %d %d %d", D, h, w);
    DV3D_REQUIRE(netp->n_levels >= 1 && netp->n_levels <= DV3D_MAX_LEVELS, "hot_path_sharded: n_levels out of range");
    DV3D_REQUIRE(world >= 2 && world <= 8 && rank >= 0 && rank < world, "hot_path_sharded: rank %d of %d (2..8 ranks)", rank, world);
    DV3D_REQUIRE(heap && peer_heaps && epoch && err_flag && ((uintptr_t)heap & 255) == 0,
                 "hot_path_sharded: the symmetric heap (dv3d_symm_alloc + dv3d_symm_register), its peers, epoch and error flag are required");
    DV3D_REQUIRE(netp->res_down[0][0][0].Wp,
                 "hot_path_sharded: needs a tensor-core GEMM mode with packed weights (only those epilogues store into peer memory)");
    Shard sh = {rank, world, (char*)heap, heap_bytes, 256, peer_heaps, world - 1, err_flag, *epoch};
    const int rc = hot_path_impl(*netp, feats_nhwc, n_imgs, Hf, Wf, rotmats, tvecs, K, ref_img, edge_rowptr, edge_src, n_ref_local,
                                 depth_batch_all, depth_start, depth_interval, D, h, w, H, W, edge_len, offsets_host, n_outer,
                                 n_inner, workspace, workspace_bytes, &sh, n_ref_total, ref_start, depth_init_out, depth_out, stream);
    *epoch = sh.epoch;  // also on failure: epochs already used must not be reused
    return rc;
}
