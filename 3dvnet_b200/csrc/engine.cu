// Engine: the whole hot path enqueued from native code by one call (dv3d_hot_path).
//
// The reference drives this path from Python, one library call per tensor op
// (eval-3dvnet.py:58-99 -> lightningmodel.py:124-242).  On a B200 the kernels of the path run
// 5-50 us each, so ~220 interpreter round trips per reference view would bound the step; here
// the same per-op entry points of this library are called back to back from C++ with a bump
// allocator over a caller-provided arena.  The sequence below mirrors, call for call, the
// Python-composed path in 3dvnet_b200/mv3d (which stays as the reference-shaped interface and
// as the parity cross-check: tests assert bit-identical depth).
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"

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
    const struct Level* out_lv;
    const struct Level* in_lv;
    int step;
};

struct Scene {
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

static int sparse_conv(const dv3d_dense_params_t& p, const float* feat, long long n_in, const KMap& km, long long n_out,
                       const float* residual, Scene& sc, float* out, void* st) {
    if (km.use_pairs && p.Wp)
        return dv3d_sparse_conv_pairs(feat, n_in, p.K / 27, km.plan, km.n_tiles, n_out, p.Wp, p.N, p.a, p.b, residual, 1,
                                      sc.pair_ws, sc.pair_ws_bytes, out, st);
    return dv3d_sparse_conv(feat, n_in, p.K / 27, km.nbr, n_out, p.W, p.Wp, p.N, p.a, p.b, residual, 1, sc.split_ws,
                            sc.split_ws_bytes, out, st);
}

// relu(x + GN2(conv2(relu(GN1(conv1(x))))))  (scenemodeling.py:16-44)
static int res_block(const dv3d_dense_params_t (&p)[2], const float* x, long long n, const KMap& nbr, Scene& sc, Arena& ar,
                     float** out, void* st) {
    const int C = p[0].N;
    float* h = ar.get<float>((size_t)n * C);
    float* y = ar.get<float>((size_t)n * C);
    ARENA_CHECK(ar);
    TRY(sparse_conv(p[0], x, n, nbr, n, nullptr, sc, h, st));
    TRY(sparse_conv(p[1], h, n, nbr, n, x, sc, y, st));
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
static int kernel_map(const Level& out_lv, const Level& in_lv, int step, bool want_plan, Arena& ar, KMap* km, void* st) {
    km->n_out = out_lv.n;
    km->nbr = ar.get<int>((size_t)out_lv.n * 27);
    km->plan = nullptr;
    km->n_tiles = 0;
    km->use_pairs = false;
    const size_t pb = want_plan ? dv3d_pair_plan_bytes(out_lv.n) : 0;
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
static int model_scene(const dv3d_net_params_t& net, const float* pts, const float* pts_feat, const long long* pts_batch,
                       long long N, double edge_len, Scene& sc, Arena& ar, void* st) {
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
    float* x0 = ar.get<float>((size_t)N * in_pad);
    float* xa = ar.get<float>((size_t)N * Hd);
    float* xb = ar.get<float>((size_t)N * Hd);
    // the four per-voxel max pools are produced by the epilogues of fc1..fc4 (dv3d_linear_pool): one 0xFF fill for all
    float* pools = ar.get<float>((size_t)4 * nv * Hd);
    float* F = ar.get<float>((size_t)nv * net.pointnet[5].N);
    ARENA_CHECK(ar);
    DV3D_CUDA(cudaMemsetAsync(pools, 0xFF, sizeof(float) * 4 * (size_t)nv * Hd, cs));
    TRY(dv3d_pointnet_input(pts, pts_feat, 32, a_pts, seg, N, 32, in_pad, x0, st));
    TRY(dv3d_linear(x0, in_pad, in_pad, nullptr, nullptr, 0, N, net.pointnet[0].W, net.pointnet[0].Wp, net.pointnet[0].b, Hd,
                    0, xa, st));
    TRY(dv3d_linear_pool(xa, Hd, Hd, nullptr, nullptr, 0, N, net.pointnet[1].W, net.pointnet[1].Wp, net.pointnet[1].b, Hd, 1,
                         xb, pools, seg, st));
    float *cur = xb, *nxt = xa;
    for (int i = 2; i <= 4; ++i) {
        float* pool_in = pools + (size_t)(i - 2) * nv * Hd;
        float* pool_out = pools + (size_t)(i - 1) * nv * Hd;
        TRY(dv3d_linear_pool(cur, Hd, Hd, pool_in, seg, Hd, N, net.pointnet[i].W, net.pointnet[i].Wp, net.pointnet[i].b, Hd, 1,
                             nxt, pool_out, seg, st));
        float* t = cur;
        cur = nxt;
        nxt = t;
    }
    TRY(dv3d_linear(pools + (size_t)3 * nv * Hd, Hd, Hd, nullptr, nullptr, 0, nv, net.pointnet[5].W, net.pointnet[5].Wp,
                    net.pointnet[5].b, net.pointnet[5].N, 1, F, st));

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
        KMap* maps[3 * DV3D_MAX_LEVELS];
        int n_maps = 0;
        for (int l = 0; l < nl; ++l) {
            TRY(kernel_map(sc.lv[l], sc.lv[l], sc.lv[l].stride, want_plan, ar, &sc.same[l], sst));
            maps[n_maps++] = &sc.same[l];
        }
        for (int l = 0; l + 1 < nl; ++l) {
            TRY(kernel_map(sc.lv[l + 1], sc.lv[l], sc.lv[l].stride, want_plan, ar, &sc.down[l], sst));
            maps[n_maps++] = &sc.down[l];
        }
        for (int l = 0; l + 1 < nl; ++l) {
            TRY(kernel_map(sc.lv[l], sc.lv[l + 1], -sc.lv[l].stride, want_plan, ar, &sc.up[l], sst));
            maps[n_maps++] = &sc.up[l];
        }
        {
            const int* co[3 * DV3D_MAX_LEVELS];
            long long no[3 * DV3D_MAX_LEVELS];
            const void* tb[3 * DV3D_MAX_LEVELS];
            size_t tbb[3 * DV3D_MAX_LEVELS];
            int stp[3 * DV3D_MAX_LEVELS];
            int* nb[3 * DV3D_MAX_LEVELS];
            for (int i = 0; i < n_maps; ++i) {
                co[i] = maps[i]->out_lv->coords;
                no[i] = maps[i]->out_lv->n;
                tb[i] = maps[i]->in_lv->table;
                tbb[i] = maps[i]->in_lv->table_bytes;
                stp[i] = maps[i]->step;
                nb[i] = maps[i]->nbr;
            }
            TRY(dv3d_kernel_map_batch(co, no, tb, tbb, stp, nb, n_maps, sst));
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
    for (int b = 0; b < net.n_res[0]; ++b) TRY(res_block(net.res_down[0][b], x, sc.lv[0].n, sc.same[0], sc, ar, &x, st));
    xs[0] = x;
    for (int i = 1; i < nl; ++i) {
        float* y = ar.get<float>((size_t)sc.lv[i].n * net.down[i - 1].N);
        ARENA_CHECK(ar);
        TRY(sparse_conv(net.down[i - 1], x, sc.lv[i - 1].n, sc.down[i - 1], sc.lv[i].n, nullptr, sc, y, st));
        x = y;
        for (int b = 0; b < net.n_res[i]; ++b) TRY(res_block(net.res_down[i][b], x, sc.lv[i].n, sc.same[i], sc, ar, &x, st));
        xs[i] = x;
    }
    sc.feats[nl - 1] = xs[nl - 1];
    sc.dims[nl - 1] = net.res_down[nl - 1][0][0].N;
    for (int i = 0; i < nl - 1; ++i) {
        const int l = nl - 2 - i;  // target (finer) level
        const int Cu = net.up[i].N;
        float* up = ar.get<float>((size_t)sc.lv[l].n * Cu);
        float* adj = ar.get<float>((size_t)sc.lv[l].n * net.feat_adj[i].N);
        ARENA_CHECK(ar);
        TRY(sparse_conv(net.up[i], x, sc.lv[l + 1].n, sc.up[l], sc.lv[l].n, nullptr, sc, up, st));
        TRY(dv3d_concat_linear_gn_relu(up, Cu, xs[l], net.feat_adj[i].K - Cu, sc.lv[l].n, net.feat_adj[i].W, net.feat_adj[i].Wp,
                                       net.feat_adj[i].N, net.feat_adj[i].a, net.feat_adj[i].b, adj, st));
        x = adj;
        // the reversed n_res list: res_up[i] has n_res[nl-2-i] blocks (scenemodeling.py:168-175)
        for (int b = 0; b < net.n_res[l]; ++b) TRY(res_block(net.res_up[i][b], x, sc.lv[l].n, sc.same[l], sc, ar, &x, st));
        sc.feats[l] = x;
        sc.dims[l] = net.feat_adj[i].N;
    }
    // position of index (0,0,0) of every batch (scenemodeling.py:211-226 / refinement.py:33)
    sc.origin = ar.get<float>((size_t)grid.n_batch * 3);
    ARENA_CHECK(ar);
    DV3D_CUDA(cudaMemsetAsync(sc.origin, 0, sizeof(float) * grid.n_batch * 3, cs));
    TRY(dv3d_batch_origin(a_pts, a_idx, a_batch, nv, (float)edge_len, sc.origin, st));
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
    const dv3d_net_params_t& net = *netp;
    DV3D_REQUIRE(net.n_levels >= 1 && net.n_levels <= DV3D_MAX_LEVELS, "hot_path: n_levels out of range");
    if (n_ref == 0) return DV3D_OK;
    cudaStream_t cs = (cudaStream_t)stream;
    Arena ar{(char*)workspace, workspace_bytes, 0, false};
    const long long P = (long long)h * w, Np = (long long)n_ref * P;

    // ================= path A: cost volume -> CostRegNet -> soft-argmin (mvsnet.py:176-229)
    float* cams = ar.get<float>((size_t)n_imgs * 36);
    float* x_var = ar.get<float>((size_t)n_ref * 32 * D * P);
    ARENA_CHECK(ar);
    TRY(dv3d_camera_tables(rotmats, tvecs, K, n_imgs, cams, stream));
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
    float* depth = ar.get<float>((size_t)Np);
    float* offs = ar.get<float>((size_t)Np);
    ARENA_CHECK(ar);
    const double depth_end = depth_start + depth_interval * (D - 1);
    {
        Prof pr(DV3D_STAGE_SOFTARGMIN, cs);
        TRY(dv3d_prob_softargmin(act[9], n_ref, net.costreg[9].Cout, D, h, w, net.prob_weight, net.prob_bias,
                                 (float)depth_start, (float)depth_end, nullptr, depth, stream));
    }
    if (depth_init_out) DV3D_CUDA(cudaMemcpyAsync(depth_init_out, depth, sizeof(float) * Np, cudaMemcpyDeviceToDevice, cs));

    // ================= path B: volumetric refinement (eval-3dvnet.py:73-99)
    if (n_outer * n_inner > 0) {
        ar.off = 0;  // path A's volumes are dead; keep only cams / depth / offs by re-reserving them first
        float* cams2 = ar.get<float>((size_t)n_imgs * 36);
        (void)cams2;  // same address as cams
        // depth / offs live behind the volumes: move them to the front of the arena
        float* depth_f = ar.get<float>((size_t)Np);
        float* offs_f = ar.get<float>((size_t)Np);
        DV3D_CUDA(cudaMemcpyAsync(depth_f, depth, sizeof(float) * Np, cudaMemcpyDeviceToDevice, cs));
        depth = depth_f;
        offs = offs_f;
        const int in_dim = net.dec[0].K / 3, var_off = in_dim - 32;
        float* operand = ar.get<float>((size_t)Np * 8 * in_dim);  // [n_pts, 8, in_dim], padding row stays zero
        float* dec_a = ar.get<float>((size_t)Np * 8 * net.dec[0].N);
        float* dec_b = ar.get<float>((size_t)Np * 8 * net.dec[1].N);
        float* pts_hyp = ar.get<float>((size_t)Np * 7 * 3);
        long long* pts_batch = ar.get<long long>((size_t)Np);
        const size_t split_ws_bytes = dv3d_sparse_conv_workspace_bytes(128);
        void* split_ws = ar.get<char>(split_ws_bytes);
        ARENA_CHECK(ar);
        DV3D_CUDA(cudaMemsetAsync(operand, 0, sizeof(float) * Np * 8 * in_dim, cs));
        DV3D_CUDA(cudaMemsetAsync(split_ws, 0, dv3d_sparse_conv_workspace_bytes(128), cs));
        DV3D_LAUNCH((expand_batch_kernel), cdiv(Np, 256), 256, 0, cs, depth_batch, (int)P, Np, pts_batch);
        DV3D_LAUNCHED();
        const size_t mark = ar.off;
        for (int o = 0; o < n_outer; ++o) {
            ar.off = mark;
            // feature-rich point cloud (lightningmodel.py:132-174)
            float* pts = ar.get<float>((size_t)Np * 3);
            float* pfeat = ar.get<float>((size_t)Np * 32);
            ARENA_CHECK(ar);
            {
                Prof pr(DV3D_STAGE_POINTCLOUD, cs);
                TRY(dv3d_points_var(feats_nhwc, n_imgs, 32, Hf, Wf, cams, ref_img, edge_rowptr, edge_src, depth, n_ref, h, w, H,
                                    W, 0, 0.0, pts, pfeat, 1, 32, 0, stream));
            }
            Scene sc;
            memset(&sc, 0, sizeof(sc));
            sc.split_ws = split_ws;
            sc.split_ws_bytes = dv3d_sparse_conv_workspace_bytes(128);
            TRY(model_scene(net, pts, pfeat, pts_batch, Np, edge_len, sc, ar, stream));
            for (int it = 0; it < n_inner; ++it) {
                const double offset = offsets_host[o * n_inner + it];
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
            }
        }
    }
    DV3D_CUDA(cudaMemcpyAsync(depth_out, depth, sizeof(float) * Np, cudaMemcpyDeviceToDevice, cs));
    return DV3D_OK;
}
