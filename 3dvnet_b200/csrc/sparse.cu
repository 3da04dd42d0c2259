// Sparse 3D-UNet plumbing: coordinate hash, kernel maps, the sparse convolution / 1x1
// "feature adjust" entry points (thin descriptors over the gather-GEMM) and the trilinear
// sparse interpolation of PointFlow.  MinkowskiEngine 0.5 semantics per SURVEY.md A.4:
// offsets enumerate x fastest (k = (dx+1) + 3(dy+1) + 9(dz+1)), no bias, strided maps are
// floor(c / s) * s, the transposed convolution reuses the finer map with the forward kernel
// map swapped, interpolation takes the 8 corners lower + {0,ts}^3 and missing voxels add 0.
#include <math.h>

#include <algorithm>

#include "gemm.cuh"

namespace dv3d {

__global__ void __launch_bounds__(256)
hash_clear_kernel(unsigned long long* keys, int* rows, size_t cap) {
    pdl_wait();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) {
        keys[i] = kEmptyKey;
        rows[i] = INT_MAX;
    }
}

__global__ void __launch_bounds__(256)
hash_insert_kernel(const int* __restrict__ coords, long long n, HashView t, int* __restrict__ err) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int b = coords[4 * i], x = coords[4 * i + 1], y = coords[4 * i + 2], z = coords[4 * i + 3];
    if (!coord_in_range(b, x, y, z)) {
        *err = 1;
        return;
    }
    unsigned long long key = coord_key(b, x, y, z);
    unsigned slot = hash_mix(key) & t.mask;
    while (true) {
        unsigned long long prev = atomicCAS(t.keys + slot, kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) {  // duplicate coordinates: lowest row wins deterministically
            atomicMin(t.rows + slot, (int)i);
            return;
        }
        slot = (slot + 1) & t.mask;
    }
}

// several tables in two launches: blockIdx.y selects the table
constexpr int HB_MAX = 4;
struct HashBatch {
    const int* coords[HB_MAX];
    long long n[HB_MAX];
    HashView t[HB_MAX];
};
__global__ void __launch_bounds__(256)
hash_clear_batch_kernel(const __grid_constant__ HashBatch b) {
    pdl_wait();
    const HashView& t = b.t[blockIdx.y];
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= (size_t)t.mask) {
        t.keys[i] = kEmptyKey;
        t.rows[i] = INT_MAX;
    }
}
__global__ void __launch_bounds__(256)
hash_insert_batch_kernel(const __grid_constant__ HashBatch b, int* __restrict__ err) {
    pdl_wait();
    const int j = blockIdx.y;
    const HashView t = b.t[j];
    const int* __restrict__ coords = b.coords[j];
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n[j]) return;
    int bb = coords[4 * i], x = coords[4 * i + 1], y = coords[4 * i + 2], z = coords[4 * i + 3];
    if (!coord_in_range(bb, x, y, z)) {
        *err = 1;
        return;
    }
    unsigned long long key = coord_key(bb, x, y, z);
    unsigned slot = hash_mix(key) & t.mask;
    while (true) {
        unsigned long long prev = atomicCAS(t.keys + slot, kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) {
            atomicMin(t.rows + slot, (int)i);
            return;
        }
        slot = (slot + 1) & t.mask;
    }
}

// nbr[o*27 + k] = row of (coords_out[o] + offset_k * step) in the input level, or -1
__global__ void __launch_bounds__(256)
kernel_map_kernel(const int* __restrict__ coords_out, long long n_out, HashView t, int step, int* __restrict__ nbr) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out * 27) return;
    long long o = i / 27;
    int k = (int)(i - o * 27);
    int dx = k % 3 - 1, dy = (k / 3) % 3 - 1, dz = k / 9 - 1;
    int b = coords_out[4 * o], x = coords_out[4 * o + 1] + dx * step, y = coords_out[4 * o + 2] + dy * step,
        z = coords_out[4 * o + 3] + dz * step;
    nbr[i] = coord_in_range(b, x, y, z) ? hash_find(t, coord_key(b, x, y, z)) : -1;
}

// all kernel maps of a scene in one launch: blockIdx.y selects the map
constexpr int KM_MAX_MAPS = 16;
struct KernelMapBatch {
    const int* coords_out[KM_MAX_MAPS];
    long long n_out[KM_MAX_MAPS];
    HashView table[KM_MAX_MAPS];
    int step[KM_MAX_MAPS];
    int* nbr[KM_MAX_MAPS];
};
__global__ void __launch_bounds__(256)
kernel_map_batch_kernel(const __grid_constant__ KernelMapBatch b) {
    pdl_wait();
    const int j = blockIdx.y;
    const long long n_out = b.n_out[j];
    const int* __restrict__ coords_out = b.coords_out[j];
    const int step = b.step[j];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_out * 27; i += (long long)gridDim.x * blockDim.x) {
        long long o = i / 27;
        int k = (int)(i - o * 27);
        int dx = k % 3 - 1, dy = (k / 3) % 3 - 1, dz = k / 9 - 1;
        int bb = coords_out[4 * o], x = coords_out[4 * o + 1] + dx * step, y = coords_out[4 * o + 2] + dy * step,
            z = coords_out[4 * o + 3] + dz * step;
        b.nbr[j][i] = coord_in_range(bb, x, y, z) ? hash_find(b.table[j], coord_key(bb, x, y, z)) : -1;
    }
}

// ---- work-balanced row ranges for a level sharded over GPUs (csrc/engine.cu, Shard) -----------------------------
// The cost of a sparse convolution over a row range follows its number of (row, neighbour) pairs, and rows are sorted
// by voxel id: equal row counts give the ranks that own the dense middle of a scene up to 1.5x the work of the edge
// ranks.  One warp counts the occupied neighbours of every `stride`-th row ...
__global__ void __launch_bounds__(256)
sample_neighbours_kernel(const int* __restrict__ coords, long long n_samples, int stride, HashView t, int step,
                         int* __restrict__ cnt) {
    pdl_wait();
    const long long s = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= n_samples) return;
    const int k = threadIdx.x & 31;
    const long long o = s * stride;
    bool hit = false;
    if (k < 27) {
        const int dx = k % 3 - 1, dy = (k / 3) % 3 - 1, dz = k / 9 - 1;
        const int b = coords[4 * o], x = coords[4 * o + 1] + dx * step, y = coords[4 * o + 2] + dy * step,
                  z = coords[4 * o + 3] + dz * step;
        hit = coord_in_range(b, x, y, z) && hash_find(t, coord_key(b, x, y, z)) >= 0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (k == 0) cnt[s] = __popc(m);
}
// ... and one block cuts the prefix sum of the samples into `world` equal parts: bounds[k] = first row of rank k
// (a multiple of 128, the GEMM row tile, except bounds[world] = n).  Integer arithmetic only: every rank derives the
// same bounds from its identical copy of the level.
__global__ void __launch_bounds__(1024)
balanced_bounds_kernel(const int* __restrict__ cnt, long long n_samples, int stride, long long n, int world,
                       int* __restrict__ bounds) {
    pdl_wait();
    __shared__ long long s_warp[32];
    __shared__ long long s_total, s_carry;
    __shared__ int s_b[17];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    long long sum = 0;
    for (long long i = tid; i < n_samples; i += 1024) sum += cnt[i];
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_warp[wid] = sum;
    if (tid <= world) s_b[tid] = tid == world ? (int)n : 0;
    __syncthreads();
    if (tid == 0) {
        long long t = 0;
        for (int w = 0; w < 32; ++w) t += s_warp[w];
        s_total = t;
        s_carry = 0;
    }
    __syncthreads();
    const long long total = s_total;
    for (long long base = 0; base < n_samples; base += 1024) {
        const long long i = base + tid;
        const long long v = i < n_samples ? cnt[i] : 0;
        long long incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const long long up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        __syncthreads();   // s_warp / s_carry of the previous chunk are consumed
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        long long off = s_carry;
        for (int w = 0; w < wid; ++w) off += s_warp[w];
        incl += off;
        const long long excl = incl - v;
        if (i < n_samples)
            for (int k = 1; k < world; ++k) {
                const long long target = total * k / world;
                if (excl < target && target <= incl) {   // the sample where the running sum crosses k / world of the work
                    long long row = (i + 1) * stride;
                    row = (row + 64) / 128 * 128;
                    s_b[k] = (int)(row > n ? n : row);
                }
            }
        __syncthreads();
        if (tid == 1023) s_carry = incl;
    }
    __syncthreads();
    if (tid == 0) {
        if (total == 0) {   // fewer rows than one sample: equal row counts
            const long long m = (n + world - 1) / world;
            for (int k = 1; k < world; ++k) s_b[k] = (int)(k * m < n ? k * m : n);
        }
        s_b[world] = (int)n;
        for (int k = 1; k <= world; ++k)   // monotone: a level too small to cut leaves some ranks without rows
            if (s_b[k] < s_b[k - 1]) s_b[k] = s_b[k - 1];
        for (int k = 0; k <= world; ++k) bounds[k] = s_b[k];
    }
}

// One warp per query point.  q = ((p - origin[b]) / res) * stride in base-voxel units
// (refinement.py:34-35); lanes 0..7 probe the 8 corners, all lanes accumulate C channels.
template <int C>
__device__ __forceinline__ void sparse_interp_body(const float* __restrict__ pts, const long long* __restrict__ pts_batch,
                                                   long long Nq, int n_hyp, int rows_per_point,
                                                   const float* __restrict__ origin, float res, int stride, HashView t,
                                                   const float* __restrict__ feat, float* __restrict__ out, int out_ld,
                                                   int out_off) {
    const int lane = threadIdx.x & 31;
    const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= Nq) return;
    const long long p = q / n_hyp;
    const int hy = (int)(q - p * n_hyp);
    const int b = (int)__ldg(pts_batch + p);
    const float ts = (float)stride;
    float qc[3];
    int lo[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float v = __fsub_rn(__ldg(pts + 3 * q + d), __ldg(origin + 3 * b + d));
        v = __fmul_rn(__fdiv_rn(v, res), ts);
        qc[d] = v;
        lo[d] = (int)(floorf(__fdiv_rn(v, ts)) * ts);
    }
    int row = -1;
    float wgt = 0.f;
    if (lane < 8) {
        int cx = lo[0] + ((lane >> 0) & 1) * stride, cy = lo[1] + ((lane >> 1) & 1) * stride,
            cz = lo[2] + ((lane >> 2) & 1) * stride;
        wgt = (1.f - fabsf(qc[0] - (float)cx) / ts) * (1.f - fabsf(qc[1] - (float)cy) / ts) *
              (1.f - fabsf(qc[2] - (float)cz) / ts);
        if (coord_in_range(b, cx, cy, cz)) row = hash_find(t, coord_key(b, cx, cy, cz));
    }
    constexpr int V = C / 32;  // channels per lane (2 or 4)
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        int r = __shfl_sync(0xffffffffu, row, c);
        float w = __shfl_sync(0xffffffffu, wgt, c);
        if (r >= 0) {
            const float* f = feat + (size_t)r * C + lane * V;
            if (V == 4) {
                float4 v = __ldg(reinterpret_cast<const float4*>(f));
                acc[0] = fmaf(w, v.x, acc[0]); acc[1] = fmaf(w, v.y, acc[1]);
                acc[2] = fmaf(w, v.z, acc[2]); acc[3 % V] = fmaf(w, v.w, acc[3 % V]);
            } else {
                float2 v = __ldg(reinterpret_cast<const float2*>(f));
                acc[0] = fmaf(w, v.x, acc[0]); acc[1] = fmaf(w, v.y, acc[1]);
            }
        }
    }
    float* o = out + ((size_t)p * rows_per_point + hy) * out_ld + out_off + lane * V;
    if (V == 4)
        *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2 % V], acc[3 % V]);
    else
        *reinterpret_cast<float2*>(o) = make_float2(acc[0], acc[1]);
}

template <int C>
__global__ void __launch_bounds__(256)
sparse_interp_kernel(const float* __restrict__ pts, const long long* __restrict__ pts_batch, long long Nq, int n_hyp,
                     int rows_per_point, const float* __restrict__ origin, float res, int stride, HashView t,
                     const float* __restrict__ feat, float* __restrict__ out, int out_ld, int out_off) {
    pdl_wait();
    sparse_interp_body<C>(pts, pts_batch, Nq, n_hyp, rows_per_point, origin, res, stride, t, feat, out, out_ld, out_off);
}

// every level of the U-Net in one launch: blockIdx.y selects the level
constexpr int SI_MAX_LEVELS = 4;
struct InterpBatch {
    float res[SI_MAX_LEVELS];
    int stride[SI_MAX_LEVELS];
    HashView table[SI_MAX_LEVELS];
    const float* feat[SI_MAX_LEVELS];
    int C[SI_MAX_LEVELS];
    int out_off[SI_MAX_LEVELS];
};
__global__ void __launch_bounds__(256)
sparse_interp_batch_kernel(const float* __restrict__ pts, const long long* __restrict__ pts_batch, long long Nq, int n_hyp,
                           int rows_per_point, const float* __restrict__ origin, const __grid_constant__ InterpBatch b,
                           float* __restrict__ out, int out_ld) {
    pdl_wait();
    const int l = blockIdx.y;
    if (b.C[l] == 64)
        sparse_interp_body<64>(pts, pts_batch, Nq, n_hyp, rows_per_point, origin, b.res[l], b.stride[l], b.table[l], b.feat[l],
                               out, out_ld, b.out_off[l]);
    else
        sparse_interp_body<128>(pts, pts_batch, Nq, n_hyp, rows_per_point, origin, b.res[l], b.stride[l], b.table[l],
                                b.feat[l], out, out_ld, b.out_off[l]);
}

}  // namespace dv3d

using namespace dv3d;

extern "C" size_t dv3d_hash_bytes(long long n_rows) { return n_rows < 0 ? 0 : hash_capacity_for(n_rows) * 12; }

extern "C" int dv3d_hash_build(const int* coords, long long n, void* table, size_t table_bytes, int* err_flag,
                               void* stream) {
    HashView t;
    DV3D_REQUIRE(coords && table && err_flag && n >= 0, "hash_build: bad arguments");
    DV3D_REQUIRE(hash_view(table, table_bytes, &t) && (size_t)t.mask + 1 >= (size_t)(2 * n),
                 "hash_build: table_bytes must be dv3d_hash_bytes(n)");
    cudaStream_t st = (cudaStream_t)stream;
    size_t cap = (size_t)t.mask + 1;
    DV3D_LAUNCH((hash_clear_kernel), cdiv(cap, 256), 256, 0, st, t.keys, t.rows, cap);
    DV3D_LAUNCHED();
    if (n == 0) return DV3D_OK;
    DV3D_LAUNCH((hash_insert_kernel), cdiv(n, 256), 256, 0, st, coords, n, t, err_flag);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_hash_build_batch(const int* const* coords, const long long* n, void* const* tables,
                                     const size_t* table_bytes, int n_tables, int* err_flag, void* stream) {
    DV3D_REQUIRE(coords && n && tables && table_bytes && err_flag && n_tables >= 1 && n_tables <= HB_MAX,
                 "hash_build_batch: bad arguments (1..%d tables)", HB_MAX);
    HashBatch b = {};
    size_t max_cap = 0;
    long long max_n = 0;
    for (int j = 0; j < n_tables; ++j) {
        DV3D_REQUIRE(coords[j] && tables[j] && n[j] >= 0, "hash_build_batch: bad table %d", j);
        DV3D_REQUIRE(hash_view(tables[j], table_bytes[j], &b.t[j]) && (size_t)b.t[j].mask + 1 >= (size_t)(2 * n[j]),
                     "hash_build: table_bytes must be dv3d_hash_bytes(n)");
        b.coords[j] = coords[j];
        b.n[j] = n[j];
        max_cap = std::max(max_cap, (size_t)b.t[j].mask + 1);
        max_n = std::max(max_n, n[j]);
    }
    cudaStream_t st = (cudaStream_t)stream;
    DV3D_LAUNCH((hash_clear_batch_kernel), dim3(cdiv(max_cap, 256), n_tables), 256, 0, st, b);
    DV3D_LAUNCHED();
    if (max_n == 0) return DV3D_OK;
    DV3D_LAUNCH((hash_insert_batch_kernel), dim3(cdiv(max_n, 256), n_tables), 256, 0, st, b, err_flag);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_kernel_map(const int* coords_out, long long n_out, const void* table_in, size_t table_bytes,
                               int step, int* nbr, void* stream) {
    HashView t;
    DV3D_REQUIRE(coords_out && table_in && nbr && n_out >= 0, "kernel_map: bad arguments");
    DV3D_REQUIRE(hash_view(const_cast<void*>(table_in), table_bytes, &t), "kernel_map: bad table size");
    if (n_out == 0) return DV3D_OK;
    DV3D_LAUNCH((kernel_map_kernel), cdiv(n_out * 27, 256), 256, 0, (cudaStream_t)stream, coords_out, n_out, t, step, nbr);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

// K-split workspace: a fixed counter block (splitting only happens below 148 row tiles), then
// room for one raw partial per resident CTA
constexpr size_t kSplitCounterBytes = 4096;

extern "C" int dv3d_kernel_map_batch(const int* const* coords_out, const long long* n_out, const void* const* table_in,
                                     const size_t* table_bytes, const int* step, int* const* nbr, int n_maps, void* stream) {
    DV3D_REQUIRE(coords_out && n_out && table_in && table_bytes && step && nbr && n_maps >= 0 && n_maps <= KM_MAX_MAPS,
                 "kernel_map_batch: bad arguments (at most %d maps per call)", KM_MAX_MAPS);
    KernelMapBatch b = {};
    long long max_n = 0;
    for (int i = 0; i < n_maps; ++i) {
        DV3D_REQUIRE(coords_out[i] && table_in[i] && nbr[i] && n_out[i] >= 0, "kernel_map_batch: bad map %d", i);
        DV3D_REQUIRE(hash_view(const_cast<void*>(table_in[i]), table_bytes[i], &b.table[i]), "kernel_map_batch: bad table size");
        b.coords_out[i] = coords_out[i];
        b.n_out[i] = n_out[i];
        b.step[i] = step[i];
        b.nbr[i] = nbr[i];
        if (n_out[i] > max_n) max_n = n_out[i];
    }
    if (n_maps == 0 || max_n == 0) return DV3D_OK;
    int gx = cdiv(max_n * 27, 256);
    if (gx > 8 * kNumSMs) gx = 8 * kNumSMs;
    DV3D_LAUNCH((kernel_map_batch_kernel), dim3(gx, n_maps), 256, 0, (cudaStream_t)stream, b);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_sparse_interp_batch(const float* pts, const long long* pts_batch, long long n_pts, int n_hyp,
                                        int rows_per_point, const float* origin, int n_levels, const float* res,
                                        const int* stride, const void* const* table, const size_t* table_bytes,
                                        const float* const* feat, const int* C, const int* out_off, float* out, int out_ld,
                                        void* stream) {
    DV3D_REQUIRE(pts && pts_batch && origin && res && stride && table && table_bytes && feat && C && out_off && out &&
                     n_pts >= 0 && n_hyp > 0 && rows_per_point >= n_hyp && n_levels >= 1 && n_levels <= SI_MAX_LEVELS,
                 "sparse_interp_batch: bad arguments (1..%d levels)", SI_MAX_LEVELS);
    InterpBatch b = {};
    for (int l = 0; l < n_levels; ++l) {
        DV3D_REQUIRE(table[l] && feat[l] && res[l] > 0.f && stride[l] > 0, "sparse_interp_batch: bad level %d", l);
        DV3D_REQUIRE(hash_view(const_cast<void*>(table[l]), table_bytes[l], &b.table[l]), "sparse_interp_batch: bad table size");
        DV3D_REQUIRE(C[l] == 64 || C[l] == 128, "sparse_interp_batch: C must be 64 or 128, got %d", C[l]);
        DV3D_REQUIRE(out_ld % 4 == 0 && out_off[l] % 4 == 0 && out_off[l] + C[l] <= out_ld, "sparse_interp_batch: bad output window");
        b.res[l] = res[l];
        b.stride[l] = stride[l];
        b.feat[l] = feat[l];
        b.C[l] = C[l];
        b.out_off[l] = out_off[l];
    }
    const long long Nq = n_pts * n_hyp;
    if (Nq == 0) return DV3D_OK;
    DV3D_LAUNCH((sparse_interp_batch_kernel), dim3(cdiv(Nq, 8), n_levels), 256, 0, (cudaStream_t)stream, pts, pts_batch, Nq, n_hyp,
                rows_per_point, origin, b, out, out_ld);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" size_t dv3d_sparse_conv_workspace_bytes(int Cout) {
    if (Cout <= 0) return 0;
    return kSplitCounterBytes + (size_t)kNumSMs * 128 * Cout * sizeof(float);
}

extern "C" int dv3d_sparse_conv(const float* feat, long long n_in, int Cin, const int* nbr, long long n_out,
                                const float* W, const void* W_packed, int Cout, const float* gn_weight,
                                const float* gn_bias, const float* residual, int relu, void* workspace,
                                size_t workspace_bytes, float* out, void* stream) {
    DV3D_REQUIRE(feat && nbr && (W || W_packed) && out && n_in >= 0 && n_out >= 0, "sparse_conv: bad arguments");
    GemmDesc d = {};
    d.n_slices = 27;
    for (int k = 0; k < 27; ++k) d.slice[k] = GemmSlice{feat, nbr + k, 27, 0, Cin, Cin};
    d.kmap = nbr;
    d.Wp = (const float*)W_packed;
    if (workspace && W_packed && workspace_bytes > kSplitCounterBytes) {
        DV3D_REQUIRE(((uintptr_t)workspace & 255) == 0, "sparse_conv: workspace must be 256-byte aligned");
        d.split_counters = (int*)workspace;
        d.split_ws = (float*)((char*)workspace + kSplitCounterBytes);
        d.split_ws_bytes = workspace_bytes - kSplitCounterBytes;
    }
    d.M = n_out;
    d.n_src_rows = n_in;
    d.N = Cout;
    d.W = W;
    d.gn_weight = gn_weight;
    d.gn_bias = gn_bias;
    d.residual = residual;
    d.res_ld = Cout;
    d.relu_out = relu;
    d.out = out;
    d.out_ld = Cout;
    return launch_gather_gemm(d, (cudaStream_t)stream);
}

extern "C" int dv3d_concat_linear_gn_relu(const float* a, int Ca, const float* b, int Cb, long long n, const float* W,
                                          const void* W_packed, int Cout, const float* gn_weight,
                                          const float* gn_bias, float* out, void* stream) {
    DV3D_REQUIRE(a && b && (W || W_packed) && out && n >= 0, "concat_linear: bad arguments");
    GemmDesc d = {};
    d.Wp = (const float*)W_packed;
    d.n_slices = 2;
    d.slice[0] = GemmSlice{a, nullptr, 0, 0, Ca, Ca};
    d.slice[1] = GemmSlice{b, nullptr, 0, 0, Cb, Cb};
    d.M = n;
    d.n_src_rows = n;
    d.N = Cout;
    d.W = W;
    d.gn_weight = gn_weight;
    d.gn_bias = gn_bias;
    d.relu_out = 1;
    d.out = out;
    d.out_ld = Cout;
    return launch_gather_gemm(d, (cudaStream_t)stream);
}

extern "C" int dv3d_sparse_interp(const float* pts, const long long* pts_batch, long long n_pts, int n_hyp,
                                  int rows_per_point, const float* origin, float res, int stride, const void* table,
                                  size_t table_bytes, const float* feat, int C, float* out, int out_ld, int out_off,
                                  void* stream) {
    HashView t;
    DV3D_REQUIRE(pts && pts_batch && origin && table && feat && out && n_pts >= 0 && n_hyp > 0 &&
                     rows_per_point >= n_hyp && res > 0.f && stride > 0,
                 "sparse_interp: bad arguments");
    DV3D_REQUIRE(hash_view(const_cast<void*>(table), table_bytes, &t), "sparse_interp: bad table size");
    DV3D_REQUIRE(C == 64 || C == 128, "sparse_interp: C must be 64 or 128, got %d", C);
    DV3D_REQUIRE(out_ld % 4 == 0 && out_off % 4 == 0 && out_off + C <= out_ld, "sparse_interp: bad output window");
    const long long Nq = n_pts * n_hyp;
    if (Nq == 0) return DV3D_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 64)
        DV3D_LAUNCH((sparse_interp_kernel<64>), cdiv(Nq, 8), 256, 0, st, pts, pts_batch, Nq, n_hyp, rows_per_point, origin, res, stride, t, feat, out, out_ld, out_off);
    else
        DV3D_LAUNCH((sparse_interp_kernel<128>), cdiv(Nq, 8), 256, 0, st, pts, pts_batch, Nq, n_hyp, rows_per_point, origin, res, stride, t, feat, out, out_ld, out_off);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" size_t dv3d_balanced_row_bounds_scratch_bytes(long long n, int sample_stride) {
    if (n < 0 || sample_stride <= 0) return 0;
    return align_up((size_t)((n + sample_stride - 1) / sample_stride + 1) * sizeof(int), 256);
}

extern "C" int dv3d_balanced_row_bounds(const int* coords, long long n, const void* table, size_t table_bytes, int step,
                                        int world, int sample_stride, void* scratch, int* bounds_dev, void* stream) {
    HashView t;
    DV3D_REQUIRE(coords && table && scratch && bounds_dev && n >= 0 && world >= 1 && world <= 16 && sample_stride >= 1,
                 "balanced_row_bounds: bad arguments");
    DV3D_REQUIRE(hash_view(const_cast<void*>(table), table_bytes, &t), "balanced_row_bounds: bad table size");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_samples = n / sample_stride;   // rows 0, s, 2 s, ... < n
    int* cnt = reinterpret_cast<int*>(scratch);
    if (n_samples > 0) {
        DV3D_LAUNCH((sample_neighbours_kernel), cdiv(n_samples, 8), 256, 0, st, coords, n_samples, sample_stride, t, step, cnt);
        DV3D_LAUNCHED();
    }
    DV3D_LAUNCH((balanced_bounds_kernel), 1, 1024, 0, st, (const int*)cnt, n_samples, sample_stride, n, world, bounds_dev);
    DV3D_LAUNCHED();
    return DV3D_OK;
}
