// Voxelisation of the back-projected point cloud (utils.py:38-64 + torch_cluster grid_cluster)
// and the coarser coordinate levels of the sparse U-Net (MinkowskiEngine stride-2 maps).
//
// The reference obtains the sorted unique voxel ids with torch.unique (a device-wide sort).
// Here the occupied cells of the bounding-box grid are marked in a bitmap; an exclusive scan
// of the per-word popcounts gives every occupied cell its rank in ascending id order, i.e.
// exactly torch.unique's ordering and inverse map, with no sort and no hash.  All index
// arithmetic follows the reference's fp32 / int64 operation order (no FMA contraction, IEEE
// division): results are bit-exact.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace dv3d {

constexpr int SCAN_T = 256;               // threads per scan block
constexpr int SCAN_W = 4;                 // words per thread
constexpr int SCAN_BLK = SCAN_T * SCAN_W;  // words per scan block
constexpr int kUntouched = 0x7f7f7f7f;     // memset(0x7f) marker of a per-batch minimum nobody wrote

struct BBoxHeader {  // device-side reduction target
    unsigned mn[3], mx[3];  // order-encoded floats
    unsigned long long bmax;
};

__global__ void bbox_init_kernel(BBoxHeader* h) {
    pdl_wait();
    if (threadIdx.x < 3) {
        h->mn[threadIdx.x] = 0xffffffffu;
        h->mx[threadIdx.x] = 0u;
    }
    if (threadIdx.x == 0) h->bmax = 0ull;
}

__global__ void __launch_bounds__(256)
bbox_kernel(const float* __restrict__ pts, const long long* __restrict__ batch, long long N, BBoxHeader* h) {
    pdl_wait();
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    long long bm = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            float v = __ldg(pts + 3 * i + d);
            mn[d] = fminf(mn[d], v);
            mx[d] = fmaxf(mx[d], v);
        }
        long long b = __ldg(batch + i);
        bm = b > bm ? b : bm;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
        long long ob = __shfl_xor_sync(0xffffffffu, bm, o);
        bm = ob > bm ? ob : bm;
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            atomicMin(&h->mn[d], f2ord(mn[d]));
            atomicMax(&h->mx[d], f2ord(mx[d]));
        }
        atomicMax(&h->bmax, (unsigned long long)bm);
    }
}

__global__ void bbox_decode_kernel(const BBoxHeader* h, float* out8) {
    pdl_wait();
    if (threadIdx.x < 3) {
        out8[threadIdx.x] = ord2f(h->mn[threadIdx.x]);
        out8[3 + threadIdx.x] = ord2f(h->mx[threadIdx.x]);
    }
    if (threadIdx.x == 0) {
        out8[6] = (float)h->bmax;
        out8[7] = 0.f;
    }
}

struct GridDev {
    float bmin[3];
    float e;
    long long n[3];     // grid_cluster cells per dim: trunc((max-min)/e) + 1
    long long g[3];     // voxelize's grid_size: ceil((max-min)/e)
};

// 1-D voxel id of a point (torch_cluster grid kernel): fp32 subtract, fp32 divide, truncate
__device__ __forceinline__ long long voxel_id(const float* p, long long b, const GridDev& G) {
    long long cx = (long long)__fdiv_rn(__fsub_rn(p[0], G.bmin[0]), G.e);
    long long cy = (long long)__fdiv_rn(__fsub_rn(p[1], G.bmin[1]), G.e);
    long long cz = (long long)__fdiv_rn(__fsub_rn(p[2], G.bmin[2]), G.e);
    long long cb = (long long)__fdiv_rn(__fsub_rn((float)b, 0.f), 1.f);
    return cx + cy * G.n[0] + cz * (G.n[0] * G.n[1]) + cb * (G.n[0] * G.n[1] * G.n[2]);
}

__global__ void __launch_bounds__(256)
mark_points_kernel(const float* __restrict__ pts, const long long* __restrict__ batch, long long N, GridDev G,
                   long long total_cells, long long* __restrict__ point_id, unsigned* __restrict__ bitmap,
                   int* __restrict__ err) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float p[3] = {__ldg(pts + 3 * i), __ldg(pts + 3 * i + 1), __ldg(pts + 3 * i + 2)};
    long long id = voxel_id(p, __ldg(batch + i), G);
    if (id < 0 || id >= total_cells) {  // cannot happen for finite inputs
        *err = 1;
        id = 0;
    }
    point_id[i] = id;
    atomicOr(bitmap + (id >> 5), 1u << (id & 31));
}

// exclusive scan of popcounts, stage 1: per block of SCAN_BLK words
__device__ __forceinline__ void scan_words_body(const unsigned* __restrict__ bitmap, long long n_words,
                                                unsigned* __restrict__ prefix, unsigned* __restrict__ block_sums) {
    __shared__ unsigned s_warp[SCAN_T / 32];
    const long long base = (long long)blockIdx.x * SCAN_BLK + (long long)threadIdx.x * SCAN_W;
    unsigned c[SCAN_W], tot = 0;
#pragma unroll
    for (int j = 0; j < SCAN_W; ++j) {
        c[j] = (base + j < n_words) ? __popc(bitmap[base + j]) : 0u;
        tot += c[j];
    }
    unsigned incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += s_warp[w];
    unsigned run = woff + incl - tot;
#pragma unroll
    for (int j = 0; j < SCAN_W; ++j) {
        if (base + j < n_words) prefix[base + j] = run;
        run += c[j];
    }
    if (threadIdx.x == SCAN_T - 1) block_sums[blockIdx.x] = woff + incl;
}
__global__ void __launch_bounds__(SCAN_T)
scan_words_kernel(const unsigned* __restrict__ bitmap, long long n_words, unsigned* __restrict__ prefix,
                  unsigned* __restrict__ block_sums) {
    pdl_wait();
    scan_words_body(bitmap, n_words, prefix, block_sums);
}

// stage 2: one block turns block_sums into exclusive offsets and writes the grand total
__device__ __forceinline__ void scan_blocks_body(unsigned* __restrict__ block_sums, int n_blocks, long long* __restrict__ total) {
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += 1024) {
        int i = base + threadIdx.x;
        unsigned v = i < n_blocks ? block_sums[i] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        unsigned woff = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += s_warp[w];
        unsigned carry = s_carry;
        if (i < n_blocks) block_sums[i] = carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = (long long)s_carry;
}
__global__ void __launch_bounds__(1024)
scan_blocks_kernel(unsigned* __restrict__ block_sums, int n_blocks, long long* __restrict__ total) {
    pdl_wait();
    scan_blocks_body(block_sums, n_blocks, total);
}

__device__ __forceinline__ unsigned rank_of(long long id, const unsigned* bitmap, const unsigned* prefix,
                                            const unsigned* block_offs) {
    long long w = id >> 5;
    unsigned bits = bitmap[w] & ((1u << (id & 31)) - 1u);
    return block_offs[w / SCAN_BLK] + prefix[w] + __popc(bits);
}

// one thread per bitmap word: emit the anchors of its set bits in ascending id order.
// min_idx (utils.py:61, the per-batch minimum index) is reduced per warp and per block before it reaches global memory:
// one atomicMin per anchor and axis on the same three addresses cost 380 us for the 194 k anchors of a 64-view scene.
__global__ void __launch_bounds__(256)
emit_anchors_kernel(const unsigned* __restrict__ bitmap, const unsigned* __restrict__ prefix,
                    const unsigned* __restrict__ block_offs, long long n_words, GridDev G, float half_e,
                    long long cap, float* __restrict__ anchor_pts, int* __restrict__ anchor_idx3d,
                    long long* __restrict__ anchor_batch, int* __restrict__ min_idx) {
    pdl_wait();
    __shared__ long long s_b[8];
    __shared__ int s_min[8][3];
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned bits = w < n_words ? bitmap[w] : 0u;
    long long b0 = -1;                      // batch of this word's first anchor
    int mx = 0x7fffffff, my = 0x7fffffff, mz = 0x7fffffff;
    if (bits) {
        long long rank = (long long)block_offs[w / SCAN_BLK] + prefix[w];
        const long long cells = G.n[0] * G.n[1] * G.n[2];
        const long long max_grid_idx = G.g[0] * G.g[1] * G.g[2];
        const long long gxy = G.g[0] * G.g[1];
        while (bits) {
            int bit = __ffs(bits) - 1;
            bits &= bits - 1;
            if (rank < cap) {
                long long id = (w << 5) + bit;
                long long b = id / cells;                 // == scatter-min of the batch ids of its points
                long long a = id - b * max_grid_idx;      // utils.py:53
                int z = (int)(a / gxy);                   // utils.py:55
                long long rem = a - (long long)z * gxy;
                int y = (int)(rem / G.g[0]);              // utils.py:56
                int x = (int)(rem % G.g[0]);              // utils.py:57
                // utils.py:58: (idx * e + bbox_min) + e/2, separate roundings
                anchor_pts[3 * rank + 0] = __fadd_rn(__fadd_rn(__fmul_rn((float)x, G.e), G.bmin[0]), half_e);
                anchor_pts[3 * rank + 1] = __fadd_rn(__fadd_rn(__fmul_rn((float)y, G.e), G.bmin[1]), half_e);
                anchor_pts[3 * rank + 2] = __fadd_rn(__fadd_rn(__fmul_rn((float)z, G.e), G.bmin[2]), half_e);
                anchor_idx3d[3 * rank + 0] = x;
                anchor_idx3d[3 * rank + 1] = y;
                anchor_idx3d[3 * rank + 2] = z;
                anchor_batch[rank] = b;
                if (b0 < 0) b0 = b;
                if (b == b0) {
                    mx = min(mx, x), my = min(my, y), mz = min(mz, z);
                } else {                                  // a word that straddles two batches (one per batch boundary)
                    atomicMin(min_idx + 3 * b + 0, x);
                    atomicMin(min_idx + 3 * b + 1, y);
                    atomicMin(min_idx + 3 * b + 2, z);
                }
            }
            ++rank;
        }
    }
    // warp: lanes that share the batch of the warp's first occupied word reduce together, the others go direct
    const unsigned occ = __ballot_sync(0xffffffffu, b0 >= 0);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    long long wb = -1;
    if (occ) {
        wb = __shfl_sync(0xffffffffu, b0, __ffs(occ) - 1);
        const bool join = b0 == wb;
        if (b0 >= 0 && !join) {
            atomicMin(min_idx + 3 * b0 + 0, mx);
            atomicMin(min_idx + 3 * b0 + 1, my);
            atomicMin(min_idx + 3 * b0 + 2, mz);
        }
        mx = __reduce_min_sync(0xffffffffu, join ? mx : 0x7fffffff);
        my = __reduce_min_sync(0xffffffffu, join ? my : 0x7fffffff);
        mz = __reduce_min_sync(0xffffffffu, join ? mz : 0x7fffffff);
    }
    if (lane == 0) {
        s_b[wid] = wb;
        s_min[wid][0] = mx, s_min[wid][1] = my, s_min[wid][2] = mz;
    }
    __syncthreads();
    if (threadIdx.x == 0) {   // block: merge runs of equal batch ids (ids ascend with the word index)
        long long cur = -1;
        int cx = 0x7fffffff, cy = 0x7fffffff, cz = 0x7fffffff;
        for (int i = 0; i <= 8; ++i) {
            const long long bi = i < 8 ? s_b[i] : -2;
            if (i < 8 && bi < 0) continue;    // a warp without anchors
            if (bi != cur) {
                if (cur >= 0) {
                    atomicMin(min_idx + 3 * cur + 0, cx);
                    atomicMin(min_idx + 3 * cur + 1, cy);
                    atomicMin(min_idx + 3 * cur + 2, cz);
                }
                cur = bi;
                cx = cy = cz = 0x7fffffff;
            }
            if (i < 8) cx = min(cx, s_min[i][0]), cy = min(cy, s_min[i][1]), cz = min(cz, s_min[i][2]);
        }
    }
}

// utils.py:61-62 and the inverse map of torch.unique
__global__ void __launch_bounds__(256)
finish_voxelize_kernel(long long n_anchors, long long N, const long long* __restrict__ anchor_batch,
                       const int* __restrict__ min_idx, int* __restrict__ anchor_idx3d,
                       const long long* __restrict__ point_id, const unsigned* __restrict__ bitmap,
                       const unsigned* __restrict__ prefix, const unsigned* __restrict__ block_offs,
                       int* __restrict__ point_anchor) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_anchors) {
        long long b = anchor_batch[i];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            int m = min_idx[3 * b + d];
            anchor_idx3d[3 * i + d] -= (m == kUntouched ? 0 : m);  // empty scatter-min slots are 0
        }
    }
    if (i < N) point_anchor[i] = (int)rank_of(point_id[i], bitmap, prefix, block_offs);
}

// ------------------------------------------------------------------ coarser sparse levels
struct LevelDims {
    int X, Y, Z;   // cells per axis of the coarse lattice
    int stride;    // new tensor stride
};

__device__ __forceinline__ void mark_coarse_body(const int* __restrict__ coords, long long n, LevelDims L, int n_batch,
                                                 unsigned* __restrict__ bitmap, int* __restrict__ err) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int b = coords[4 * i], x = coords[4 * i + 1], y = coords[4 * i + 2], z = coords[4 * i + 3];
    // floor(c / stride) for the non-negative lattice the voxeliser produces
    int cx = x / L.stride, cy = y / L.stride, cz = z / L.stride;
    if (b < 0 || b >= n_batch || x < 0 || y < 0 || z < 0 || cx >= L.X || cy >= L.Y || cz >= L.Z) {
        *err = 1;
        return;
    }
    long long cell = (((long long)b * L.Z + cz) * L.Y + cy) * L.X + cx;  // (b, z, y, x) order
    atomicOr(bitmap + (cell >> 5), 1u << (cell & 31));
}
__global__ void __launch_bounds__(256)
mark_coarse_kernel(const int* __restrict__ coords, long long n, LevelDims L, int n_batch,
                   unsigned* __restrict__ bitmap, int* __restrict__ err) {
    pdl_wait();
    mark_coarse_body(coords, n, L, n_batch, bitmap, err);
}

__device__ __forceinline__ void emit_coarse_body(const unsigned* __restrict__ bitmap, const unsigned* __restrict__ prefix,
                                                 const unsigned* __restrict__ block_offs, long long n_words, LevelDims L,
                                                 long long cap, int* __restrict__ coarse) {
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    unsigned bits = bitmap[w];
    if (!bits) return;
    long long rank = (long long)block_offs[w / SCAN_BLK] + prefix[w];
    while (bits) {
        int bit = __ffs(bits) - 1;
        bits &= bits - 1;
        if (rank < cap) {
            long long cell = (w << 5) + bit;
            int x = (int)(cell % L.X);
            long long r = cell / L.X;
            int y = (int)(r % L.Y);
            r /= L.Y;
            int z = (int)(r % L.Z);
            int b = (int)(r / L.Z);
            coarse[4 * rank + 0] = b;
            coarse[4 * rank + 1] = x * L.stride;
            coarse[4 * rank + 2] = y * L.stride;
            coarse[4 * rank + 3] = z * L.stride;
        }
        ++rank;
    }
}
__global__ void __launch_bounds__(256)
emit_coarse_kernel(const unsigned* __restrict__ bitmap, const unsigned* __restrict__ prefix,
                   const unsigned* __restrict__ block_offs, long long n_words, LevelDims L, long long cap,
                   int* __restrict__ coarse) {
    pdl_wait();
    emit_coarse_body(bitmap, prefix, block_offs, n_words, L, cap, coarse);
}

// every coarser level of a scene in one pass over the finest coordinates: blockIdx.y selects the level
constexpr int CB_MAX = 4;
struct CoarsenBatch {
    LevelDims L[CB_MAX];
    long long n_words[CB_MAX];
    unsigned* bitmap[CB_MAX];
    unsigned* prefix[CB_MAX];
    unsigned* block_sums[CB_MAX];
    long long* total[CB_MAX];
    int* err[CB_MAX];
    int* coarse[CB_MAX];
};
__global__ void __launch_bounds__(256)
mark_coarse_batch_kernel(const int* __restrict__ coords, long long n, int n_batch, const __grid_constant__ CoarsenBatch b) {
    pdl_wait();
    const int l = blockIdx.y;
    mark_coarse_body(coords, n, b.L[l], n_batch, b.bitmap[l], b.err[l]);
}
__global__ void __launch_bounds__(SCAN_T)
scan_words_batch_kernel(const __grid_constant__ CoarsenBatch b) {
    pdl_wait();
    const int l = blockIdx.y;
    if ((long long)blockIdx.x * SCAN_BLK >= b.n_words[l]) return;
    scan_words_body(b.bitmap[l], b.n_words[l], b.prefix[l], b.block_sums[l]);
}
__global__ void __launch_bounds__(1024)
scan_blocks_batch_kernel(const __grid_constant__ CoarsenBatch b) {
    pdl_wait();
    const int l = blockIdx.x;
    scan_blocks_body(b.block_sums[l], (int)((b.n_words[l] + SCAN_BLK - 1) / SCAN_BLK), b.total[l]);
}
__global__ void __launch_bounds__(256)
emit_coarse_batch_kernel(const __grid_constant__ CoarsenBatch b, long long cap) {
    pdl_wait();
    const int l = blockIdx.y;
    emit_coarse_body(b.bitmap[l], b.prefix[l], b.block_sums[l], b.n_words[l], b.L[l], cap, b.coarse[l]);
}

// positions and reference-style views of a level (scenemodeling.py:211-226):
// x_pts = idx * res + pts_min[b],  pts_min[b] = anchor_pts[first voxel of b] - idx[first] * res
__global__ void batch_origin_kernel(const float* __restrict__ anchor_pts, const int* __restrict__ idx3d,
                                    const long long* __restrict__ batch, long long n, float res,
                                    float* __restrict__ origin) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long b = batch[i];
    if (i == 0 || batch[i - 1] != b) {  // rows are sorted by batch: first voxel of this batch
#pragma unroll
        for (int d = 0; d < 3; ++d)
            origin[3 * b + d] = __fsub_rn(anchor_pts[3 * i + d], __fmul_rn((float)idx3d[3 * i + d], res));
    }
}

__global__ void __launch_bounds__(256)
level_points_kernel(const int* __restrict__ coords, long long n, const float* __restrict__ origin, float res,
                    float* __restrict__ pts, long long* __restrict__ idx_out, long long* __restrict__ batch_out) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int b = coords[4 * i];
    if (batch_out) batch_out[i] = b;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        int c = coords[4 * i + 1 + d];
        if (idx_out) idx_out[3 * i + d] = c;
        pts[3 * i + d] = __fadd_rn(__fmul_rn((float)c, res), origin[3 * b + d]);
    }
}

__global__ void __launch_bounds__(256)
make_coords_kernel(const int* __restrict__ idx3d, const long long* __restrict__ batch, long long n,
                   int* __restrict__ coords) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    coords[4 * i] = (int)batch[i];
    coords[4 * i + 1] = idx3d[3 * i];
    coords[4 * i + 2] = idx3d[3 * i + 1];
    coords[4 * i + 3] = idx3d[3 * i + 2];
}

// workspace carving shared by voxelize and coarsen
struct ScanSpace {
    unsigned* bitmap;
    unsigned* prefix;
    unsigned* block_sums;
    long long* total;
    int* err;
    char* extra;  // caller-specific tail
    size_t extra_off;
};
static size_t scan_space_bytes(long long n_words) {
    size_t nb = (size_t)cdiv(n_words, SCAN_BLK);
    return align_up((size_t)n_words * 4, 256) * 2 + align_up(nb * 4, 256) + 256;
}
static ScanSpace carve(void* ws, long long n_words) {
    ScanSpace s;
    char* p = (char*)ws;
    size_t nb = (size_t)cdiv(n_words, SCAN_BLK);
    s.bitmap = (unsigned*)p;
    p += align_up((size_t)n_words * 4, 256);
    s.prefix = (unsigned*)p;
    p += align_up((size_t)n_words * 4, 256);
    s.block_sums = (unsigned*)p;
    p += align_up(nb * 4, 256);
    s.total = (long long*)p;
    s.err = (int*)(p + 8);
    p += 256;
    s.extra = p;
    s.extra_off = (size_t)(p - (char*)ws);
    return s;
}

static int run_scan(const ScanSpace& s, long long n_words, cudaStream_t st) {
    int nb = cdiv(n_words, SCAN_BLK);
    DV3D_LAUNCH((scan_words_kernel), nb, SCAN_T, 0, st, s.bitmap, n_words, s.prefix, s.block_sums);
    DV3D_LAUNCHED();
    DV3D_LAUNCH((scan_blocks_kernel), 1, 1024, 0, st, s.block_sums, nb, s.total);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

}  // namespace dv3d

using namespace dv3d;

extern "C" int dv3d_voxel_grid(const float* pts, const long long* batch, long long N, float edge_len,
                               dv3d_voxel_grid_t* grid_host, void* scratch64, void* stream) {
    DV3D_REQUIRE(pts && batch && grid_host && scratch64 && N > 0 && edge_len > 0.f, "voxel_grid: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    BBoxHeader* h = (BBoxHeader*)scratch64;
    float* out8 = (float*)((char*)scratch64 + 32);
    DV3D_LAUNCH((bbox_init_kernel), 1, 32, 0, st, h);
    DV3D_LAUNCHED();
    int blocks = cdiv(N, 256 * 4);
    if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
    DV3D_LAUNCH((bbox_kernel), blocks, 256, 0, st, pts, batch, N, h);
    DV3D_LAUNCHED();
    DV3D_LAUNCH((bbox_decode_kernel), 1, 32, 0, st, h, out8);
    DV3D_LAUNCHED();
    float host8[8];
    {
        ReadItem it = {out8, host8, (int)sizeof(host8)};
        int rc = read_back(&it, 1, st);
        if (rc) return rc;
    }
    dv3d_voxel_grid_t g;
    memset(&g, 0, sizeof(g));
    g.edge_len = edge_len;
    g.total_cells = 1;
    for (int d = 0; d < 3; ++d) {
        g.bbox_min[d] = host8[d];
        g.bbox_max[d] = host8[3 + d];
        // fp32 subtract, fp32 divide as torch does on float32 tensors
        volatile float ext = g.bbox_max[d] - g.bbox_min[d];
        volatile float q = ext / edge_len;
        DV3D_REQUIRE(isfinite(q) && q < 1e9f, "voxel_grid: non-finite or absurd bounding box");
        g.n_cells[d] = (long long)q + 1;         // grid_cluster
        g.grid_size[d] = (long long)ceilf(q);    // utils.py:41
        g.total_cells *= g.n_cells[d];
    }
    g.n_batch = (long long)host8[6] + 1;
    g.total_cells *= g.n_batch;
    *grid_host = g;
    return DV3D_OK;
}

extern "C" size_t dv3d_voxelize_workspace_bytes(const dv3d_voxel_grid_t* grid, long long N) {
    if (!grid || N < 0) return 0;
    long long n_words = (grid->total_cells + 31) / 32;
    return scan_space_bytes(n_words) + align_up((size_t)N * 8, 256) + align_up((size_t)grid->n_batch * 12, 256);
}

extern "C" int dv3d_voxelize(const float* pts, const long long* batch, long long N, const dv3d_voxel_grid_t* grid,
                             void* workspace, size_t workspace_bytes, long long cap, long long* n_anchors_host,
                             float* anchor_pts, int* anchor_idx3d, long long* anchor_batch, int* point_anchor,
                             void* stream) {
    DV3D_REQUIRE(pts && batch && grid && workspace && n_anchors_host && anchor_pts && anchor_idx3d && anchor_batch &&
                     point_anchor && N > 0 && cap > 0,
                 "voxelize: bad arguments");
    DV3D_REQUIRE(grid->grid_size[0] > 0 && grid->grid_size[1] > 0 && grid->grid_size[2] > 0,
                 "voxelize: degenerate bounding box (zero extent): the reference divides by zero here "
                 "(utils.py:55-57)");
    DV3D_REQUIRE(grid->total_cells > 0 && grid->total_cells < (1ll << 36), "voxelize: %lld cells is out of range",
                 grid->total_cells);
    if (workspace_bytes < dv3d_voxelize_workspace_bytes(grid, N)) {
        set_error("voxelize: workspace of %zu bytes is too small, need %zu", workspace_bytes,
                  dv3d_voxelize_workspace_bytes(grid, N));
        return DV3D_ENOSPC;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_words = (grid->total_cells + 31) / 32;
    ScanSpace s = carve(workspace, n_words);
    long long* point_id = (long long*)s.extra;
    int* min_idx = (int*)(s.extra + align_up((size_t)N * 8, 256));

    GridDev G;
    for (int d = 0; d < 3; ++d) {
        G.bmin[d] = grid->bbox_min[d];
        G.n[d] = grid->n_cells[d];
        G.g[d] = grid->grid_size[d];
    }
    G.e = grid->edge_len;

    DV3D_CUDA(cudaMemsetAsync(workspace, 0, s.extra_off, st));                         // bitmap, sums, total, err
    DV3D_CUDA(cudaMemsetAsync(min_idx, 0x7f, (size_t)grid->n_batch * 12, st));        // > any index
    DV3D_LAUNCH((mark_points_kernel), cdiv(N, 256), 256, 0, st, pts, batch, N, G, grid->total_cells, point_id, s.bitmap, s.err);
    DV3D_LAUNCHED();
    int rc = run_scan(s, n_words, st);
    if (rc) return rc;
    DV3D_LAUNCH((emit_anchors_kernel), cdiv(n_words, 256), 256, 0, st, s.bitmap, s.prefix, s.block_sums, n_words, G, (float)(grid->edge_len / 2.0), cap, anchor_pts, anchor_idx3d, anchor_batch, min_idx);
    DV3D_LAUNCHED();
    long long host[2] = {0, 0};
    {
        ReadItem it = {s.total, host, 16};
        int rc = read_back(&it, 1, st);
        if (rc) return rc;
    }
    *n_anchors_host = host[0];
    DV3D_REQUIRE((int)(host[1] & 0xffffffff) == 0, "voxelize: a point fell outside the bounding-box grid (NaN input?)");
    if (host[0] > cap) {
        set_error("voxelize: %lld anchors exceed the output capacity %lld", host[0], cap);
        return DV3D_ENOSPC;
    }
    long long m = host[0] > N ? host[0] : N;
    DV3D_LAUNCH((finish_voxelize_kernel), cdiv(m, 256), 256, 0, st, host[0], N, anchor_batch, min_idx, anchor_idx3d, point_id, s.bitmap, s.prefix, s.block_sums, point_anchor);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_make_coords(const int* idx3d, const long long* batch, long long n, int* coords, void* stream) {
    DV3D_REQUIRE(idx3d && batch && coords && n >= 0, "make_coords: bad arguments");
    if (n == 0) return DV3D_OK;
    DV3D_LAUNCH((make_coords_kernel), cdiv(n, 256), 256, 0, (cudaStream_t)stream, idx3d, batch, n, coords);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" size_t dv3d_coarsen_workspace_bytes(int dim_x, int dim_y, int dim_z, int n_batch, int new_stride) {
    if (dim_x <= 0 || dim_y <= 0 || dim_z <= 0 || n_batch <= 0 || new_stride <= 0) return 0;
    long long X = cdiv(dim_x, new_stride), Y = cdiv(dim_y, new_stride), Z = cdiv(dim_z, new_stride);
    long long n_words = (X * Y * Z * n_batch + 31) / 32;
    return scan_space_bytes(n_words);
}

// the kernels of a coarsening step, without the read-back of the count
extern "C" int dv3d_coarsen_enqueue(const int* coords, long long n, int new_stride, int dim_x, int dim_y, int dim_z,
                                    int n_batch, void* workspace, size_t workspace_bytes, long long cap,
                                    int* coarse_coords, void* stream) {
    DV3D_REQUIRE(coords && workspace && coarse_coords && n > 0 && cap > 0 && new_stride > 0, "coarsen: bad arguments");
    size_t need = dv3d_coarsen_workspace_bytes(dim_x, dim_y, dim_z, n_batch, new_stride);
    DV3D_REQUIRE(need > 0, "coarsen: bad lattice dimensions");
    if (workspace_bytes < need) {
        set_error("coarsen: workspace of %zu bytes is too small, need %zu", workspace_bytes, need);
        return DV3D_ENOSPC;
    }
    cudaStream_t st = (cudaStream_t)stream;
    LevelDims L = {cdiv(dim_x, new_stride), cdiv(dim_y, new_stride), cdiv(dim_z, new_stride), new_stride};
    const long long n_words = ((long long)L.X * L.Y * L.Z * n_batch + 31) / 32;
    ScanSpace s = carve(workspace, n_words);
    DV3D_CUDA(cudaMemsetAsync(workspace, 0, s.extra_off, st));
    DV3D_LAUNCH((mark_coarse_kernel), cdiv(n, 256), 256, 0, st, coords, n, L, n_batch, s.bitmap, s.err);
    DV3D_LAUNCHED();
    int rc = run_scan(s, n_words, st);
    if (rc) return rc;
    DV3D_LAUNCH((emit_coarse_kernel), cdiv(n_words, 256), 256, 0, st, s.bitmap, s.prefix, s.block_sums, n_words, L, cap, coarse_coords);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

// several coarser levels from the same (finest) coordinates with four launches in total
extern "C" int dv3d_coarsen_enqueue_batch(const int* coords, long long n, const int* new_strides, int n_levels, int dim_x,
                                          int dim_y, int dim_z, int n_batch, void* const* workspaces,
                                          const size_t* workspace_bytes, long long cap, int* const* coarse_coords,
                                          void* stream) {
    DV3D_REQUIRE(coords && new_strides && workspaces && workspace_bytes && coarse_coords && n > 0 && cap > 0 && n_levels >= 1 &&
                     n_levels <= CB_MAX,
                 "coarsen_batch: bad arguments (1..%d levels)", CB_MAX);
    cudaStream_t st = (cudaStream_t)stream;
    CoarsenBatch b = {};
    long long max_words = 0;
    for (int l = 0; l < n_levels; ++l) {
        DV3D_REQUIRE(workspaces[l] && coarse_coords[l] && new_strides[l] > 0, "coarsen_batch: bad level %d", l);
        const size_t need = dv3d_coarsen_workspace_bytes(dim_x, dim_y, dim_z, n_batch, new_strides[l]);
        DV3D_REQUIRE(need > 0, "coarsen: bad lattice dimensions");
        if (workspace_bytes[l] < need) {
            set_error("coarsen: workspace of %zu bytes is too small, need %zu", workspace_bytes[l], need);
            return DV3D_ENOSPC;
        }
        b.L[l] = LevelDims{cdiv(dim_x, new_strides[l]), cdiv(dim_y, new_strides[l]), cdiv(dim_z, new_strides[l]), new_strides[l]};
        b.n_words[l] = ((long long)b.L[l].X * b.L[l].Y * b.L[l].Z * n_batch + 31) / 32;
        ScanSpace s = carve(workspaces[l], b.n_words[l]);
        b.bitmap[l] = s.bitmap;
        b.prefix[l] = s.prefix;
        b.block_sums[l] = s.block_sums;
        b.total[l] = s.total;
        b.err[l] = s.err;
        b.coarse[l] = coarse_coords[l];
        if (b.n_words[l] > max_words) max_words = b.n_words[l];
        // contiguous workspaces are cleared by the memset of the first one
        const bool merged = l > 0 && (char*)workspaces[l] == (char*)workspaces[l - 1] + workspace_bytes[l - 1];
        if (!merged) {
            size_t span = workspace_bytes[l];
            for (int k = l + 1; k < n_levels && (char*)workspaces[k] == (char*)workspaces[k - 1] + workspace_bytes[k - 1]; ++k)
                span += workspace_bytes[k];
            DV3D_CUDA(cudaMemsetAsync(workspaces[l], 0, span, st));
        }
    }
    DV3D_LAUNCH((mark_coarse_batch_kernel), dim3(cdiv(n, 256), n_levels), 256, 0, st, coords, n, n_batch, b);
    DV3D_LAUNCHED();
    DV3D_LAUNCH((scan_words_batch_kernel), dim3(cdiv(max_words, SCAN_BLK), n_levels), SCAN_T, 0, st, b);
    DV3D_LAUNCHED();
    DV3D_LAUNCH((scan_blocks_batch_kernel), n_levels, 1024, 0, st, b);
    DV3D_LAUNCHED();
    DV3D_LAUNCH((emit_coarse_batch_kernel), dim3(cdiv(max_words, 256), n_levels), 256, 0, st, b, cap);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

// reads the count of an enqueued coarsening step back (SYNCS)
extern "C" int dv3d_coarsen_finish(const void* workspace, int new_stride, int dim_x, int dim_y, int dim_z, int n_batch,
                                   long long cap, long long* n_coarse_host, void* stream) {
    DV3D_REQUIRE(workspace && n_coarse_host && new_stride > 0, "coarsen_finish: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    LevelDims L = {cdiv(dim_x, new_stride), cdiv(dim_y, new_stride), cdiv(dim_z, new_stride), new_stride};
    const long long n_words = ((long long)L.X * L.Y * L.Z * n_batch + 31) / 32;
    ScanSpace s = carve(const_cast<void*>(workspace), n_words);
    long long host[2] = {0, 0};
    {
        ReadItem it = {s.total, host, 16};
        int rc = read_back(&it, 1, st);
        if (rc) return rc;
    }
    *n_coarse_host = host[0];
    DV3D_REQUIRE((int)(host[1] & 0xffffffff) == 0, "coarsen: a coordinate lies outside the declared lattice");
    if (host[0] > cap) {
        set_error("coarsen: %lld coarse voxels exceed the output capacity %lld", host[0], cap);
        return DV3D_ENOSPC;
    }
    return DV3D_OK;
}

// the counts of several enqueued coarsening steps with ONE read-back
extern "C" int dv3d_coarsen_finish_batch(const void* const* workspaces, const int* new_strides, int n_levels, int dim_x,
                                         int dim_y, int dim_z, int n_batch, long long cap, long long* n_coarse_host,
                                         void* stream) {
    DV3D_REQUIRE(workspaces && new_strides && n_coarse_host && n_levels >= 1 && n_levels <= CB_MAX,
                 "coarsen_finish_batch: bad arguments (1..%d levels)", CB_MAX);
    long long host[CB_MAX][2] = {};
    ReadItem items[CB_MAX];
    for (int l = 0; l < n_levels; ++l) {
        DV3D_REQUIRE(workspaces[l] && new_strides[l] > 0, "coarsen_finish_batch: bad level %d", l);
        LevelDims L = {cdiv(dim_x, new_strides[l]), cdiv(dim_y, new_strides[l]), cdiv(dim_z, new_strides[l]), new_strides[l]};
        const long long n_words = ((long long)L.X * L.Y * L.Z * n_batch + 31) / 32;
        ScanSpace s = carve(const_cast<void*>(workspaces[l]), n_words);
        items[l] = ReadItem{s.total, host[l], 16};
    }
    int rc = read_back(items, n_levels, (cudaStream_t)stream);
    if (rc) return rc;
    for (int l = 0; l < n_levels; ++l) {
        n_coarse_host[l] = host[l][0];
        DV3D_REQUIRE((int)(host[l][1] & 0xffffffff) == 0, "coarsen: a coordinate lies outside the declared lattice");
        if (host[l][0] > cap) {
            set_error("coarsen: %lld coarse voxels exceed the output capacity %lld", host[l][0], cap);
            return DV3D_ENOSPC;
        }
    }
    return DV3D_OK;
}

extern "C" int dv3d_coarsen(const int* coords, long long n, int new_stride, int dim_x, int dim_y, int dim_z,
                            int n_batch, void* workspace, size_t workspace_bytes, long long cap, int* coarse_coords,
                            long long* n_coarse_host, void* stream) {
    DV3D_REQUIRE(n_coarse_host, "coarsen: bad arguments");
    int rc = dv3d_coarsen_enqueue(coords, n, new_stride, dim_x, dim_y, dim_z, n_batch, workspace, workspace_bytes, cap,
                                  coarse_coords, stream);
    if (rc) return rc;
    return dv3d_coarsen_finish(workspace, new_stride, dim_x, dim_y, dim_z, n_batch, cap, n_coarse_host, stream);
}

extern "C" int dv3d_batch_origin(const float* anchor_pts, const int* idx3d, const long long* batch, long long n,
                                 float res, float* origin, void* stream) {
    DV3D_REQUIRE(anchor_pts && idx3d && batch && origin && n > 0, "batch_origin: bad arguments");
    DV3D_LAUNCH((batch_origin_kernel), cdiv(n, 256), 256, 0, (cudaStream_t)stream, anchor_pts, idx3d, batch, n, res, origin);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_level_points(const int* coords, long long n, const float* origin, float res, float* pts,
                                 long long* idx_out, long long* batch_out, void* stream) {
    DV3D_REQUIRE(coords && origin && pts && n >= 0, "level_points: bad arguments");
    if (n == 0) return DV3D_OK;
    DV3D_LAUNCH((level_points_kernel), cdiv(n, 256), 256, 0, (cudaStream_t)stream, coords, n, origin, res, pts, idx_out, batch_out);
    DV3D_LAUNCHED();
    return DV3D_OK;
}
