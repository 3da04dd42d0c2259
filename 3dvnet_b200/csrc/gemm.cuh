// Gather-GEMM: the one dense contraction behind every learned layer of the volumetric
// refinement path.
//
//     out[m, :] = epilogue( sum_s  A_s[row_s(m), 0:K_s] @ W[off_s : off_s + K_s, :] )
//
// A "slice" s names a source matrix, how output row m maps to a source row (an index table
// or a constant shift; a negative / out-of-range row contributes zero) and the K extent of
// its weight block.  That covers
//   * sparse 3x3x3 convolutions     27 slices, row_s(m) = kernel_map[m][s]       (scenemodeling.py:27-39)
//   * Conv1d(k=3) over hypotheses    3 slices, row_s(m) = m + s - 1               (refinement.py:8-13)
//   * ME.cat + 1x1 convolution       2 slices, identity rows                      (scenemodeling.py:206)
//   * PointNet Linear on [x|pool[v]] 2 slices, second gathered by voxel id        (scenemodeling.py:130-141)
// The epilogue fuses bias / folded BatchNorm, per-row GroupNorm (16-channel groups),
// residual add, ReLU and the zeroing of padding rows.
#pragma once
#include "common.cuh"

namespace dv3d {

constexpr int kMaxSlices = 27;

struct GemmSlice {
    const float* src;   // [rows, ld]
    const int* idx;     // row table (idx[m * idx_stride]) or nullptr -> row = m + shift
    int idx_stride;
    int shift;
    int ld;             // row pitch of src in floats
    int K;              // multiple of 16
};

constexpr int kMaxPeers = 7;
struct GemmDesc {
    GemmSlice slice[kMaxSlices];
    int n_slices;
    long long M;            // output rows
    long long n_src_rows;   // shifted (non-table) rows outside [0, n_src_rows) read as zero
    int N;                  // output channels: 64 or 128
    const float* W;         // [sum_s K_s, N] row-major (fp32 CUDA-core kernel)
    const float* Wp;        // packed by dv3d_gemm_pack_weights: selects the tcgen05 kernel when set
    const int* kmap;        // set when every slice s gathers through kmap[m * n_slices + s]
    const int* tile_wslice; // tcgen05 kernel, n_slices == 1: 128-row tile t contracts with the weight block
                            // W[tile_wslice[t] * K : (tile_wslice[t] + 1) * K, :] (pair-major sparse convolution)
    const float* scale;     // per-channel multiplier (folded BN) or nullptr
    const float* shift;     // per-channel addend (bias / folded BN) or nullptr
    const float* gn_weight; // GroupNorm affine (groups of 16 channels) or nullptr
    const float* gn_bias;
    const float* residual;  // added before the ReLU, [M, res_ld], or nullptr
    int res_ld;
    int relu_in;            // ReLU on the gathered inputs  (fc(relu(x)))
    int relu_out;
    int zero_row_mod;       // rows with m % zero_row_mod == zero_row_val are written as 0
    int zero_row_val;
    float* split_ws;        // optional K-split workspace of the tcgen05 kernel (see gemm_tc.cu)
    size_t split_ws_bytes;
    int* split_counters;    // one int per 128-row tile, zero on entry, left zero on exit
    int split_hint;         // > 1: use exactly this many K splits (caller sized the workspace); 0: automatic
    float* out;             // [M, out_ld]
    int out_ld;
    // Row-sharded layers of a scene spanning GPUs (symm.cu): the finished rows are also stored into the same rows
    // of every peer's copy of the buffer over NVLink, straight from the epilogue.  Filled in by the launchers from
    // the registered symmetric regions (symm_attach); callers leave it zero.
    float* peer_out[kMaxPeers];
    int n_peers;
    // Halo exchange: when set, peer p only receives the rows lo <= row_base + m < hi with lo = peer_rows[2 p],
    // hi = peer_rows[2 p + 1] (device ints: the rows of this level that rank gathers through its kernel maps).
    // nullptr: every row goes to every peer.  Set by symm_set_halo for the launches that follow.
    const int* peer_rows;
    long long row_base;     // global row index of out row 0
    // Optional fused segment max (PointNet, scenemodeling.py:129): every finished row m is also max-reduced into
    // pool_out[pool_seg[m], :] (row pitch = out_ld).  pool_out must be filled with 0xFF bytes beforehand (see
    // atomic_max_f32) and every segment must own at least one row.
    float* pool_out;
    const int* pool_seg;
};

// max(*addr, v) on a float slot initialised to 0xFFFFFFFF (a negative NaN: the smallest int, the largest unsigned):
// non-negative floats order like ints, negative floats like unsigned ints reversed.  Exact and order independent.
__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
    v += 0.f;  // -0 -> +0
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

// Epilogue of one 16-byte unit (4 channels c0..c0+3 of output row m).  The 4 lanes that hold
// the 16 channels of a GroupNorm group are consecutive and aligned, so the group statistics are
// two xor-shuffles; every lane of the warp must call this (live = false rows only skip the store).
__device__ __forceinline__ void epilogue4(const GemmDesc& d, float4 y, long long m, int c0, bool live, bool zero_row) {
    if (d.scale) {
        const float4 s = __ldg(reinterpret_cast<const float4*>(d.scale + c0));
        y.x *= s.x; y.y *= s.y; y.z *= s.z; y.w *= s.w;
    }
    if (d.shift) {
        const float4 s = __ldg(reinterpret_cast<const float4*>(d.shift + c0));
        y.x += s.x; y.y += s.y; y.z += s.z; y.w += s.w;
    }
    if (d.gn_weight) {
        float sum = (y.x + y.y) + (y.z + y.w);
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const float mean = sum * (1.f / 16.f);
        const float dx = y.x - mean, dy = y.y - mean, dz = y.z - mean, dw = y.w - mean;
        float q = fmaf(dx, dx, dy * dy) + fmaf(dz, dz, dw * dw);
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        const float rstd = 1.f / sqrtf(q * (1.f / 16.f) + 1e-5f);
        const float4 gw = __ldg(reinterpret_cast<const float4*>(d.gn_weight + c0));
        const float4 gb = __ldg(reinterpret_cast<const float4*>(d.gn_bias + c0));
        y.x = fmaf(dx * rstd, gw.x, gb.x);
        y.y = fmaf(dy * rstd, gw.y, gb.y);
        y.z = fmaf(dz * rstd, gw.z, gb.z);
        y.w = fmaf(dw * rstd, gw.w, gb.w);
    }
    if (!live) return;
    if (d.residual) {
        const float4 rv = __ldg(reinterpret_cast<const float4*>(d.residual + (size_t)m * d.res_ld + c0));
        y.x += rv.x; y.y += rv.y; y.z += rv.z; y.w += rv.w;
    }
    if (d.relu_out) {
        y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f);
    }
    if (zero_row) y = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t at = (size_t)m * d.out_ld + c0;
    *reinterpret_cast<float4*>(d.out + at) = y;
    if (d.pool_out) {
        float* o = d.pool_out + (size_t)__ldg(d.pool_seg + m) * d.out_ld + c0;
        atomic_max_f32(o, y.x);
        atomic_max_f32(o + 1, y.y);
        atomic_max_f32(o + 2, y.z);
        atomic_max_f32(o + 3, y.w);
    }
    if (d.n_peers) {
        const long long g = d.row_base + m;
        for (int p = 0; p < d.n_peers; ++p)
            if (!d.peer_rows || (g >= __ldg(d.peer_rows + 2 * p) && g < __ldg(d.peer_rows + 2 * p + 1)))
                *reinterpret_cast<float4*>(d.peer_out[p] + at) = y;
    }
}

// d.Wp == nullptr: fp32 CUDA-core kernel (gemm.cu); otherwise the tcgen05 kernel (gemm_tc.cu),
// 3xTF32 or TF32 according to dv3d_set_gemm_precision.
// fills d.peer_out / d.n_peers when d.out lies inside a region registered with dv3d_symm_register
void symm_attach(GemmDesc& d);
// rows that go to each peer for the launches that follow on this host thread (nullptr: all rows); see GemmDesc
void symm_set_halo(const int* peer_rows, long long row_base);
int launch_gather_gemm(const GemmDesc& d, cudaStream_t st);
int launch_gather_gemm_tc(const GemmDesc& d, cudaStream_t st);
int gather_gemm_tc_splits(long long M, int n_slices);

}  // namespace dv3d
