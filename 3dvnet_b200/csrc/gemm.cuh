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

struct GemmDesc {
    GemmSlice slice[kMaxSlices];
    int n_slices;
    long long M;            // output rows
    long long n_src_rows;   // shifted (non-table) rows outside [0, n_src_rows) read as zero
    int N;                  // output channels: 64 or 128
    const float* W;         // [sum_s K_s, N] row-major (fp32 CUDA-core kernel)
    const float* Wp;        // packed by dv3d_gemm_pack_weights: selects the tcgen05 kernel when set
    const int* kmap;        // set when every slice s gathers through kmap[m * n_slices + s]
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
    float* out;             // [M, out_ld]
    int out_ld;
};

// d.Wp == nullptr: fp32 CUDA-core kernel (gemm.cu); otherwise the tcgen05 kernel (gemm_tc.cu),
// 3xTF32 or TF32 according to dv3d_set_gemm_precision.
int launch_gather_gemm(const GemmDesc& d, cudaStream_t st);
int launch_gather_gemm_tc(const GemmDesc& d, cudaStream_t st);
int gather_gemm_tc_splits(long long M, int n_slices);

}  // namespace dv3d
