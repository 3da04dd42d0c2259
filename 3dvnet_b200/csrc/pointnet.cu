// Per-voxel PointNet (scenemodeling.py:116-144, lightningmodel.py:182): input assembly,
// Linear layers on [x | pool[voxel]] through the gather-GEMM, and the per-voxel max pool.
#include <math.h>

#include "gemm.cuh"

namespace dv3d {

// rows [pts - anchor_pts[seg] | pts_feat | 0-padding]: out [N, ld]
__global__ void __launch_bounds__(256)
pointnet_input_kernel(const float* __restrict__ pts, const float* __restrict__ pts_feat, int feat_ld,
                      const float* __restrict__ anchor_pts, const int* __restrict__ seg, long long N, int C, int ld,
                      float* __restrict__ out) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * ld) return;
    long long p = i / ld;
    int c = (int)(i - p * ld);
    float v = 0.f;
    if (c < 3)
        v = __fsub_rn(__ldg(pts + 3 * p + c), __ldg(anchor_pts + 3 * (long long)__ldg(seg + p) + c));
    else if (c < 3 + C)
        v = __ldg(pts_feat + p * feat_ld + (c - 3));
    out[i] = v;
}

__global__ void __launch_bounds__(256)
fill_kernel(unsigned* p, long long n, unsigned v) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// torch_scatter 'max': out[seg[i], c] = max_i x[i, c]; order-encoded atomicMax, decoded afterwards
__global__ void __launch_bounds__(256)
segment_max_kernel(const float* __restrict__ x, const int* __restrict__ seg, long long N, int C,
                   unsigned* __restrict__ out) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * (C / 4)) return;
    long long p = i / (C / 4);
    int c4 = (int)(i - p * (C / 4));
    float4 v = __ldg(reinterpret_cast<const float4*>(x + p * C) + c4);
    unsigned* o = out + (long long)__ldg(seg + p) * C + c4 * 4;
    atomicMax(o + 0, f2ord(v.x));
    atomicMax(o + 1, f2ord(v.y));
    atomicMax(o + 2, f2ord(v.z));
    atomicMax(o + 3, f2ord(v.w));
}

__global__ void __launch_bounds__(256)
segment_max_decode_kernel(unsigned* __restrict__ out, long long n) {
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned u = out[i];
    reinterpret_cast<float*>(out)[i] = (u == 0u) ? 0.f : ord2f(u);  // empty segments are 0 (torch_scatter)
}

}  // namespace dv3d

using namespace dv3d;

extern "C" int dv3d_pointnet_input(const float* pts, const float* pts_feat, int feat_ld, const float* anchor_pts,
                                   const int* seg, long long N, int C, int out_ld, float* out, void* stream) {
    DV3D_REQUIRE(pts && pts_feat && anchor_pts && seg && out && N >= 0 && C > 0 && out_ld >= 3 + C && feat_ld >= C,
                 "pointnet_input: bad arguments");
    if (N == 0) return DV3D_OK;
    DV3D_LAUNCH((pointnet_input_kernel), cdiv(N * out_ld, 256), 256, 0, (cudaStream_t)stream, pts, pts_feat, feat_ld, anchor_pts, seg, N, C, out_ld, out);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_linear(const float* x_a, int Ca, int lda, const float* pool, const int* seg, int Cb, long long N,
                           const float* weight_kn, const void* W_packed, const float* bias, int Cout, int relu_input,
                           float* y, void* stream) {
    DV3D_REQUIRE(x_a && (weight_kn || W_packed) && y && N >= 0 && Ca > 0 && Cb >= 0, "linear: bad arguments");
    DV3D_REQUIRE(Cb == 0 || (pool && seg), "linear: the pooled half needs pool and seg");
    GemmDesc d = {};
    d.n_slices = Cb ? 2 : 1;
    d.slice[0] = GemmSlice{x_a, nullptr, 0, 0, lda, Ca};
    if (Cb) d.slice[1] = GemmSlice{pool, seg, 1, 0, Cb, Cb};
    d.M = N;
    d.n_src_rows = N;
    d.N = Cout;
    d.W = weight_kn;
    d.Wp = (const float*)W_packed;
    d.shift = bias;
    d.relu_in = relu_input;
    d.out = y;
    d.out_ld = Cout;
    return launch_gather_gemm(d, (cudaStream_t)stream);
}

extern "C" int dv3d_linear_pool(const float* x_a, int Ca, int lda, const float* pool, const int* seg, int Cb, long long N,
                                const float* weight_kn, const void* W_packed, const float* bias, int Cout, int relu_input,
                                float* y, float* pool_out, const int* pool_seg, void* stream) {
    DV3D_REQUIRE(x_a && (weight_kn || W_packed) && y && N >= 0 && Ca > 0 && Cb >= 0, "linear_pool: bad arguments");
    DV3D_REQUIRE(Cb == 0 || (pool && seg), "linear_pool: the pooled half needs pool and seg");
    DV3D_REQUIRE(pool_out && pool_seg, "linear_pool: pool_out (0xFF-filled) and pool_seg are required");
    GemmDesc d = {};
    d.n_slices = Cb ? 2 : 1;
    d.slice[0] = GemmSlice{x_a, nullptr, 0, 0, lda, Ca};
    if (Cb) d.slice[1] = GemmSlice{pool, seg, 1, 0, Cb, Cb};
    d.M = N;
    d.n_src_rows = N;
    d.N = Cout;
    d.W = weight_kn;
    d.Wp = (const float*)W_packed;
    d.shift = bias;
    d.relu_in = relu_input;
    d.out = y;
    d.out_ld = Cout;
    d.pool_out = pool_out;
    d.pool_seg = pool_seg;
    return launch_gather_gemm(d, (cudaStream_t)stream);
}

extern "C" int dv3d_segment_max(const float* x, const int* seg, long long N, int C, long long n_seg, float* out,
                                void* stream) {
    DV3D_REQUIRE(x && seg && out && N >= 0 && n_seg >= 0 && C > 0 && C % 4 == 0, "segment_max: bad arguments");
    if (n_seg == 0) return DV3D_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DV3D_CUDA(cudaMemsetAsync(out, 0, (size_t)n_seg * C * 4, st));  // 0 orders below every encoded float
    if (N > 0) {
        DV3D_LAUNCH((segment_max_kernel), cdiv(N * (C / 4), 256), 256, 0, st, x, seg, N, C, reinterpret_cast<unsigned*>(out));
        DV3D_LAUNCHED();
    }
    DV3D_LAUNCH((segment_max_decode_kernel), cdiv(n_seg * C, 256), 256, 0, st, reinterpret_cast<unsigned*>(out), n_seg * C);
    DV3D_LAUNCHED();
    return DV3D_OK;
}
