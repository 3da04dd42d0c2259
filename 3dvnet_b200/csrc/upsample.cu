// Coarse-to-fine depth upsampling: PropagationNet (upsampling.py:14-36) and the nearest
// upsampling in front of it (eval-3dvnet.py:101-125, lightningmodel.py:84-112) - the step
// immediately after the hot path.
//
// The four Conv2d(3x3)+BN+ReLU layers are 9-slice gather-GEMMs on the tcgen05 kernel
// (gemm_tc.cu): activations are channels-last rows [pixel][C], the slice of tap t gathers the
// row of pixel (y+dy, x+dx) through a dense-grid neighbour table (-1 = zero padding), weights
// are [9*Cin, 64] with the output channels zero-padded to the kernel's 64-column tile, folded
// BatchNorm + ReLU in the epilogue.  Two small kernels surround them: the input rows
// [features | nearest-upsampled depth | 0] and the softmax over the 9 logits + weighted sum of
// the replicate-padded 3x3 depth neighbourhood.
#include <math.h>

#include "gemm.cuh"

namespace dv3d {

// nbr[(n*H+y)*W+x][t] = row of pixel (y+dy, x+dx), t = (dy+1)*3 + (dx+1), or -1 outside the image
__global__ void __launch_bounds__(256)
grid_map_2d_kernel(int n, int H, int W, int* __restrict__ nbr) {
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long M = (long long)n * H * W;
    if (i >= M * 9) return;
    const long long m = i / 9;
    const int t = (int)(i - m * 9);
    const int x = (int)(m % W), y = (int)((m / W) % H);
    const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
    nbr[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (int)(m + (long long)(t / 3 - 1) * W + (t % 3 - 1)) : -1;
}

// F.interpolate(mode='nearest') source index: min(floor(dst * (in / out)), in - 1) in fp32
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
    const int s = (int)floorf(__fmul_rn((float)dst, scale));
    return s < in_size - 1 ? s : in_size - 1;
}

// x[m][0..C-1] = features[n][c][y][x], x[m][C] = depth_up[m] = depth_lo[n][src(y)][src(x)], x[m][C+1..ld-1] = 0
__global__ void __launch_bounds__(256)
propagation_input_kernel(const float* __restrict__ feats, int C, const float* __restrict__ depth_lo, int n, int h, int w,
                         int H, int W, float sy, float sx, int ld, float* __restrict__ x, float* __restrict__ depth_up) {
    pdl_wait();
    __shared__ float tile[32][33];
    const int nn = blockIdx.z;
    const long long HW = (long long)H * W;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    // coalesced read of 32 channels x 32 pixels, transposed write of 32 pixels x 32 channels
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i;
        const long long p = p0 + threadIdx.x;
        float v = 0.f;
        if (p < HW) {
            if (c < C) {
                v = __ldg(feats + ((size_t)nn * C + c) * HW + p);
            } else if (c == C) {
                const int y = (int)(p / W), xx = (int)(p % W);
                v = __ldg(depth_lo + ((size_t)nn * h + nearest_src(y, sy, h)) * w + nearest_src(xx, sx, w));
                depth_up[(size_t)nn * HW + p] = v;
            }
        }
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long p = p0 + i;
        const int c = c0 + threadIdx.x;
        if (p < HW && c < ld) x[((size_t)nn * HW + p) * ld + c] = tile[threadIdx.x][i];
    }
}

// out = sum_t softmax(logits[m][0..8])_t * depth(y+dy, x+dx) with replicate padding (upsampling.py:26-36)
__global__ void __launch_bounds__(256)
propagation_output_kernel(const float* __restrict__ logits, int ld, const float* __restrict__ depth, int n, int H, int W,
                          float* __restrict__ out) {
    pdl_wait();
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= (long long)n * H * W) return;
    const int x = (int)(m % W), y = (int)((m / W) % H);
    const float* dn = depth + (m - (long long)y * W - x);
    float l[9];
    const float4 a = __ldg(reinterpret_cast<const float4*>(logits + (size_t)m * ld));
    const float4 b = __ldg(reinterpret_cast<const float4*>(logits + (size_t)m * ld) + 1);
    l[0] = a.x; l[1] = a.y; l[2] = a.z; l[3] = a.w; l[4] = b.x; l[5] = b.y; l[6] = b.z; l[7] = b.w;
    l[8] = __ldg(logits + (size_t)m * ld + 8);
    float mx = l[0];
#pragma unroll
    for (int t = 1; t < 9; ++t) mx = fmaxf(mx, l[t]);
    float s = 0.f, acc = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const float e = expf(l[t] - mx);
        const int yy = min(max(y + t / 3 - 1, 0), H - 1), xx = min(max(x + t % 3 - 1, 0), W - 1);
        s += e;
        acc = fmaf(e, __ldg(dn + (size_t)yy * W + xx), acc);
    }
    out[m] = acc / s;
}

}  // namespace dv3d

using namespace dv3d;

extern "C" int dv3d_grid_map_2d(int n, int H, int W, int* nbr, void* stream) {
    DV3D_REQUIRE(nbr && n >= 0 && H > 0 && W > 0 && (long long)n * H * W < (1ll << 31), "grid_map_2d: bad arguments");
    if (n == 0) return DV3D_OK;
    DV3D_LAUNCH((grid_map_2d_kernel), cdiv((long long)n * H * W * 9, 256), 256, 0, (cudaStream_t)stream, n, H, W, nbr);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_propagation_input(const float* feats_nchw, int C, const float* depth_lo, int n, int h, int w, int H,
                                      int W, int ld, float* x, float* depth_up, void* stream) {
    DV3D_REQUIRE(feats_nchw && depth_lo && x && depth_up && n >= 0 && C > 0 && h > 0 && w > 0 && H > 0 && W > 0,
                 "propagation_input: bad arguments");
    DV3D_REQUIRE(ld % 32 == 0 && ld > C, "propagation_input: ld must be a multiple of 32 larger than C (C=%d ld=%d)", C, ld);
    if (n == 0) return DV3D_OK;
    DV3D_REQUIRE(n <= 65535, "propagation_input: n > 65535");
    dim3 grid(cdiv((long long)H * W, 32), ld / 32, n), block(32, 8);
    // ATen: scale = (float)input_size / output_size
    DV3D_LAUNCH((propagation_input_kernel), grid, block, 0, (cudaStream_t)stream, feats_nchw, C, depth_lo, n, h, w, H, W,
                (float)h / (float)H, (float)w / (float)W, ld, x, depth_up);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_conv2d3x3_bn_relu_rows(const float* x, long long M, int Cin, int ldx, const int* nbr, const float* W_kn,
                                           const void* W_packed, const float* scale, const float* shift, int Cout,
                                           float* y, int ldy, void* stream) {
    DV3D_REQUIRE(x && nbr && (W_kn || W_packed) && scale && shift && y && M >= 0, "conv2d3x3: bad arguments");
    DV3D_REQUIRE(Cout == 64 || Cout == 128, "conv2d3x3: the output tile is 64 or 128 columns (zero-pad the layer), got %d", Cout);
    GemmDesc d = {};
    d.n_slices = 9;
    for (int t = 0; t < 9; ++t) d.slice[t] = GemmSlice{x, nbr + t, 9, 0, ldx, Cin};
    d.kmap = nbr;
    d.W = W_kn;
    d.Wp = (const float*)W_packed;
    d.M = M;
    d.n_src_rows = M;
    d.N = Cout;
    d.scale = scale;
    d.shift = shift;
    d.relu_out = 1;
    d.out = y;
    d.out_ld = ldy;
    return launch_gather_gemm(d, (cudaStream_t)stream);
}

extern "C" int dv3d_propagation_output(const float* logits, int ld, const float* depth, int n, int H, int W, float* out,
                                       void* stream) {
    DV3D_REQUIRE(logits && depth && out && n >= 0 && H > 0 && W > 0 && ld >= 12 && ld % 4 == 0 && ((uintptr_t)logits & 15) == 0,
                 "propagation_output: bad arguments");
    if (n == 0) return DV3D_OK;
    DV3D_LAUNCH((propagation_output_kernel), cdiv((long long)n * H * W, 256), 256, 0, (cudaStream_t)stream, logits, ld, depth,
                n, H, W, out);
    DV3D_LAUNCHED();
    return DV3D_OK;
}
