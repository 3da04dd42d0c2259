// CostRegNet layers (inference): 3x3x3 convolutions / transposed convolutions with folded
// BatchNorm + ReLU (+ skip), and the fused prob-conv + softmax(-x) + depth expectation.
// Replaces the cuDNN Conv3d/ConvTranspose3d + BatchNorm3d + ReLU + softmax chain of
// mvsnet.py:18-36,133-163,219-227.  NCDHW fp32, fp32 accumulation on the CUDA cores.
//
// Two kernels:
//  * conv3d_s1_tiled: stride-1 layers on large volumes (conv0 = 68 % of the FLOPs, conv2).
//    CTA tile 4x14x28 outputs x 8 output channels; the haloed input tile of 8 input channels
//    and the matching weights are staged in shared memory; each thread owns a 4(x) x 2(y)
//    x 8(co) register tile = 64 accumulators, 576 FMAs per 26 shared loads.
//  * conv3d_generic: stride-2, transposed and small-volume layers: one thread per output
//    voxel x CO_T output channels, weights of the channel group in shared memory.
#include "common.cuh"

namespace dv3d {

// ------------------------------------------------------------------ tiled stride-1 conv
constexpr int TZ = 4, TY = 14, TX = 28;            // output tile
constexpr int IZ = TZ + 2, IY = TY + 2, IXP = 32;  // haloed input tile, row pitch padded 30 -> 32
constexpr int CIC = 8;                             // input channels staged per pass
constexpr int COT = 8;                             // output channels per CTA
constexpr int S1_THREADS = (TX / 4) * (TY / 2) * TZ;  // 7 * 7 * 4 = 196
constexpr size_t S1_SMEM = sizeof(float) * (CIC * IZ * IY * IXP + CIC * 27 * COT);

__global__ void __launch_bounds__(S1_THREADS, 2)
conv3d_s1_tiled_kernel(const float* __restrict__ x, int Cin, int D, int H, int W, const float* __restrict__ wgt,
                       const float* __restrict__ scale, const float* __restrict__ shift, int Cout,
                       const float* __restrict__ skip, float* __restrict__ y, int tiles_x, int tiles_y) {
    extern __shared__ __align__(16) float smem[];
    float* s_in = smem;                          // [CIC][IZ][IY][IXP]
    float* s_w = smem + CIC * IZ * IY * IXP;     // [CIC][27][COT]

    const int tid = threadIdx.x;
    const int tx = tid % (TX / 4), ty = (tid / (TX / 4)) % (TY / 2), tz = tid / ((TX / 4) * (TY / 2));
    const int bx = blockIdx.x % tiles_x, by = blockIdx.x / tiles_x;
    const int x0 = bx * TX, y0 = by * TY, z0 = blockIdx.y * TZ;
    const int cog = blockIdx.z % (Cout / COT), n = blockIdx.z / (Cout / COT);
    const size_t plane = (size_t)H * W, vol = plane * D;
    const float* xn = x + (size_t)n * Cin * vol;

    float acc[2][4][COT];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int c = 0; c < COT; ++c) acc[a][b][c] = 0.f;

    for (int c0 = 0; c0 < Cin; c0 += CIC) {
        __syncthreads();  // previous pass consumed
        // stage the haloed input tile (zero outside the volume)
        for (int i = tid; i < CIC * IZ * IY * (TX + 2); i += S1_THREADS) {
            int ix = i % (TX + 2);
            int r = i / (TX + 2);
            int iy = r % IY;
            r /= IY;
            int iz = r % IZ, ci = r / IZ;
            int gx = x0 + ix - 1, gy = y0 + iy - 1, gz = z0 + iz - 1;
            float v = 0.f;
            if (gx >= 0 && gx < W && gy >= 0 && gy < H && gz >= 0 && gz < D)
                v = __ldg(xn + (size_t)(c0 + ci) * vol + (size_t)gz * plane + (size_t)gy * W + gx);
            s_in[((ci * IZ + iz) * IY + iy) * IXP + ix] = v;
        }
        // weights of this pass: s_w[ci][tap][co] = wgt[cog*8+co][c0+ci][tap]
        for (int i = tid; i < CIC * 27 * COT; i += S1_THREADS) {
            int co = i % COT, tap = (i / COT) % 27, ci = i / (COT * 27);
            s_w[i] = __ldg(wgt + ((size_t)(cog * COT + co) * Cin + c0 + ci) * 27 + tap);
        }
        __syncthreads();

        for (int ci = 0; ci < CIC; ++ci) {
#pragma unroll
            for (int kd = 0; kd < 3; ++kd) {
                float in[4][6];
                const float* row = s_in + ((ci * IZ + tz + kd) * IY + 2 * ty) * IXP + 4 * tx;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    float4 a = *reinterpret_cast<const float4*>(row + r * IXP);
                    float2 b = *reinterpret_cast<const float2*>(row + r * IXP + 4);
                    in[r][0] = a.x; in[r][1] = a.y; in[r][2] = a.z; in[r][3] = a.w; in[r][4] = b.x; in[r][5] = b.y;
                }
                const float* wp = s_w + (ci * 27 + kd * 9) * COT;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        float4 w0 = *reinterpret_cast<const float4*>(wp + (kh * 3 + kw) * COT);
                        float4 w1 = *reinterpret_cast<const float4*>(wp + (kh * 3 + kw) * COT + 4);
                        float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                        for (int a = 0; a < 2; ++a)
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                float v = in[a + kh][b + kw];
#pragma unroll
                                for (int c = 0; c < COT; ++c) acc[a][b][c] = fmaf(v, wv[c], acc[a][b][c]);
                            }
                    }
                }
            }
        }
    }

    // epilogue: folded BN, ReLU, optional skip, store
    const int gz = z0 + tz;
    if (gz >= D) return;
    const size_t ovol = vol;
    float* yn = y + (size_t)n * Cout * ovol;
    const float* sn = skip ? skip + (size_t)n * Cout * ovol : nullptr;
#pragma unroll
    for (int c = 0; c < COT; ++c) {
        const int co = cog * COT + c;
        const float sc = __ldg(scale + co), sh = __ldg(shift + co);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int gy = y0 + 2 * ty + a;
            if (gy >= H) continue;
            const size_t base = (size_t)co * ovol + (size_t)gz * plane + (size_t)gy * W + x0 + 4 * tx;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if (x0 + 4 * tx + b < W) {
                    float v = fmaxf(fmaf(acc[a][b][c], sc, sh), 0.f);
                    if (sn) v += __ldg(sn + base + b);
                    yn[base + b] = v;
                }
            }
        }
    }
}

// ------------------------------------------------------------------ generic kernel
enum ConvMode { kConvS1 = 0, kConvS2 = 1, kDeconvS2 = 2 };

// weights in shared memory as [ci][tap][co_t]; x is [n,Cin,Di,Hi,Wi]; y is [n,Cout,Do,Ho,Wo]
template <int MODE, int CO_T>
__global__ void __launch_bounds__(128)
conv3d_generic_kernel(const float* __restrict__ x, int Cin, int Di, int Hi, int Wi, const float* __restrict__ wgt,
                      const float* __restrict__ scale, const float* __restrict__ shift, int Cout, int Do, int Ho,
                      int Wo, const float* __restrict__ skip, float* __restrict__ y, long long n_vox_total) {
    extern __shared__ __align__(16) float s_w[];
    const int cog = blockIdx.y;
    for (int i = threadIdx.x; i < Cin * 27 * CO_T; i += blockDim.x) {
        int co = i % CO_T, tap = (i / CO_T) % 27, ci = i / (CO_T * 27);
        size_t src = (MODE == kDeconvS2) ? ((size_t)ci * Cout + cog * CO_T + co) * 27 + tap
                                         : ((size_t)(cog * CO_T + co) * Cin + ci) * 27 + tap;
        s_w[i] = __ldg(wgt + src);
    }
    __syncthreads();

    const size_t ivol = (size_t)Di * Hi * Wi, iplane = (size_t)Hi * Wi;
    const size_t ovol = (size_t)Do * Ho * Wo;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_vox_total;
         v += (long long)gridDim.x * blockDim.x) {
        int ox = (int)(v % Wo);
        long long r = v / Wo;
        int oy = (int)(r % Ho);
        r /= Ho;
        int oz = (int)(r % Do);
        int n = (int)(r / Do);
        const float* xn = x + (size_t)n * Cin * ivol;

        // per-dimension tap -> input index (or -1)
        int iz[3], iy[3], ixx[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (MODE == kDeconvS2) {
                int tz = oz + 1 - k, ty = oy + 1 - k, tx = ox + 1 - k;  // o = 2 i - 1 + k
                iz[k] = (tz >= 0 && !(tz & 1) && (tz >> 1) < Di) ? (tz >> 1) : -1;
                iy[k] = (ty >= 0 && !(ty & 1) && (ty >> 1) < Hi) ? (ty >> 1) : -1;
                ixx[k] = (tx >= 0 && !(tx & 1) && (tx >> 1) < Wi) ? (tx >> 1) : -1;
            } else {
                const int s = (MODE == kConvS2) ? 2 : 1;
                int tz = oz * s + k - 1, ty = oy * s + k - 1, tx = ox * s + k - 1;
                iz[k] = (tz >= 0 && tz < Di) ? tz : -1;
                iy[k] = (ty >= 0 && ty < Hi) ? ty : -1;
                ixx[k] = (tx >= 0 && tx < Wi) ? tx : -1;
            }
        }
        float acc[CO_T];
#pragma unroll
        for (int c = 0; c < CO_T; ++c) acc[c] = 0.f;
        for (int ci = 0; ci < Cin; ++ci) {
            const float* xc = xn + (size_t)ci * ivol;
            const float* wc = s_w + ci * 27 * CO_T;
#pragma unroll
            for (int kd = 0; kd < 3; ++kd) {
                if (iz[kd] < 0) continue;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    if (iy[kh] < 0) continue;
                    const float* xr = xc + (size_t)iz[kd] * iplane + (size_t)iy[kh] * Wi;
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        if (ixx[kw] < 0) continue;
                        float in = __ldg(xr + ixx[kw]);
                        const float* wp = wc + ((kd * 3 + kh) * 3 + kw) * CO_T;
#pragma unroll
                        for (int c4 = 0; c4 < CO_T; c4 += 4) {
                            float4 w = *reinterpret_cast<const float4*>(wp + c4);
                            acc[c4] = fmaf(in, w.x, acc[c4]);
                            acc[c4 + 1] = fmaf(in, w.y, acc[c4 + 1]);
                            acc[c4 + 2] = fmaf(in, w.z, acc[c4 + 2]);
                            acc[c4 + 3] = fmaf(in, w.w, acc[c4 + 3]);
                        }
                    }
                }
            }
        }
        const size_t o = ((size_t)n * Cout + cog * CO_T) * ovol + ((size_t)oz * Ho + oy) * Wo + ox;
#pragma unroll
        for (int c = 0; c < CO_T; ++c) {
            const int co = cog * CO_T + c;
            float val = fmaxf(fmaf(acc[c], __ldg(scale + co), __ldg(shift + co)), 0.f);
            if (skip) val += __ldg(skip + o + (size_t)c * ovol);
            y[o + (size_t)c * ovol] = val;
        }
    }
}

// ------------------------------------------------------------------ prob conv + soft-argmin
// CTA = 16 consecutive pixels of one image row x 16 depth lanes.  A thread convolves the planes
// d = lane, lane + 16, ... of its pixel (8 -> 1 channels, 3x3x3, bias) and folds them into an
// online softmax(-x) state with the plane depth as value; the 16 states of a pixel are merged
// through shared memory.  The regularised volume itself is only written on request.
constexpr int PS_TX = 16, PS_DL = 16;

__global__ void __launch_bounds__(PS_TX * PS_DL)
prob_softargmin_kernel(const float* __restrict__ x, int Cin, int D, int H, int W, const float* __restrict__ wgt,
                       float bias, float d_start, float d_end, float* __restrict__ x_reg, float* __restrict__ depth) {
    extern __shared__ float s_w[];  // [Cin][27]
    __shared__ float s_m[PS_DL][PS_TX], s_s[PS_DL][PS_TX], s_t[PS_DL][PS_TX];
    for (int i = threadIdx.x; i < Cin * 27; i += blockDim.x) s_w[i] = __ldg(wgt + i);
    __syncthreads();
    const int px = threadIdx.x % PS_TX, dl = threadIdx.x / PS_TX;
    const int ox = blockIdx.x * PS_TX + px, oy = blockIdx.y, n = blockIdx.z;
    const size_t plane = (size_t)H * W, vol = plane * D;
    const float* xn = x + (size_t)n * Cin * vol;
    const bool inside = ox < W;

    float m = -INFINITY, s = 0.f, t = 0.f;
    if (inside) {
        for (int d = dl; d < D; d += PS_DL) {
            float acc = bias;
            for (int ci = 0; ci < Cin; ++ci) {
                const float* xc = xn + (size_t)ci * vol;
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) {
                    const int z = d + kd - 1;
                    if (z < 0 || z >= D) continue;
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const int yy = oy + kh - 1;
                        if (yy < 0 || yy >= H) continue;
                        const float* xr = xc + (size_t)z * plane + (size_t)yy * W;
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            const int xx = ox + kw - 1;
                            if (xx < 0 || xx >= W) continue;
                            acc = fmaf(__ldg(xr + xx), s_w[ci * 27 + (kd * 3 + kh) * 3 + kw], acc);
                        }
                    }
                }
            }
            if (x_reg) x_reg[((size_t)n * D + d) * plane + (size_t)oy * W + ox] = acc;
            const float v = -acc;
            const float mn = fmaxf(m, v);
            const float corr = expf(m - mn), e = expf(v - mn);
            s = s * corr + e;
            t = t * corr + e * linspace_torch(d_start, d_end, D, d);
            m = mn;
        }
    }
    s_m[dl][px] = m;
    s_s[dl][px] = s;
    s_t[dl][px] = t;
    __syncthreads();
    if (dl == 0 && inside) {
        float M = -INFINITY;
#pragma unroll
        for (int i = 0; i < PS_DL; ++i) M = fmaxf(M, s_m[i][px]);
        float S = 0.f, T = 0.f;
#pragma unroll
        for (int i = 0; i < PS_DL; ++i) {
            const float c = expf(s_m[i][px] - M);  // lanes that saw no plane carry m = -inf -> 0
            S = fmaf(s_s[i][px], c, S);
            T = fmaf(s_t[i][px], c, T);
        }
        depth[((size_t)n * H + oy) * W + ox] = T / S;
    }
}

static int fold_check(const float* scale, const float* shift) { return scale && shift; }

}  // namespace dv3d

using namespace dv3d;

template <int MODE, int CO_T>
static int launch_generic(const float* x, int n, int Cin, int Di, int Hi, int Wi, const float* w, const float* scale,
                          const float* shift, int Cout, int Do, int Ho, int Wo, const float* skip, float* y,
                          cudaStream_t st) {
    const size_t smem = sizeof(float) * Cin * 27 * CO_T;
    DV3D_CUDA(cudaFuncSetAttribute(conv3d_generic_kernel<MODE, CO_T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    long long total = (long long)n * Do * Ho * Wo;
    int blocks = cdiv(total, 128);
    int cap = kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    dim3 grid(blocks, Cout / CO_T);
    conv3d_generic_kernel<MODE, CO_T><<<grid, 128, smem, st>>>(x, Cin, Di, Hi, Wi, w, scale, shift, Cout, Do, Ho, Wo,
                                                               skip, y, total);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_conv3d_bn_relu(const float* x, int n, int Cin, int D, int H, int W, const float* weight,
                                   const float* scale, const float* shift, int Cout, int stride, const float* skip,
                                   float* y, void* stream) {
    DV3D_REQUIRE(x && weight && y && fold_check(scale, shift), "conv3d: null pointer");
    DV3D_REQUIRE(n >= 0 && Cin > 0 && Cout > 0 && Cout % 8 == 0 && D > 0 && H > 0 && W > 0, "conv3d: bad shape");
    DV3D_REQUIRE(stride == 1 || stride == 2, "conv3d: stride must be 1 or 2");
    DV3D_REQUIRE(Cin <= 128, "conv3d: Cin > 128 unsupported");
    if (n == 0) return DV3D_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (stride == 1 && Cin % CIC == 0 && (long long)D * H * W >= 16384) {
        static bool attr = false;
        if (!attr) {
            DV3D_CUDA(cudaFuncSetAttribute(conv3d_s1_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)S1_SMEM));
            attr = true;
        }
        int tiles_x = cdiv(W, TX), tiles_y = cdiv(H, TY);
        dim3 grid(tiles_x * tiles_y, cdiv(D, TZ), n * (Cout / COT));
        DV3D_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv3d: grid too large");
        conv3d_s1_tiled_kernel<<<grid, S1_THREADS, S1_SMEM, st>>>(x, Cin, D, H, W, weight, scale, shift, Cout, skip, y,
                                                                  tiles_x, tiles_y);
        DV3D_LAUNCHED();
        return DV3D_OK;
    }
    if (stride == 1) {
        if (Cout % 16 == 0)
            return launch_generic<kConvS1, 16>(x, n, Cin, D, H, W, weight, scale, shift, Cout, D, H, W, skip, y, st);
        return launch_generic<kConvS1, 8>(x, n, Cin, D, H, W, weight, scale, shift, Cout, D, H, W, skip, y, st);
    }
    const int Do = (D + 1) / 2, Ho = (H + 1) / 2, Wo = (W + 1) / 2;  // floor((D + 2 - 3)/2) + 1
    if (Cout % 16 == 0)
        return launch_generic<kConvS2, 16>(x, n, Cin, D, H, W, weight, scale, shift, Cout, Do, Ho, Wo, skip, y, st);
    return launch_generic<kConvS2, 8>(x, n, Cin, D, H, W, weight, scale, shift, Cout, Do, Ho, Wo, skip, y, st);
}

extern "C" int dv3d_deconv3d_bn_relu(const float* x, int n, int Cin, int D, int H, int W, const float* weight,
                                     const float* scale, const float* shift, int Cout, const float* skip, float* y,
                                     void* stream) {
    DV3D_REQUIRE(x && weight && y && fold_check(scale, shift), "deconv3d: null pointer");
    DV3D_REQUIRE(n >= 0 && Cin > 0 && Cin <= 128 && Cout > 0 && Cout % 8 == 0 && D > 0 && H > 0 && W > 0,
                 "deconv3d: bad shape");
    if (n == 0) return DV3D_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (Cout % 16 == 0)
        return launch_generic<kDeconvS2, 16>(x, n, Cin, D, H, W, weight, scale, shift, Cout, 2 * D, 2 * H, 2 * W, skip,
                                             y, st);
    return launch_generic<kDeconvS2, 8>(x, n, Cin, D, H, W, weight, scale, shift, Cout, 2 * D, 2 * H, 2 * W, skip, y,
                                        st);
}

extern "C" int dv3d_prob_softargmin(const float* x, int n, int Cin, int D, int H, int W, const float* weight,
                                    float bias, float depth_start, float depth_end, float* x_reg_out,
                                    float* depth_out, void* stream) {
    DV3D_REQUIRE(x && weight && depth_out, "prob_softargmin: null pointer");
    DV3D_REQUIRE(n >= 0 && Cin > 0 && Cin <= 64 && D > 0 && H > 0 && W > 0, "prob_softargmin: bad shape");
    if (n == 0) return DV3D_OK;
    DV3D_REQUIRE(H <= 65535 && n <= 65535, "prob_softargmin: H or n > 65535");
    dim3 grid(cdiv(W, PS_TX), H, n);
    prob_softargmin_kernel<<<grid, PS_TX * PS_DL, sizeof(float) * Cin * 27, (cudaStream_t)stream>>>(
        x, Cin, D, H, W, weight, bias, depth_start, depth_end, x_reg_out, depth_out);
    DV3D_LAUNCHED();
    return DV3D_OK;
}
