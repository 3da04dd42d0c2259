// CostRegNet layers (inference): 3x3x3 convolutions / transposed convolutions with folded
// BatchNorm + ReLU (+ skip), and the fused prob-conv + softmax(-x) + depth expectation.
// Replaces the cuDNN Conv3d/ConvTranspose3d + BatchNorm3d + ReLU + softmax chain of
// mvsnet.py:18-36,133-163,219-227.  NCDHW fp32, fp32 accumulation on the CUDA cores.
//
// Two kernels:
//  * conv3d_s1_tiled: stride-1 layers on large volumes (conv0 = 68 % of the FLOPs).
//    CTA tile TZ x14x28 outputs x 8 output channels; the haloed input tile of 8 input channels
//    and the matching weights are staged in shared memory; each thread owns a 4(x) x 2(y)
//    x 8(co) register tile = 64 accumulators, 576 FMAs per 26 shared loads.  TZ is chosen
//    per launch so that the CTAs fill the 148 SMs (2 CTAs each) in whole rounds.
//  * conv3d_direct: every other layer (stride-2, transposed, small volumes): thread = output
//    voxel x CO_T output channels x a slice of the input channels (K split, reduced through
//    shared memory), weights of the channel group in shared memory.
#include "common.cuh"

namespace dv3d {

// ------------------------------------------------------------------ tiled stride-1 conv
constexpr int TY = 14, TX = 28;       // output tile in y, x (TZ is a template parameter)
constexpr int IY = TY + 2, IXP = 32;  // haloed input tile, row pitch padded 30 -> 32
constexpr int CIC = 4;                // input channels staged per pass (two passes are resident)
constexpr int COT = 8;                // output channels per CTA
__host__ __device__ constexpr int s1_threads(int TZ) { return (TX / 4) * (TY / 2) * TZ; }  // 49 per plane
__host__ __device__ constexpr size_t s1_smem(int TZ) {
    return 2 * sizeof(float) * (CIC * (TZ + 2) * IY * IXP + CIC * 27 * COT);  // double buffered
}

// 4-byte asynchronous copy global -> shared (LDGSTS); src_bytes = 0 writes a zero instead
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)),
                 "l"(src), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int TZ>
__global__ void __launch_bounds__(s1_threads(TZ), 2)
conv3d_s1_tiled_kernel(const float* __restrict__ x, int Cin, int D, int H, int W, const float* __restrict__ wgt,
                       const float* __restrict__ scale, const float* __restrict__ shift, int Cout,
                       const float* __restrict__ skip, float* __restrict__ y, int tiles_x, int tiles_y) {
    pdl_wait();
    constexpr int IZ = TZ + 2;
    constexpr int S1_THREADS = s1_threads(TZ);
    constexpr int IN_F = CIC * IZ * IY * IXP, W_F = CIC * 27 * COT;  // floats per buffer
    extern __shared__ __align__(16) float smem[];
    // [2] x { input tile [CIC][IZ][IY][IXP] | weights [CIC][27][COT] }

    const int tid = threadIdx.x;
    const int tx = tid % (TX / 4), ty = (tid / (TX / 4)) % (TY / 2), tz = tid / ((TX / 4) * (TY / 2));
    const int bx = blockIdx.x % tiles_x, by = blockIdx.x / tiles_x;
    const int x0 = bx * TX, y0 = by * TY, z0 = blockIdx.y * TZ;
    const int cog = blockIdx.z % (Cout / COT), n = blockIdx.z / (Cout / COT);
    const size_t plane = (size_t)H * W, vol = plane * D;
    const float* xn = x + (size_t)n * Cin * vol;

    // Staging: every thread owns a few fixed (y, x) positions of the haloed plane and copies
    // them for every (channel, z) of a pass with 4-byte cp.async (zero fill outside the volume),
    // so that all the copies of a pass are in flight at once and the next pass loads while the
    // current one is being computed.
    constexpr int PLANE_E = IY * (TX + 2);
    constexpr int SLOTS = (PLANE_E + S1_THREADS - 1) / S1_THREADS;
    int soff[SLOTS], goff[SLOTS];
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
        const int e = tid + k * S1_THREADS;
        const int iy = e / (TX + 2), ix = e - iy * (TX + 2);
        const int gy = y0 + iy - 1, gx = x0 + ix - 1;
        soff[k] = e < PLANE_E ? iy * IXP + ix : -1;
        goff[k] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? gy * W + gx : -1;
    }
    auto stage = [&](int c0, int b) {
        float* s_in = smem + b * (IN_F + W_F);
        float* s_w = s_in + IN_F;
        for (int ci = 0; ci < CIC; ++ci) {
#pragma unroll
            for (int iz = 0; iz < IZ; ++iz) {
                const int gz = z0 + iz - 1;
                const bool zok = gz >= 0 && gz < D;
                const float* src = xn + (size_t)(c0 + ci) * vol + (size_t)(zok ? gz : 0) * plane;
                float* dst = s_in + (ci * IZ + iz) * IY * IXP;
#pragma unroll
                for (int k = 0; k < SLOTS; ++k) {
                    if (soff[k] >= 0) {
                        const bool ok = zok && goff[k] >= 0;
                        cp_async4(dst + soff[k], ok ? src + goff[k] : xn, ok ? 4 : 0);
                    }
                }
            }
        }
        // weights of this pass: s_w[ci][tap][co] = wgt[cog*8+co][c0+ci][tap] (CIC*27 contiguous floats per co)
        for (int i = tid; i < W_F; i += S1_THREADS) {
            const int co = i / (CIC * 27), r = i - co * (CIC * 27);
            cp_async4(s_w + r * COT + co, wgt + ((size_t)(cog * COT + co) * Cin + c0) * 27 + r, 4);
        }
        cp_async_commit();
    };

    // accumulators as channel pairs: FFMA2 (packed fp32, scalar-broadcast input operand) does
    // two output channels per issue slot
    float2 acc[2][4][COT / 2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int c = 0; c < COT / 2; ++c) acc[a][b][c] = make_float2(0.f, 0.f);

    stage(0, 0);
    for (int c0 = 0, pass = 0; c0 < Cin; c0 += CIC, ++pass) {
        if (c0 + CIC < Cin) {
            stage(c0 + CIC, (pass + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();  // this pass's tile has landed for every thread
        const float* s_in = smem + (pass & 1) * (IN_F + W_F);
        const float* s_w = s_in + IN_F;

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
                        const float2 wv[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w),
                                              make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
#pragma unroll
                        for (int a = 0; a < 2; ++a)
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                const float v = in[a + kh][b + kw];
                                const float2 vv = make_float2(v, v);
#pragma unroll
                                for (int c = 0; c < COT / 2; ++c) acc[a][b][c] = __ffma2_rn(vv, wv[c], acc[a][b][c]);
                            }
                    }
                }
            }
        }
        __syncthreads();  // pass consumed: its buffer is restaged two passes later
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
                    const float2 pr = acc[a][b][c >> 1];
                    float v = fmaxf(fmaf((c & 1) ? pr.y : pr.x, sc, sh), 0.f);
                    if (sn) v += __ldg(sn + base + b);
                    yn[base + b] = v;
                }
            }
        }
    }
}

// ------------------------------------------------------------------ direct kernels, optional K split
enum ConvMode { kConvS1 = 0, kConvS2 = 1, kDeconvS2 = 2 };

constexpr int DC_THREADS = 256;

// weights of one output-channel group into shared memory as [ci][tap][CO_T], by 4-byte
// cp.async in global order (all copies in flight at once)
template <bool TRANSPOSED, int CO_T>
__device__ __forceinline__ void stage_group_weights(float* s_w, const float* __restrict__ wgt, int Cin, int Cout, int cog) {
    for (int i = threadIdx.x; i < Cin * 27 * CO_T; i += DC_THREADS) {
        const int tap = i % 27;
        int ci, co;
        size_t src;
        if (TRANSPOSED) {  // ConvTranspose3d weight [Cin][Cout][27]
            co = (i / 27) % CO_T;
            ci = i / (27 * CO_T);
            src = ((size_t)ci * Cout + cog * CO_T + co) * 27 + tap;
        } else {  // Conv3d weight [Cout][Cin][27]
            ci = (i / 27) % Cin;
            co = i / (27 * Cin);
            src = ((size_t)(cog * CO_T + co) * Cin + ci) * 27 + tap;
        }
        cp_async4(s_w + (ci * 27 + tap) * CO_T + co, wgt + src, 4);
    }
    cp_async_commit();
}

// Convolution, stride 1 or 2.  thread = (output voxel v, input-channel slice s); CTA =
// (256 / n_slices) voxels x n_slices.  Per input channel the 27 taps of the voxel are fetched
// by 27 independent predicated loads (validity mask and offsets are per-thread constants), then
// contracted with the group's weights by FFMA2; the partial sums of the slices are added in
// slice order through shared memory.  The K split gives the small levels of the U-Net
// (588 .. 4704 voxels) enough threads to fill 148 SMs.
// x is [n,Cin,Di,Hi,Wi]; y is [n,Cout,Do,Ho,Wo]
template <int STRIDE, int CO_T>
__global__ void __launch_bounds__(DC_THREADS)
conv3d_direct_kernel(const float* __restrict__ x, int Cin, int Di, int Hi, int Wi, const float* __restrict__ wgt,
                     const float* __restrict__ scale, const float* __restrict__ shift, int Cout, int Do, int Ho,
                     int Wo, const float* __restrict__ skip, float* __restrict__ y, long long n_vox_total,
                     int n_slices) {
    pdl_wait();
    extern __shared__ __align__(16) float smem[];
    float* s_w = smem;                      // [Cin][27][CO_T]
    float* s_red = smem + Cin * 27 * CO_T;  // [n_slices][CO_T][vox]
    const int tid = threadIdx.x;
    const int cog = blockIdx.y;
    stage_group_weights<false, CO_T>(s_w, wgt, Cin, Cout, cog);

    const int vox = DC_THREADS / n_slices;
    const int v = tid % vox, s = tid / vox;
    const int cps = Cin / n_slices;
    const size_t ivol = (size_t)Di * Hi * Wi;
    const int iplane = Hi * Wi;
    const size_t ovol = (size_t)Do * Ho * Wo;
    const long long gv = (long long)blockIdx.x * vox + v;
    const bool live = gv < n_vox_total;

    // per-thread tap table: validity bits per dimension and the centre offset
    int n = 0, centre = 0;
    unsigned zm = 0, ym = 0, xm = 0;
    if (live) {
        n = (int)(gv / (long long)ovol);
        const int sp = (int)(gv - (long long)n * ovol);
        const int ox = sp % Wo, oy = (sp / Wo) % Ho, oz = sp / (Wo * Ho);
        const int cz = oz * STRIDE, cy = oy * STRIDE, cx = ox * STRIDE;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            zm |= (unsigned)(cz + k - 1 >= 0 && cz + k - 1 < Di) << k;
            ym |= (unsigned)(cy + k - 1 >= 0 && cy + k - 1 < Hi) << k;
            xm |= (unsigned)(cx + k - 1 >= 0 && cx + k - 1 < Wi) << k;
        }
        centre = (cz * Hi + cy) * Wi + cx;
    }
    const float* xs = x + ((size_t)n * Cin + (size_t)s * cps) * ivol + centre;

    float2 acc2[CO_T / 2];  // channel pairs for FFMA2
#pragma unroll
    for (int c = 0; c < CO_T / 2; ++c) acc2[c] = make_float2(0.f, 0.f);
    cp_async_wait<0>();
    __syncthreads();
    const float* ws = s_w + (size_t)s * cps * 27 * CO_T;
    for (int ci = 0; ci < cps; ++ci) {
        const float* xc = xs + (size_t)ci * ivol;
        float in[27];
#pragma unroll
        for (int kd = 0; kd < 3; ++kd)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const bool ok = ((zm >> kd) & (ym >> kh) & (xm >> kw) & 1u) != 0;
                    in[(kd * 3 + kh) * 3 + kw] = ok ? __ldg(xc + (kd - 1) * iplane + (kh - 1) * Wi + (kw - 1)) : 0.f;
                }
        const float* wc = ws + ci * 27 * CO_T;
#pragma unroll
        for (int t = 0; t < 27; ++t) {
            const float2 in2 = make_float2(in[t], in[t]);
#pragma unroll
            for (int c4 = 0; c4 < CO_T; c4 += 4) {
                const float4 w = *reinterpret_cast<const float4*>(wc + t * CO_T + c4);
                acc2[c4 / 2] = __ffma2_rn(in2, make_float2(w.x, w.y), acc2[c4 / 2]);
                acc2[c4 / 2 + 1] = __ffma2_rn(in2, make_float2(w.z, w.w), acc2[c4 / 2 + 1]);
            }
        }
    }
    float acc[CO_T];
#pragma unroll
    for (int c = 0; c < CO_T / 2; ++c) acc[2 * c] = acc2[c].x, acc[2 * c + 1] = acc2[c].y;
    if (n_slices == 1) {
        if (!live) return;
        const size_t o = ((size_t)n * Cout + cog * CO_T) * ovol + (size_t)(gv - (long long)n * ovol);
#pragma unroll
        for (int c = 0; c < CO_T; ++c) {
            const int co = cog * CO_T + c;
            float val = fmaxf(fmaf(acc[c], __ldg(scale + co), __ldg(shift + co)), 0.f);
            if (skip) val += __ldg(skip + o + (size_t)c * ovol);
            y[o + (size_t)c * ovol] = val;
        }
        return;
    }
#pragma unroll
    for (int c = 0; c < CO_T; ++c) s_red[(s * CO_T + c) * vox + v] = acc[c];
    __syncthreads();
    for (int o = tid; o < vox * CO_T; o += DC_THREADS) {
        const int vv = o % vox, c = o / vox;
        const long long g = (long long)blockIdx.x * vox + vv;
        if (g >= n_vox_total) continue;
        float sum = 0.f;
        for (int ss = 0; ss < n_slices; ++ss) sum += s_red[(ss * CO_T + c) * vox + vv];
        const int nn = (int)(g / (long long)ovol);
        const int co = cog * CO_T + c;
        const size_t idx = ((size_t)nn * Cout + co) * ovol + (size_t)(g - (long long)nn * ovol);
        float val = fmaxf(fmaf(sum, __ldg(scale + co), __ldg(shift + co)), 0.f);
        if (skip) val += __ldg(skip + idx);
        y[idx] = val;
    }
}

// Transposed convolution, stride 2, padding 1, output_padding 1 (output = 2 x input).
// thread = (INPUT voxel (i,j,k), input-channel slice): it produces the 2x2x2 output block
// (2i+a, 2j+b, 2k+c) from the 2x2x2 input neighbourhood (i+dz, j+dy, k+dx).  Along one axis
// output parity a and input step d pair with exactly one tap: (a=0,d=0) -> k=1, (a=1,d=0) ->
// k=2, (a=1,d=1) -> k=0 (from o = 2 i' - 1 + k), so the 27 taps are each used exactly once per
// block and there is no parity divergence and no wasted tap test.
__device__ __forceinline__ constexpr int deconv_tap(int a, int d) { return a == 0 ? (d == 0 ? 1 : -1) : (d == 0 ? 2 : 0); }

template <int CO_T>
__global__ void __launch_bounds__(DC_THREADS, 2)
deconv3d_block_kernel(const float* __restrict__ x, int Cin, int Di, int Hi, int Wi, const float* __restrict__ wgt,
                      const float* __restrict__ scale, const float* __restrict__ shift, int Cout,
                      const float* __restrict__ skip, float* __restrict__ y, long long n_vox_total, int n_slices) {
    pdl_wait();
    extern __shared__ __align__(16) float smem[];
    float* s_w = smem;                      // [Cin][27][CO_T]
    float* s_red = smem + Cin * 27 * CO_T;  // [n_slices][CO_T][vox] per output parity
    const int tid = threadIdx.x;
    const int cog = blockIdx.y;
    stage_group_weights<true, CO_T>(s_w, wgt, Cin, Cout, cog);

    const int vox = DC_THREADS / n_slices;
    const int v = tid % vox, s = tid / vox;
    const int cps = Cin / n_slices;
    const size_t ivol = (size_t)Di * Hi * Wi;
    const int iplane = Hi * Wi;
    const int Do = 2 * Di, Ho = 2 * Hi, Wo = 2 * Wi;
    const size_t ovol = 8 * ivol;
    const long long gv = (long long)blockIdx.x * vox + v;  // input voxel
    const bool live = gv < n_vox_total;
    int n = 0, iz = 0, iy = 0, ix = 0;
    if (live) {
        n = (int)(gv / (long long)ivol);
        const int sp = (int)(gv - (long long)n * ivol);
        ix = sp % Wi, iy = (sp / Wi) % Hi, iz = sp / iplane;
    }
    const bool z1 = live && iz + 1 < Di, y1 = iy + 1 < Hi, x1 = ix + 1 < Wi;
    const float* xs = x + ((size_t)n * Cin + (size_t)s * cps) * ivol + (size_t)iz * iplane + iy * Wi + ix;

    float2 acc[8][CO_T / 2];  // [output parity a*4+b*2+c][channel pair]
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int c = 0; c < CO_T / 2; ++c) acc[q][c] = make_float2(0.f, 0.f);
    cp_async_wait<0>();
    __syncthreads();
    const float* ws = s_w + (size_t)s * cps * 27 * CO_T;
    for (int ci = 0; ci < cps; ++ci) {
        const float* xc = xs + (size_t)ci * ivol;
        float in[8];  // [dz*4 + dy*2 + dx]
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int dz = q >> 2, dy = (q >> 1) & 1, dx = q & 1;
            const bool ok = live && (!dz || z1) && (!dy || y1) && (!dx || x1);
            in[q] = ok ? __ldg(xc + dz * iplane + dy * Wi + dx) : 0.f;
        }
        const float* wc = ws + ci * 27 * CO_T;
#pragma unroll
        for (int q = 0; q < 8; ++q) {        // output parity
#pragma unroll
            for (int e = 0; e < 8; ++e) {    // input neighbour
                const int kz = deconv_tap(q >> 2, e >> 2), ky = deconv_tap((q >> 1) & 1, (e >> 1) & 1),
                          kx = deconv_tap(q & 1, e & 1);
                if (kz < 0 || ky < 0 || kx < 0) continue;  // compile-time
                const float2 in2 = make_float2(in[e], in[e]);
                const float* wt = wc + ((kz * 3 + ky) * 3 + kx) * CO_T;
#pragma unroll
                for (int c4 = 0; c4 < CO_T; c4 += 4) {
                    const float4 w = *reinterpret_cast<const float4*>(wt + c4);
                    acc[q][c4 / 2] = __ffma2_rn(in2, make_float2(w.x, w.y), acc[q][c4 / 2]);
                    acc[q][c4 / 2 + 1] = __ffma2_rn(in2, make_float2(w.z, w.w), acc[q][c4 / 2 + 1]);
                }
            }
        }
    }
    // epilogue, one output parity at a time (keeps the reduction buffer at 256 * CO_T floats)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int a = q >> 2, b = (q >> 1) & 1, c1 = q & 1;
        if (n_slices == 1) {
            if (live) {
                const size_t o = ((size_t)n * Cout + cog * CO_T) * ovol + ((size_t)(2 * iz + a) * Ho + 2 * iy + b) * Wo +
                                 2 * ix + c1;
#pragma unroll
                for (int c = 0; c < CO_T; ++c) {
                    const int co = cog * CO_T + c;
                    const float r = (c & 1) ? acc[q][c >> 1].y : acc[q][c >> 1].x;
                    float val = fmaxf(fmaf(r, __ldg(scale + co), __ldg(shift + co)), 0.f);
                    if (skip) val += __ldg(skip + o + (size_t)c * ovol);
                    y[o + (size_t)c * ovol] = val;
                }
            }
            continue;
        }
        if (q) __syncthreads();  // previous parity's sums have been read
#pragma unroll
        for (int c = 0; c < CO_T; ++c) s_red[(s * CO_T + c) * vox + v] = (c & 1) ? acc[q][c >> 1].y : acc[q][c >> 1].x;
        __syncthreads();
        for (int o = tid; o < vox * CO_T; o += DC_THREADS) {
            const int vv = o % vox, c = o / vox;
            const long long g = (long long)blockIdx.x * vox + vv;
            if (g >= n_vox_total) continue;
            float sum = 0.f;
            for (int ss = 0; ss < n_slices; ++ss) sum += s_red[(ss * CO_T + c) * vox + vv];
            const int nn = (int)(g / (long long)ivol);
            const int sp = (int)(g - (long long)nn * ivol);
            const int jx = sp % Wi, jy = (sp / Wi) % Hi, jz = sp / iplane;
            const int co = cog * CO_T + c;
            const size_t idx = ((size_t)nn * Cout + co) * ovol + ((size_t)(2 * jz + a) * Ho + 2 * jy + b) * Wo + 2 * jx + c1;
            float val = fmaxf(fmaf(sum, __ldg(scale + co), __ldg(shift + co)), 0.f);
            if (skip) val += __ldg(skip + idx);
            y[idx] = val;
        }
    }
}

// ------------------------------------------------------------------ prob conv + soft-argmin
// Thread = (pixel, depth segment of 8 planes); CTA = 8 consecutive pixels of one image row x all
// the segments of their depth column.  A thread walks the 10 input planes of its segment once:
// the 3x3xCin neighbourhood of a plane (72 loads at Cin = 8) feeds the three outputs it touches
// (kd = 0,1,2), so every loaded value is used three times instead of once.  The 8 logits of the
// segment are folded into an online softmax(-x) state with the plane depth as value and the
// states of a pixel are merged through shared memory.  The regularised volume itself is only
// written on request.
constexpr int PS_TX = 8, PS_SEG = 8, PS_MAXSEG = 32;

template <int CIN>
__global__ void __launch_bounds__(PS_TX * PS_MAXSEG)
prob_softargmin_kernel(const float* __restrict__ x, int D, int H, int W, const float* __restrict__ wgt, float bias,
                       float d_start, float d_end, float* __restrict__ x_reg, float* __restrict__ depth) {
    pdl_wait();
    __shared__ __align__(16) float s_w[27 * CIN];  // [kd][kh][kw][ci]
    __shared__ float s_m[PS_MAXSEG][PS_TX], s_s[PS_MAXSEG][PS_TX], s_t[PS_MAXSEG][PS_TX];
    const int tid = threadIdx.y * PS_TX + threadIdx.x;
    for (int i = tid; i < CIN * 27; i += PS_TX * blockDim.y) s_w[(i % 27) * CIN + i / 27] = __ldg(wgt + i);
    __syncthreads();
    const int px = threadIdx.x, sg = threadIdx.y;
    const int ox = blockIdx.x * PS_TX + px, oy = blockIdx.y, n = blockIdx.z;
    const size_t plane = (size_t)H * W, vol = plane * D;
    const float* xn = x + (size_t)n * CIN * vol;
    const bool inside = ox < W;
    const int n_seg = (D + PS_SEG - 1) / PS_SEG;

    float m = -INFINITY, s = 0.f, t = 0.f;
    for (int seg = sg; seg < n_seg && inside; seg += blockDim.y) {
        const int d0 = seg * PS_SEG;
        float acc[PS_SEG];
#pragma unroll
        for (int j = 0; j < PS_SEG; ++j) acc[j] = 0.f;
#pragma unroll
        for (int zi = 0; zi < PS_SEG + 2; ++zi) {  // input plane z = d0 - 1 + zi
            const int z = d0 - 1 + zi;
            if (z < 0 || z >= D) continue;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int yy = oy + kh - 1;
                if (yy < 0 || yy >= H) continue;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int xx = ox + kw - 1;
                    if (xx < 0 || xx >= W) continue;
                    const float* xp = xn + (size_t)z * plane + (size_t)yy * W + xx;
                    float v[CIN];
#pragma unroll
                    for (int ci = 0; ci < CIN; ++ci) v[ci] = __ldg(xp + (size_t)ci * vol);
#pragma unroll
                    for (int kd = 0; kd < 3; ++kd) {  // output plane d = z - kd + 1, j = d - d0 = zi - kd
                        const int j = zi - kd;
                        if (j < 0 || j >= PS_SEG) continue;  // compile-time
                        const float* wp = s_w + ((kd * 3 + kh) * 3 + kw) * CIN;
#pragma unroll
                        for (int ci = 0; ci < CIN; ++ci) acc[j] = fmaf(v[ci], wp[ci], acc[j]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < PS_SEG; ++j) {
            const int d = d0 + j;
            if (d >= D) break;
            const float a = bias + acc[j];
            if (x_reg) x_reg[((size_t)n * D + d) * plane + (size_t)oy * W + ox] = a;
            const float v = -a;
            const float mn = fmaxf(m, v);
            const float corr = expf(m - mn), e = expf(v - mn);
            s = s * corr + e;
            t = t * corr + e * linspace_torch(d_start, d_end, D, d);
            m = mn;
        }
    }
    s_m[sg][px] = m;
    s_s[sg][px] = s;
    s_t[sg][px] = t;
    __syncthreads();
    if (sg == 0 && inside) {
        float M = -INFINITY;
        for (int i = 0; i < (int)blockDim.y; ++i) M = fmaxf(M, s_m[i][px]);
        float S = 0.f, T = 0.f;
        for (int i = 0; i < (int)blockDim.y; ++i) {
            const float c = expf(s_m[i][px] - M);  // segments that saw no plane carry m = -inf -> 0
            S = fmaf(s_s[i][px], c, S);
            T = fmaf(s_t[i][px], c, T);
        }
        depth[((size_t)n * H + oy) * W + ox] = T / S;
    }
}

static int fold_check(const float* scale, const float* shift) { return scale && shift; }

}  // namespace dv3d

using namespace dv3d;

// K split: the smallest power of two that gives the 148 SMs `min_ctas` CTAs (or 8)
static int pick_slices(long long threads_voxels, int Cin, int cogs, int min_ctas) {
    int ns = 1;
    while (ns < 8 && Cin % (2 * ns) == 0 && (long long)cdiv(threads_voxels, DC_THREADS / ns) * cogs < min_ctas) ns *= 2;
    return ns;
}

template <int STRIDE, int CO_T>
static int launch_direct(const float* x, int n, int Cin, int Di, int Hi, int Wi, const float* w, const float* scale,
                         const float* shift, int Cout, int Do, int Ho, int Wo, const float* skip, float* y,
                         cudaStream_t st) {
    const long long total = (long long)n * Do * Ho * Wo;
    const int cogs = Cout / CO_T;
    const int ns = pick_slices(total, Cin, cogs, 2 * kNumSMs);
    const size_t smem = sizeof(float) * ((size_t)Cin * 27 * CO_T + (ns > 1 ? DC_THREADS * CO_T : 0));
    DV3D_REQUIRE(smem <= 200 * 1024, "conv3d: weights of one channel group (%zu bytes) do not fit shared memory", smem);
    static std::atomic<unsigned long long> attr{0};  // per instantiation; the opt-in covers the 200 KB bound above
    DV3D_FUNC_SMEM_ONCE(attr, (conv3d_direct_kernel<STRIDE, CO_T>), 200 * 1024);
    dim3 grid(cdiv(total, DC_THREADS / ns), cogs);
    DV3D_REQUIRE(grid.y <= 65535, "conv3d: too many channel groups");
    DV3D_LAUNCH((conv3d_direct_kernel<STRIDE, CO_T>), grid, DC_THREADS, smem, st, x, Cin, Di, Hi, Wi, w, scale, shift, Cout, Do, Ho, Wo, skip, y, total, ns);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

// 16 output channels per thread halve the input loads per FMA; the weights of the group must
// leave room for several CTAs per SM
template <int STRIDE>
static int launch_direct_any(const float* x, int n, int Cin, int Di, int Hi, int Wi, const float* w, const float* scale,
                             const float* shift, int Cout, int Do, int Ho, int Wo, const float* skip, float* y,
                             cudaStream_t st) {
    if (Cout % 16 == 0 && Cin <= 32)
        return launch_direct<STRIDE, 16>(x, n, Cin, Di, Hi, Wi, w, scale, shift, Cout, Do, Ho, Wo, skip, y, st);
    return launch_direct<STRIDE, 8>(x, n, Cin, Di, Hi, Wi, w, scale, shift, Cout, Do, Ho, Wo, skip, y, st);
}

static int launch_deconv(const float* x, int n, int Cin, int Di, int Hi, int Wi, const float* w, const float* scale,
                         const float* shift, int Cout, const float* skip, float* y, cudaStream_t st) {
    constexpr int CO_T = 8;
    const long long total = (long long)n * Di * Hi * Wi;  // threads are INPUT voxels
    const int cogs = Cout / CO_T;
    // a thread produces 8 output voxels x 8 channels: one CTA per SM is already a full round
    const int ns = pick_slices(total, Cin, cogs, kNumSMs - 8);
    const size_t smem = sizeof(float) * ((size_t)Cin * 27 * CO_T + (ns > 1 ? DC_THREADS * CO_T : 0));
    DV3D_REQUIRE(smem <= 200 * 1024, "deconv3d: weights of one channel group (%zu bytes) do not fit shared memory", smem);
    static std::atomic<unsigned long long> attr{0};
    DV3D_FUNC_SMEM_ONCE(attr, (deconv3d_block_kernel<CO_T>), 200 * 1024);
    dim3 grid(cdiv(total, DC_THREADS / ns), cogs);
    DV3D_REQUIRE(grid.y <= 65535, "deconv3d: too many channel groups");
    DV3D_LAUNCH((deconv3d_block_kernel<CO_T>), grid, DC_THREADS, smem, st, x, Cin, Di, Hi, Wi, w, scale, shift, Cout, skip, y, total, ns);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

template <int TZ>
static int launch_s1_tiled(const float* x, int n, int Cin, int D, int H, int W, const float* weight, const float* scale,
                           const float* shift, int Cout, const float* skip, float* y, cudaStream_t st) {
    static std::atomic<unsigned long long> attr{0};
    DV3D_FUNC_SMEM_ONCE(attr, (conv3d_s1_tiled_kernel<TZ>), (int)s1_smem(TZ));
    const int tiles_x = cdiv(W, TX), tiles_y = cdiv(H, TY);
    dim3 grid(tiles_x * tiles_y, cdiv(D, TZ), n * (Cout / COT));
    DV3D_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv3d: grid too large");
    DV3D_LAUNCH((conv3d_s1_tiled_kernel<TZ>), grid, s1_threads(TZ), s1_smem(TZ), st, x, Cin, D, H, W, weight, scale, shift, Cout, skip, y, tiles_x, tiles_y);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

namespace dv3d {
// csrc/conv3d_tc.cu: the 32 -> 8 first layer on tcgen05
int conv3d_tc_mode();
int launch_conv3d_c32_c8_tc(const float* x, int n, int D, int H, int W, const float* weight, const float* scale,
                            const float* shift, float* y, cudaStream_t st);
}  // namespace dv3d

extern "C" int dv3d_conv3d_bn_relu(const float* x, int n, int Cin, int D, int H, int W, const float* weight,
                                   const float* scale, const float* shift, int Cout, int stride, const float* skip,
                                   float* y, void* stream) {
    DV3D_REQUIRE(x && weight && y && fold_check(scale, shift), "conv3d: null pointer");
    DV3D_REQUIRE(n >= 0 && Cin > 0 && Cout > 0 && Cout % 8 == 0 && D > 0 && H > 0 && W > 0, "conv3d: bad shape");
    DV3D_REQUIRE(stride == 1 || stride == 2, "conv3d: stride must be 1 or 2");
    DV3D_REQUIRE(Cin <= 128, "conv3d: Cin > 128 unsupported");
    if (n == 0) return DV3D_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (stride == 1 && Cin == 32 && Cout == 8 && !skip && conv3d_tc_mode() == 0)
        return launch_conv3d_c32_c8_tc(x, n, D, H, W, weight, scale, shift, y, st);
    if (stride == 1 && Cin % CIC == 0 && (long long)D * H * W >= 100000) {
        // depth of the CTA tile: fewest rounds of 2 CTAs per SM, weighted by the planes per CTA
        const long long per_plane = (long long)cdiv(W, TX) * cdiv(H, TY) * n * (Cout / COT);
        int best = 4;
        long long best_cost = -1;
        for (int tz = 4; tz >= 2; --tz) {
            const long long ctas = per_plane * cdiv(D, tz);
            const long long cost = ((ctas + 2 * kNumSMs - 1) / (2 * kNumSMs)) * (tz + 1);
            if (best_cost < 0 || cost < best_cost) best = tz, best_cost = cost;
        }
        if (best == 2) return launch_s1_tiled<2>(x, n, Cin, D, H, W, weight, scale, shift, Cout, skip, y, st);
        if (best == 3) return launch_s1_tiled<3>(x, n, Cin, D, H, W, weight, scale, shift, Cout, skip, y, st);
        return launch_s1_tiled<4>(x, n, Cin, D, H, W, weight, scale, shift, Cout, skip, y, st);
    }
    if (stride == 1)
        return launch_direct_any<1>(x, n, Cin, D, H, W, weight, scale, shift, Cout, D, H, W, skip, y, st);
    const int Do = (D + 1) / 2, Ho = (H + 1) / 2, Wo = (W + 1) / 2;  // floor((D + 2 - 3)/2) + 1
    return launch_direct_any<2>(x, n, Cin, D, H, W, weight, scale, shift, Cout, Do, Ho, Wo, skip, y, st);
}

extern "C" int dv3d_deconv3d_bn_relu(const float* x, int n, int Cin, int D, int H, int W, const float* weight,
                                     const float* scale, const float* shift, int Cout, const float* skip, float* y,
                                     void* stream) {
    DV3D_REQUIRE(x && weight && y && fold_check(scale, shift), "deconv3d: null pointer");
    DV3D_REQUIRE(n >= 0 && Cin > 0 && Cin <= 128 && Cout > 0 && Cout % 8 == 0 && D > 0 && H > 0 && W > 0,
                 "deconv3d: bad shape");
    if (n == 0) return DV3D_OK;
    return launch_deconv(x, n, Cin, D, H, W, weight, scale, shift, Cout, skip, y, (cudaStream_t)stream);
}

extern "C" int dv3d_prob_softargmin(const float* x, int n, int Cin, int D, int H, int W, const float* weight,
                                    float bias, float depth_start, float depth_end, float* x_reg_out,
                                    float* depth_out, void* stream) {
    DV3D_REQUIRE(x && weight && depth_out, "prob_softargmin: null pointer");
    DV3D_REQUIRE(n >= 0 && (Cin == 8 || Cin == 16) && D > 0 && H > 0 && W > 0,
                 "prob_softargmin: bad shape (Cin must be 8 or 16, got %d)", Cin);
    if (n == 0) return DV3D_OK;
    DV3D_REQUIRE(H <= 65535 && n <= 65535, "prob_softargmin: H or n > 65535");
    const int n_seg = cdiv(D, PS_SEG);
    dim3 grid(cdiv(W, PS_TX), H, n), block(PS_TX, n_seg < PS_MAXSEG ? n_seg : PS_MAXSEG);
    if (Cin == 8)
        DV3D_LAUNCH((prob_softargmin_kernel<8>), grid, block, 0, (cudaStream_t)stream, x, D, H, W, weight, bias, depth_start,
                    depth_end, x_reg_out, depth_out);
    else
        DV3D_LAUNCH((prob_softargmin_kernel<16>), grid, block, 0, (cudaStream_t)stream, x, D, H, W, weight, bias, depth_start,
                    depth_end, x_reg_out, depth_out);
    DV3D_LAUNCHED();
    return DV3D_OK;
}
