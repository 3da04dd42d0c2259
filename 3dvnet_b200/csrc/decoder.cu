// PointFlow hypothesis decoder (refinement.py:8-13,20-25,42-44; lightningmodel.py:238-241).
// The decoder operand is kept as [n_pts, 8, C]: the 7 hypotheses of a point plus one all-zero
// padding row, so that the k=3 convolution over the hypothesis axis is three row-shifted GEMM
// slices whose out-of-range taps land on a zero row (no per-row masking in the main loop).
#include <math.h>

#include "gemm.cuh"

namespace dv3d {

// last Conv1d (C -> 1, bias) + softmax over the hypotheses + expected offset; one warp per point
__global__ void __launch_bounds__(256)
decoder_head_kernel(const float* __restrict__ x, long long Np, int n_hyp, int rows_per_point, int C, int ld,
                    const float* __restrict__ w, float bias, double offset, float* __restrict__ prob_out,
                    float* __restrict__ offset_out) {
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= Np) return;
    // logit[h] = bias + sum_c sum_t w[c][t] x[h + t - 1][c]  (w is [1, C, 3])
    float logit[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) logit[h] = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float w0 = __ldg(w + 3 * c), w1 = __ldg(w + 3 * c + 1), w2 = __ldg(w + 3 * c + 2);
        float xv[8];
#pragma unroll
        for (int h = 0; h < 8; ++h) xv[h] = (h < n_hyp) ? __ldg(x + ((size_t)p * rows_per_point + h) * ld + c) : 0.f;
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            float a = w1 * xv[h];
            if (h > 0) a = fmaf(w0, xv[h - 1], a);
            if (h < 7) a = fmaf(w2, xv[h + 1], a);
            logit[h] += a;
        }
    }
#pragma unroll
    for (int h = 0; h < 8; ++h) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) logit[h] += __shfl_xor_sync(0xffffffffu, logit[h], o);
        logit[h] += bias;
    }
    float m = -INFINITY;
#pragma unroll
    for (int h = 0; h < 8; ++h)
        if (h < n_hyp) m = fmaxf(m, logit[h]);
    float e[8], s = 0.f;
#pragma unroll
    for (int h = 0; h < 8; ++h) {
        e[h] = (h < n_hyp) ? expf(logit[h] - m) : 0.f;
        s += e[h];
    }
    const int n_side = (n_hyp - 1) / 2;
    const float lo = (float)(-(double)n_side * offset), hi = (float)((double)n_side * offset);
    float acc = 0.f;
#pragma unroll
    for (int h = 0; h < 8; ++h) {
        if (h < n_hyp) {
            float pr = e[h] / s;
            if (prob_out && lane == 0) prob_out[p * n_hyp + h] = pr;
            acc += linspace_torch(lo, hi, n_hyp, h) * pr;
        }
    }
    if (lane == 0) offset_out[p] = acc;
}

}  // namespace dv3d

using namespace dv3d;

extern "C" int dv3d_conv1d_bn_relu(const float* x, long long n_pts, int rows_per_point, int Cin, int ldx,
                                   const float* weight_tkn, const void* W_packed, const float* scale,
                                   const float* shift, int Cout, float* y, int ldy, void* workspace,
                                   size_t workspace_bytes, void* stream) {
    DV3D_REQUIRE(x && (weight_tkn || W_packed) && scale && shift && y && n_pts >= 0, "conv1d: bad arguments");
    DV3D_REQUIRE(rows_per_point == 8, "conv1d: the operand layout is [n_pts, 8, C] (7 hypotheses + 1 zero row)");
    GemmDesc d = {};
    d.n_slices = 3;
    for (int t = 0; t < 3; ++t) d.slice[t] = GemmSlice{x, nullptr, 0, t - 1, ldx, Cin};
    d.M = n_pts * rows_per_point;
    d.n_src_rows = d.M;
    d.N = Cout;
    d.W = weight_tkn;
    d.Wp = (const float*)W_packed;
    d.scale = scale;
    d.shift = shift;
    d.relu_out = 1;
    d.zero_row_mod = rows_per_point;
    d.zero_row_val = rows_per_point - 1;
    d.out = y;
    d.out_ld = ldy;
    cudaStream_t st = (cudaStream_t)stream;
    // Wave quantisation: the tensor-core kernel runs one 128-row tile per SM, so 148 q + r tiles take
    // q + 1 rounds.  When the last round is less than half full its tiles are launched separately
    // with the three taps split over three CTAs each (partials added in tap order by the last CTA
    // to arrive: deterministic) - C2: 196 tiles = one full round + 48 tiles x 3 splits.
    // A tile boundary is a point boundary (128 = 16 points x 8 rows), so the row shifts of the
    // second launch see the same zero rows as in the single launch.  Measured at C2: 81 -> 71 us for
    // K = 3 x 352, but a loss for K = 3 x 128 (the extra launch and the reduction cost more than the
    // saved half round), hence the K threshold.
    const long long tiles = (d.M + 127) / 128, q = tiles / kNumSMs, r = tiles % kNumSMs;
    const int split = r > 0 ? (int)(kNumSMs / r) : 1;
    constexpr size_t kCounterBytes = 4096;
    if (W_packed && workspace && split >= 2 && 3 * Cin >= 768 && workspace_bytes >= kCounterBytes + (size_t)r * 3 * 128 * Cout * sizeof(float) &&
        r * sizeof(int) <= kCounterBytes) {
        DV3D_REQUIRE(((uintptr_t)workspace & 255) == 0, "conv1d: workspace must be 256-byte aligned");
        const long long M_main = q * kNumSMs * 128;
        if (q > 0) {
            GemmDesc m = d;
            m.M = M_main;
            int rc = launch_gather_gemm(m, st);
            if (rc) return rc;
        }
        GemmDesc t = d;
        for (int i = 0; i < 3; ++i) t.slice[i].src = x + (size_t)M_main * ldx;
        t.out = y + (size_t)M_main * ldy;
        t.M = d.M - M_main;
        t.n_src_rows = t.M;
        t.split_counters = (int*)workspace;
        t.split_ws = (float*)((char*)workspace + kCounterBytes);
        t.split_ws_bytes = workspace_bytes - kCounterBytes;
        t.split_hint = split < 3 ? split : 3;
        return launch_gather_gemm(t, st);
    }
    return launch_gather_gemm(d, st);
}

extern "C" int dv3d_decoder_head(const float* x, long long n_pts, int n_hyp, int rows_per_point, int Cin, int ldx,
                                 const float* weight, float bias, double offset, float* prob_out, float* offset_out,
                                 void* stream) {
    DV3D_REQUIRE(x && weight && offset_out && n_pts >= 0 && n_hyp >= 1 && n_hyp <= 7 && (n_hyp & 1) &&
                     rows_per_point >= n_hyp && Cin > 0 && ldx >= Cin,
                 "decoder_head: bad arguments (n_hyp must be odd and <= 7)");
    if (n_pts == 0) return DV3D_OK;
    DV3D_LAUNCH((decoder_head_kernel), cdiv(n_pts, 8), 256, 0, (cudaStream_t)stream, x, n_pts, n_hyp, rows_per_point, Cin, ldx, weight, bias, offset, prob_out, offset_out);
    DV3D_LAUNCHED();
    return DV3D_OK;
}
