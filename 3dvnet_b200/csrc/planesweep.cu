// Fused plane-sweep warp + variance, and the point-level variant used for the feature-rich
// point cloud and the PointFlow hypotheses.
//
// Replaces, per reference view, the reference's chain
//   numpy plane points -> torch.inverse/bmm x3 -> F.grid_sample on a gathered feature copy
//   -> torch_scatter mean x2 -> x**2, sub              (mvsnet.py:187-216, utils.py:86-108)
// which materialises x_vox [E,C,D,h,w] and re-reads it three times, by one kernel that
// reads every source feature map through L1/L2 and writes the [C,D,h,w] variance slab once.
//
// Work decomposition (sm_100a, 148 SMs):
//   CTA = 256 threads = 32 consecutive plane pixels x 8 channel groups (4 channels each,
//   one float4 = 16 B per tap per thread; the 8 lanes of a pixel read one full 128-byte
//   NHWC line per tap).  The CTA marches through depth in chunks of 8 planes.
//   Phase 1 (per chunk): thread (pixel, k) projects plane k of its pixel into every source
//     view of the reference and leaves a 20-byte sample record (tap base, 4 weights) in
//     shared memory — the projection (2 IEEE divides + grid normalisation in the
//     reference's operation order) is computed once per (pixel, plane, edge), not once per
//     channel group.
//   Phase 2: thread (pixel, channel group) walks edges x planes, reloads its four taps only
//     when the record's tap base changes (consecutive depth planes mostly fall into the
//     same 2x2 footprint), and accumulates sum / sum of squares in registers.
//   Phase 3: variance -> padded shared tile [c][k][pixel] -> 128-byte coalesced row stores
//     into x_var[r][c][d][p0..p0+31].
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace dv3d {

constexpr int TP = 32;     // pixels per CTA
constexpr int KD = 8;      // planes (or hypotheses) per chunk
constexpr int EMAX = 8;    // edges whose records are staged per pass
constexpr int CS = KD * TP + 1;  // padded channel stride of the output tile (bank-conflict free)

struct SampleGeom {
    int Hf, Wf;
    float wm1, hm1;    // W-1, H-1 of the FULL image (mvsnet.py:205-206)
    double rwm1, rhm1; // 1 / (W-1), 1 / (H-1) in fp64 (div_small_int)
    float wfm1, hfm1;  // Wf-1, Hf-1 (grid_sample align_corners=True un-normalisation)
};

// The reference's fp32 operation chain is kept instruction for instruction, so that sample
// positions (and with them x_var and the back-projected points) are BIT-IDENTICAL to the CPU
// PyTorch path and not merely close: torch.bmm on K = 3/4 is a sequential FMA chain
// (a0*b0, then fma(a_k, b_k, acc) for k = 1..) - pinned empirically against MKL - and
// torch.inverse of a zero-skew pinhole K is [[1/fx, 0, -(cx/fx)], [0, 1/fy, -(cy/fy)], [0,0,1]].
// Per image the host-visible table holds Kinv (9) | P = K [R|t] (12) | R (9) | t (3).
constexpr int CAM_STRIDE = 36;
constexpr int CAM_KINV = 0, CAM_P = 9, CAM_R = 21, CAM_T = 30;

// a / n correctly rounded to fp32 for a SMALL INTEGER n (< 2^20), without the ~10-instruction IEEE
// division sequence: d = (double)a * RN64(1/n) is within 2^-52 of a/n, and a/n with a 24-bit
// numerator can be neither an fp32 rounding tie (that needs n | 2^k (2j+1) with an impossible
// magnitude) nor closer than ~2^-24/n to one, so rounding d to fp32 equals rounding a/n
// (double rounding is innocuous here).  Used for the divisions by the image size (W-1, H-1) and by the
// edge count; bit-exactness against the reference is asserted by the golden tests.
__device__ __forceinline__ float div_small_int(float a, double rcp_n) { return (float)__dmul_rn((double)a, rcp_n); }

__device__ __forceinline__ float chain3(float a0, float a1, float a2, float b0, float b1, float b2) {
    return __fmaf_rn(a2, b2, __fmaf_rn(a1, b1, __fmul_rn(a0, b0)));
}

// world point X = R^T (Kinv p - t) of the homogeneous pixel * depth vector p
// (utils.py:102-106 / lightningmodel.py:142-144)
__device__ __forceinline__ void backproject(const float* __restrict__ cam, float p0, float p1, float p2, float& X0,
                                            float& X1, float& X2) {
    const float* Ki = cam + CAM_KINV;
    const float* R = cam + CAM_R;
    const float* t = cam + CAM_T;
    float c0 = __fsub_rn(chain3(__ldg(Ki + 0), __ldg(Ki + 1), __ldg(Ki + 2), p0, p1, p2), __ldg(t + 0));
    float c1 = __fsub_rn(chain3(__ldg(Ki + 3), __ldg(Ki + 4), __ldg(Ki + 5), p0, p1, p2), __ldg(t + 1));
    float c2 = __fsub_rn(chain3(__ldg(Ki + 6), __ldg(Ki + 7), __ldg(Ki + 8), p0, p1, p2), __ldg(t + 2));
    X0 = chain3(__ldg(R + 0), __ldg(R + 3), __ldg(R + 6), c0, c1, c2);  // rows of R^T = columns of R
    X1 = chain3(__ldg(R + 1), __ldg(R + 4), __ldg(R + 7), c0, c1, c2);
    X2 = chain3(__ldg(R + 2), __ldg(R + 5), __ldg(R + 8), c0, c1, c2);
}

// One (pixel, plane, edge) sample: q = P_src [X;1] (mvsnet.py:199), z = |q_z| + 1e-8, normalise
// by the full image size, un-normalise as grid_sample does, split into tap base + bilinear
// weights.  Taps outside the map get weight 0 and a clamped address (zero padding per tap).
__device__ __forceinline__ void make_record(const float* __restrict__ P, float X0, float X1, float X2,
                                            const SampleGeom& g, int& rec, float4& wt) {
    // P is the CTA's shared-memory copy of the source camera's projection matrix (broadcast reads)
    float qx = __fmaf_rn(P[3], 1.f, chain3(P[0], P[1], P[2], X0, X1, X2));
    float qy = __fmaf_rn(P[7], 1.f, chain3(P[4], P[5], P[6], X0, X1, X2));
    float qz = __fmaf_rn(P[11], 1.f, chain3(P[8], P[9], P[10], X0, X1, X2));
    float zz = __fadd_rn(fabsf(qz), 1e-8f);
    float x = __fdiv_rn(qx, zz), y = __fdiv_rn(qy, zz);
    float gx = __fsub_rn(__fmul_rn(div_small_int(x, g.rwm1), 2.f), 1.f);
    float gy = __fsub_rn(__fmul_rn(div_small_int(y, g.rhm1), 2.f), 1.f);
    float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), g.wfm1);
    float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), g.hfm1);
    rec = 0;
    wt = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ix > -1.f && ix < (float)g.Wf && iy > -1.f && iy < (float)g.Hf) {
        float fx0 = floorf(ix), fy0 = floorf(iy);
        float tx = __fsub_rn(ix, fx0), ty = __fsub_rn(iy, fy0);
        int x0 = (int)fx0, y0 = (int)fy0;
        bool l = x0 >= 0, r = x0 + 1 <= g.Wf - 1, t = y0 >= 0, b = y0 + 1 <= g.Hf - 1;
        float e1 = __fsub_rn(1.f, tx), s1 = __fsub_rn(1.f, ty);  // distances to the east / south taps
        float wl = l ? e1 : 0.f, wr = r ? tx : 0.f;
        float wtp = t ? s1 : 0.f, wb = b ? ty : 0.f;
        wt = make_float4(__fmul_rn(wtp, wl), __fmul_rn(wtp, wr), __fmul_rn(wb, wl), __fmul_rn(wb, wr));  // nw ne sw se
        int cx = l ? x0 : 0, cy = t ? y0 : 0;
        int dx = (l && r) ? 1 : 0, dy = (t && b) ? 1 : 0;
        // when the left/top tap is the clamped one its weight is 0 and the right/bottom tap
        // must still address column x0+1 / row y0+1 == cx / cy: dx = dy = 0 does that.
        rec = ((cy * g.Wf + cx) << 2) | (dy << 1) | dx;
    }
}

// Packed fp32 (FFMA2 / FMUL2 / FADD2, sm_100): two channels per issue slot, every lane rounded
// to nearest exactly like the scalar instruction, so bit-exactness is unaffected.
// x * x with its own rounding.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2
// (even with --fmad=false and explicit .rn), which would skip the rounding of the square the
// reference performs before its scatter-add; scalar FMULs are never contracted.
__device__ __forceinline__ float2 sq2_rn(const float2& a) { return make_float2(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)); }
__device__ __forceinline__ float2 lo2(const float4& t) { return make_float2(t.x, t.y); }
__device__ __forceinline__ float2 hi2(const float4& t) { return make_float2(t.z, t.w); }

// Phase 2 for one staged pass of edges. NK = number of live planes/hypotheses in the chunk.
// Accumulators are channel pairs: acc_s[k][0] = channels (0,1), [1] = (2,3) of the thread's group.
// REUSE: keep the four taps while consecutive planes fall into the same 2x2 footprint (plane sweep: most
// far planes do).  Without it (PointFlow hypotheses, few and far apart) every sample issues its four
// independent loads unconditionally, so the unrolled loop keeps dozens of loads in flight instead of
// serialising 49 load -> use round trips per thread.
template <int NK, bool REUSE, bool FAST = false>
__device__ __forceinline__ void consume_edges(const float4* __restrict__ feats, const int* __restrict__ esrc,
                                              int e_begin, int n_e, int img_stride4, int Wf, int v, int g,
                                              const int (*s_rec)[KD][TP], const float4 (*s_wt)[KD][TP],
                                              float2 (&acc_s)[KD][2], float2 (&acc_q)[KD][2]) {
#pragma unroll(REUSE ? 1 : 2)
    for (int e = 0; e < n_e; ++e) {
        const float4* base = feats + (size_t)__ldg(esrc + e_begin + e) * img_stride4 + g;
        int prev = -1;
        float4 t00 = make_float4(0, 0, 0, 0), t01 = t00, t10 = t00, t11 = t00;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            int rec = s_rec[e][k][v];
            float4 wt = s_wt[e][k][v];
            if (!REUSE || rec != prev) {
                const float4* p = base + (size_t)(rec >> 2) * 8;
                int dx = (rec & 1) * 8, dy = ((rec >> 1) & 1) * Wf * 8;
                t00 = ldg4(p);
                t01 = ldg4(p + dx);
                t10 = ldg4(p + dy);
                t11 = ldg4(p + dy + dx);
                prev = rec;
            }
            // grid_sample (ATen CPU kernel): fma(se, w_se, fma(sw, w_sw, fma(ne, w_ne, nw * w_nw)))
            const float2 wx = make_float2(wt.x, wt.x), wy = make_float2(wt.y, wt.y), wz = make_float2(wt.z, wt.z),
                         ww = make_float2(wt.w, wt.w);
            float2 va = __fmul2_rn(lo2(t00), wx), vb = __fmul2_rn(hi2(t00), wx);
            va = __ffma2_rn(lo2(t01), wy, va); vb = __ffma2_rn(hi2(t01), wy, vb);
            va = __ffma2_rn(lo2(t10), wz, va); vb = __ffma2_rn(hi2(t10), wz, vb);
            va = __ffma2_rn(lo2(t11), ww, va); vb = __ffma2_rn(hi2(t11), ww, vb);
            // scatter 'mean' of x and of x**2: plain sums in edge order (mvsnet.py:214-215)
            acc_s[k][0] = __fadd2_rn(acc_s[k][0], va);
            acc_s[k][1] = __fadd2_rn(acc_s[k][1], vb);
            if (FAST) {   // tolerance mode: x*x + acc as one FFMA2 (the exact mode rounds the square first)
                acc_q[k][0] = __ffma2_rn(va, va, acc_q[k][0]);
                acc_q[k][1] = __ffma2_rn(vb, vb, acc_q[k][1]);
            } else {
                acc_q[k][0] = __fadd2_rn(acc_q[k][0], sq2_rn(va));
                acc_q[k][1] = __fadd2_rn(acc_q[k][1], sq2_rn(vb));
            }
        }
    }
}

__device__ __forceinline__ float var_of(float s, float q, double rcp_n) {
    float m = div_small_int(s, rcp_n);
    return __fsub_rn(div_small_int(q, rcp_n), __fmul_rn(m, m));  // E[x^2] - E[x]^2 (mvsnet.py:216)
}

// numpy.linspace(start, stop, n, dtype=float32): float64 arithmetic, last point == stop
__device__ __forceinline__ float linspace_np(double start, double stop, int n, int i) {
    if (n == 1) return (float)start;
    if (i == n - 1) return (float)stop;
    double step = (stop - start) / (double)(n - 1);
    return (float)__dadd_rn(__dmul_rn((double)i, step), start);
}

__global__ void __launch_bounds__(256, 2)
planesweep_var_kernel(const float4* __restrict__ feats, SampleGeom geom, const float* __restrict__ cams,
                      const int* __restrict__ ref_img, const int* __restrict__ rowptr, const int* __restrict__ esrc,
                      double d0, double d1, int D, int h, int w, int H, int W, int chunks_per_cta,
                      float* __restrict__ out) {
    pdl_wait();
    // dynamic shared memory (73 KB > the 48 KB static limit): weights | records | output tile
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 (*s_wt)[KD][TP] = reinterpret_cast<float4 (*)[KD][TP]>(smem_raw);
    int (*s_rec)[KD][TP] = reinterpret_cast<int (*)[KD][TP]>(smem_raw + sizeof(float4) * EMAX * KD * TP);
    float* s_out = reinterpret_cast<float*>(smem_raw + (sizeof(float4) + sizeof(int)) * EMAX * KD * TP);
    __shared__ float s_P[EMAX][12];  // projection matrices of the staged edges

    const int tid = threadIdx.x;
    const int r = blockIdx.y;
    const int P = h * w;
    const int p0 = blockIdx.x * TP;
    const int e0 = rowptr[r], e1 = rowptr[r + 1];
    const double n_edges = 1.0 / (double)(e1 - e0);  // reciprocal of the edge count (var_of)
    const int img_stride4 = geom.Hf * geom.Wf * 8;

    // phase-1 role: warp = plane pk, lane = pixel pv (conflict-free record stores);
    // phase-2 role: (pixel v, channel group g), the 8 lanes of a pixel share one 128 B tap line
    const int pv = tid & 31, pk = tid >> 5;
    const int v = tid >> 3, g = tid & 7;
    const int p = min(p0 + pv, P - 1);
    const float u = linspace_np(0.0, (double)(W - 1), w, p % w);
    const float vv = linspace_np(0.0, (double)(H - 1), h, p / w);

    const int chunk_begin = blockIdx.z * chunks_per_cta;
    const int n_chunks = (D + KD - 1) / KD;
    for (int ch = chunk_begin; ch < min(chunk_begin + chunks_per_cta, n_chunks); ++ch) {
        const int dbase = ch * KD;
        float2 acc_s[KD][2], acc_q[KD][2];
#pragma unroll
        for (int k = 0; k < KD; ++k) acc_s[k][0] = acc_s[k][1] = acc_q[k][0] = acc_q[k][1] = make_float2(0, 0);
        const float z = linspace_np(d0, d1, D, min(dbase + pk, D - 1));
        // frustum point of (pixel, plane): pixel * depth formed in fp64, rounded once (utils.py:96-100)
        float X0, X1, X2;
        backproject(cams + (size_t)__ldg(ref_img + r) * CAM_STRIDE, (float)((double)u * (double)z),
                    (float)((double)vv * (double)z), z, X0, X1, X2);

        for (int eb = e0; eb < e1; eb += EMAX) {
            const int n_e = min(EMAX, e1 - eb);
            __syncthreads();  // previous pass / chunk fully consumed (records, camera rows, output tile)
            if (tid < n_e * 12) s_P[tid / 12][tid % 12] = __ldg(cams + (size_t)__ldg(esrc + eb + tid / 12) * CAM_STRIDE + CAM_P + tid % 12);
            __syncthreads();
            for (int e = 0; e < n_e; ++e) {
                int rec;
                float4 wt;
                make_record(s_P[e], X0, X1, X2, geom, rec, wt);
                s_rec[e][pk][pv] = rec;
                s_wt[e][pk][pv] = wt;
            }
            __syncthreads();
            consume_edges<KD, true>(feats, esrc, eb, n_e, img_stride4, geom.Wf, v, g, s_rec, s_wt, acc_s, acc_q);
        }

        // phase 3: variance -> shared tile -> coalesced rows
#pragma unroll
        for (int k = 0; k < KD; ++k) {
            float* o = s_out + (4 * g) * CS + k * TP + v;
            o[0] = var_of(acc_s[k][0].x, acc_q[k][0].x, n_edges);
            o[CS] = var_of(acc_s[k][0].y, acc_q[k][0].y, n_edges);
            o[2 * CS] = var_of(acc_s[k][1].x, acc_q[k][1].x, n_edges);
            o[3 * CS] = var_of(acc_s[k][1].y, acc_q[k][1].y, n_edges);
        }
        __syncthreads();
        {
            const int warp = tid >> 5, lane = tid & 31;
            const bool ok = p0 + lane < P;
#pragma unroll 4
            for (int row = warp; row < 32 * KD; row += 8) {
                int c = row >> 3, k = row & 7;
                int d = dbase + k;
                if (ok && d < D)
                    out[(((size_t)r * 32 + c) * D + d) * P + p0 + lane] = s_out[c * CS + k * TP + lane];
            }
        }
        // the __syncthreads at the top of the next chunk's first pass orders these reads of s_out
        // (and the consumption of the records) before anything is rewritten.
    }
}

// ------------------------------------------------------------------------ tolerance mode of the plane sweep
// Same decomposition as planesweep_var_kernel, but the projection is evaluated the way the geometry allows instead
// of the way the reference happens to round it: the frustum point of (pixel, plane) is X(z) = z * dir + o with
// dir = R^T Kinv [u, v, 1] per pixel and o = -R^T t per reference, hence for a source view
//     q(z) = z * (M dir) + (M o + p4),        M = P_src[:, :3] with rows 0 / 1 pre-scaled by (Wf-1)/(W-1), (Hf-1)/(H-1)
// and the un-normalised sampling position is (q_x, q_y) * rcp(|q_z| + 1e-8): 3 FMA + 1 reciprocal + 2 multiplies
// per (pixel, plane, edge) instead of 12 FMA, two IEEE divisions and two fp64 divisions by the image size; the
// variance is q/n - (s/n)^2 with one reciprocal multiply.  Positions agree with the reference chain to ~2e-5 of a
// feature pixel, x_var to ~2e-5 of its scale (asserted in tests/test_gpu_warp_fast.py).
// MEASURED (profiles/r2_11_sweep_planesweep.md): 65.5 -> 57.3 us at C2, 625 -> 539 us at C5 - the projection was
// never the bottleneck; the kernel is bound by the per-sample gather + FMA work of phase 2 (4 L1 wavefronts and ~20
// issue slots per (pixel, plane, edge, 4 channels)), which both modes share.  12 % is not worth giving up the
// bit-exact slab, so this mode is an OPT-IN (DV3D_WARP=fast / dv3d_set_warp_mode(1)); exact is the default.
struct FastEdge {
    float m[3][3];  // M (rows 0, 1 pre-scaled)
    float b[3];     // M o + p4 (rows 0, 1 pre-scaled)
};

__device__ __forceinline__ void make_record_fast(float qx, float qy, float qz, const SampleGeom& g, int& rec, float4& wt) {
    const float r = __frcp_rn(fabsf(qz) + 1e-8f);
    const float ix = qx * r, iy = qy * r;
    rec = 0;
    wt = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ix > -1.f && ix < (float)g.Wf && iy > -1.f && iy < (float)g.Hf) {
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const float tx = ix - fx0, ty = iy - fy0;
        const int x0 = (int)fx0, y0 = (int)fy0;
        const bool l = x0 >= 0, rr = x0 + 1 <= g.Wf - 1, t = y0 >= 0, b = y0 + 1 <= g.Hf - 1;
        const float wl = l ? 1.f - tx : 0.f, wr = rr ? tx : 0.f;
        const float wtp = t ? 1.f - ty : 0.f, wb = b ? ty : 0.f;
        wt = make_float4(wtp * wl, wtp * wr, wb * wl, wb * wr);
        const int cx = l ? x0 : 0, cy = t ? y0 : 0;
        const int dx = (l && rr) ? 1 : 0, dy = (t && b) ? 1 : 0;
        rec = ((cy * g.Wf + cx) << 2) | (dy << 1) | dx;
    }
}

__global__ void __launch_bounds__(256, 2)
planesweep_var_fast_kernel(const float4* __restrict__ feats, SampleGeom geom, const float* __restrict__ cams,
                           const int* __restrict__ ref_img, const int* __restrict__ rowptr, const int* __restrict__ esrc,
                           float d0, float dstep, int D, int h, int w, int H, int W, int chunks_per_cta,
                           float* __restrict__ out) {
    pdl_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 (*s_wt)[KD][TP] = reinterpret_cast<float4 (*)[KD][TP]>(smem_raw);
    int (*s_rec)[KD][TP] = reinterpret_cast<int (*)[KD][TP]>(smem_raw + sizeof(float4) * EMAX * KD * TP);
    float* s_out = reinterpret_cast<float*>(smem_raw + (sizeof(float4) + sizeof(int)) * EMAX * KD * TP);
    __shared__ FastEdge s_E[EMAX];
    __shared__ float s_o[3];

    const int tid = threadIdx.x;
    const int r = blockIdx.y;
    const int P = h * w;
    const int p0 = blockIdx.x * TP;
    const int e0 = rowptr[r], e1 = rowptr[r + 1];
    const float inv_n = 1.f / (float)(e1 - e0);
    const int img_stride4 = geom.Hf * geom.Wf * 8;
    const float sx = geom.wfm1 / geom.wm1, sy = geom.hfm1 / geom.hm1;
    const float* cref = cams + (size_t)__ldg(ref_img + r) * CAM_STRIDE;

    const int pv = tid & 31, pk = tid >> 5;
    const int v = tid >> 3, g = tid & 7;
    const int p = min(p0 + pv, P - 1);
    const float u = linspace_np(0.0, (double)(W - 1), w, p % w);
    const float vv = linspace_np(0.0, (double)(H - 1), h, p / w);
    // dir = R^T (Kinv [u, v, 1]) of this thread's pixel; o = -R^T t of the reference
    float dir0, dir1, dir2;
    {
        const float* Ki = cref + CAM_KINV;
        const float* R = cref + CAM_R;
        const float c0 = fmaf(__ldg(Ki + 0), u, fmaf(__ldg(Ki + 1), vv, __ldg(Ki + 2)));
        const float c1 = fmaf(__ldg(Ki + 3), u, fmaf(__ldg(Ki + 4), vv, __ldg(Ki + 5)));
        const float c2 = fmaf(__ldg(Ki + 6), u, fmaf(__ldg(Ki + 7), vv, __ldg(Ki + 8)));
        dir0 = fmaf(__ldg(R + 0), c0, fmaf(__ldg(R + 3), c1, __ldg(R + 6) * c2));
        dir1 = fmaf(__ldg(R + 1), c0, fmaf(__ldg(R + 4), c1, __ldg(R + 7) * c2));
        dir2 = fmaf(__ldg(R + 2), c0, fmaf(__ldg(R + 5), c1, __ldg(R + 8) * c2));
    }
    if (tid < 3) {
        const float* R = cref + CAM_R;
        const float* t = cref + CAM_T;
        s_o[tid] = -(__ldg(R + tid) * __ldg(t + 0) + __ldg(R + 3 + tid) * __ldg(t + 1) + __ldg(R + 6 + tid) * __ldg(t + 2));
    }

    const int chunk_begin = blockIdx.z * chunks_per_cta;
    const int n_chunks = (D + KD - 1) / KD;
    for (int ch = chunk_begin; ch < min(chunk_begin + chunks_per_cta, n_chunks); ++ch) {
        const int dbase = ch * KD;
        float2 acc_s[KD][2], acc_q[KD][2];
#pragma unroll
        for (int k = 0; k < KD; ++k) acc_s[k][0] = acc_s[k][1] = acc_q[k][0] = acc_q[k][1] = make_float2(0, 0);
        const float z = fmaf((float)min(dbase + pk, D - 1), dstep, d0);

        for (int eb = e0; eb < e1; eb += EMAX) {
            const int n_e = min(EMAX, e1 - eb);
            __syncthreads();
            if (tid < n_e * 12) {   // M (pre-scaled) of every staged edge, then b = M o + p4
                const int e = tid / 12, i = tid % 12, row = i >> 2, col = i & 3;
                const float sc = row == 0 ? sx : (row == 1 ? sy : 1.f);
                const float val = __ldg(cams + (size_t)__ldg(esrc + eb + e) * CAM_STRIDE + CAM_P + i) * sc;
                if (col < 3) s_E[e].m[row][col] = val;
                else s_E[e].b[row] = val;
            }
            __syncthreads();
            if (tid < n_e * 3) {
                const int e = tid / 3, row = tid % 3;
                s_E[e].b[row] += s_E[e].m[row][0] * s_o[0] + s_E[e].m[row][1] * s_o[1] + s_E[e].m[row][2] * s_o[2];
            }
            __syncthreads();
            for (int e = 0; e < n_e; ++e) {
                const FastEdge& E = s_E[e];
                const float ax = fmaf(E.m[0][0], dir0, fmaf(E.m[0][1], dir1, E.m[0][2] * dir2));
                const float ay = fmaf(E.m[1][0], dir0, fmaf(E.m[1][1], dir1, E.m[1][2] * dir2));
                const float az = fmaf(E.m[2][0], dir0, fmaf(E.m[2][1], dir1, E.m[2][2] * dir2));
                int rec;
                float4 wt;
                make_record_fast(fmaf(z, ax, E.b[0]), fmaf(z, ay, E.b[1]), fmaf(z, az, E.b[2]), geom, rec, wt);
                s_rec[e][pk][pv] = rec;
                s_wt[e][pk][pv] = wt;
            }
            __syncthreads();
            consume_edges<KD, true, true>(feats, esrc, eb, n_e, img_stride4, geom.Wf, v, g, s_rec, s_wt, acc_s, acc_q);
        }
#pragma unroll
        for (int k = 0; k < KD; ++k) {
            float* o = s_out + (4 * g) * CS + k * TP + v;
            float m;
            m = acc_s[k][0].x * inv_n; o[0] = fmaf(acc_q[k][0].x, inv_n, -m * m);
            m = acc_s[k][0].y * inv_n; o[CS] = fmaf(acc_q[k][0].y, inv_n, -m * m);
            m = acc_s[k][1].x * inv_n; o[2 * CS] = fmaf(acc_q[k][1].x, inv_n, -m * m);
            m = acc_s[k][1].y * inv_n; o[3 * CS] = fmaf(acc_q[k][1].y, inv_n, -m * m);
        }
        __syncthreads();
        {
            const int warp = tid >> 5, lane = tid & 31;
            const bool ok = p0 + lane < P;
#pragma unroll 4
            for (int row = warp; row < 32 * KD; row += 8) {
                const int c = row >> 3, k = row & 7, d = dbase + k;
                if (ok && d < D) out[(((size_t)r * 32 + c) * D + d) * P + p0 + lane] = s_out[c * CS + k * TP + lane];
            }
        }
    }
}

// Point-level variant: the "planes" of a pixel are its 2n+1 depth hypotheses around the
// current depth estimate (lightningmodel.py:201-205); outputs are point-major.
// MINB = 1: up to 255 registers, dozens of tap loads in flight per thread - best while the grid does not fill the
// GPU (98 CTAs at C2); MINB = 2: two CTAs per SM once there are more CTAs than SMs (8+ reference views per call)
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
points_var_kernel(const float4* __restrict__ feats, SampleGeom geom, const float* __restrict__ cams,
                  const int* __restrict__ ref_img, const int* __restrict__ rowptr, const int* __restrict__ esrc,
                  const float* __restrict__ depth, int h, int w, int H, int W, int n_side, double offset,
                  float* __restrict__ pts_out, float* __restrict__ feat_out, int rows_per_point, int feat_stride,
                  int feat_off) {
    pdl_wait();
    __shared__ int s_rec[EMAX][KD][TP];
    __shared__ float4 s_wt[EMAX][KD][TP];
    __shared__ float s_P[EMAX][12];

    const int tid = threadIdx.x;
    const int r = blockIdx.y;
    const int P = h * w;
    const int p0 = blockIdx.x * TP;
    const int e0 = rowptr[r], e1 = rowptr[r + 1];
    const double n_edges = 1.0 / (double)(e1 - e0);  // reciprocal of the edge count (var_of)
    const int img_stride4 = geom.Hf * geom.Wf * 8;
    const int n_hyp = 2 * n_side + 1;

    const int pv = tid & 31, pk = tid >> 5;   // phase 1: warp = hypothesis, lane = pixel
    const int v = tid >> 3, g = tid & 7;       // phase 2: (pixel, channel group)
    const int p = min(p0 + pv, P - 1);
    const bool live = p0 + pv < P;
    const float u = linspace_np(0.0, (double)(W - 1), w, p % w);
    const float vv = linspace_np(0.0, (double)(H - 1), h, p / w);
    // hypothesis depth: depth + i*offset with i*offset rounded to fp32 first (python float
    // times int, then a tensor + scalar add, lightningmodel.py:203)
    const float dpt = __ldg(depth + (size_t)r * P + p);
    const float z = __fadd_rn(dpt, (float)((double)(pk - n_side) * offset));

    // world point of hypothesis pk: pts_img * depth is an fp32 product here (lightningmodel.py:142)
    float X0, X1, X2;
    backproject(cams + (size_t)__ldg(ref_img + r) * CAM_STRIDE, __fmul_rn(u, z), __fmul_rn(vv, z), z, X0, X1, X2);
    if (live && pk < n_hyp) {
        float* o = pts_out + (((size_t)r * P + p) * n_hyp + pk) * 3;
        o[0] = X0;
        o[1] = X1;
        o[2] = X2;
    }

    float2 acc_s[KD][2], acc_q[KD][2];
#pragma unroll
    for (int k = 0; k < KD; ++k) acc_s[k][0] = acc_s[k][1] = acc_q[k][0] = acc_q[k][1] = make_float2(0, 0);

    for (int eb = e0; eb < e1; eb += EMAX) {
        const int n_e = min(EMAX, e1 - eb);
        __syncthreads();
        if (tid < n_e * 12) s_P[tid / 12][tid % 12] = __ldg(cams + (size_t)__ldg(esrc + eb + tid / 12) * CAM_STRIDE + CAM_P + tid % 12);
        __syncthreads();
        if (pk < n_hyp) {
            for (int e = 0; e < n_e; ++e) {
                int rec;
                float4 wt;
                make_record(s_P[e], X0, X1, X2, geom, rec, wt);
                s_rec[e][pk][pv] = rec;
                s_wt[e][pk][pv] = wt;
            }
        }
        __syncthreads();
        if (n_hyp == 1)
            consume_edges<1, false>(feats, esrc, eb, n_e, img_stride4, geom.Wf, v, g, s_rec, s_wt, acc_s, acc_q);
        else
            consume_edges<7, false>(feats, esrc, eb, n_e, img_stride4, geom.Wf, v, g, s_rec, s_wt, acc_s, acc_q);
    }
    if (p0 + v < P) {
        const int p = p0 + v;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            if (k < n_hyp) {
                float4 o;
                o.x = var_of(acc_s[k][0].x, acc_q[k][0].x, n_edges);
                o.y = var_of(acc_s[k][0].y, acc_q[k][0].y, n_edges);
                o.z = var_of(acc_s[k][1].x, acc_q[k][1].x, n_edges);
                o.w = var_of(acc_s[k][1].y, acc_q[k][1].y, n_edges);
                *reinterpret_cast<float4*>(feat_out + (((size_t)r * P + p) * rows_per_point + k) * feat_stride +
                                           feat_off + 4 * g) = o;
            }
        }
    }
}

// ------------------------------------------------------------------------ backward (w.r.t. the feature maps)
// The reference trains through F.grid_sample + the two scatter means (mvsnet.py:209-216, lightningmodel.py:165-169;
// the sampling grid itself is built under no_grad, mvsnet.py:187 / lightningmodel.py:147,190):
//     x_var = mean_e(x_e^2) - mean_e(x_e)^2   =>   d x_var / d x_e = (2/n) (x_e - mean),
//     d x_e / d F[src_e, tap] = bilinear weight of the tap.
// The backward kernels recompute the sample records and the samples exactly as the forward kernels do (nothing is
// saved but the inputs), form g * (2/n) (x_e - mean) per (pixel, plane, edge, channel) and add w_tap times that to
// the four taps of the NHWC gradient map with 16-byte vector atomics.  Consecutive planes that fall into the same
// 2x2 footprint are summed in registers first (one flush per footprint instead of one per plane).  Like ATen's own
// CUDA grid_sampler backward, the accumulation order of the atomics is not fixed: results are reproducible to
// fp32 rounding, not bit for bit.
__device__ __forceinline__ void red_add4(float* addr, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void fma4(float4& a, float w, const float4& g) {
    a.x = fmaf(w, g.x, a.x); a.y = fmaf(w, g.y, a.y); a.z = fmaf(w, g.z, a.z); a.w = fmaf(w, g.w, a.w);
}

// sum over a staged pass of edges of the sampled feature (4 channels of group g) for NK planes
template <int NK>
__device__ __forceinline__ void sum_edges(const float4* __restrict__ feats, const int* __restrict__ esrc, int e_begin,
                                          int n_e, int img_stride4, int Wf, int v, int g, const int (*s_rec)[KD][TP],
                                          const float4 (*s_wt)[KD][TP], float4 (&sum)[KD]) {
    for (int e = 0; e < n_e; ++e) {
        const float4* base = feats + (size_t)__ldg(esrc + e_begin + e) * img_stride4 + g;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const int rec = s_rec[e][k][v];
            const float4 wt = s_wt[e][k][v];
            const float4* p = base + (size_t)(rec >> 2) * 8;
            const int dx = (rec & 1) * 8, dy = ((rec >> 1) & 1) * Wf * 8;
            const float4 t00 = ldg4(p), t01 = ldg4(p + dx), t10 = ldg4(p + dy), t11 = ldg4(p + dy + dx);
            float4 x = make_float4(t00.x * wt.x, t00.y * wt.x, t00.z * wt.x, t00.w * wt.x);
            fma4(x, wt.y, t01);
            fma4(x, wt.z, t10);
            fma4(x, wt.w, t11);
            sum[k].x += x.x; sum[k].y += x.y; sum[k].z += x.z; sum[k].w += x.w;
        }
    }
}

// scatter pass: grad_x_e = gsc[k] * (x_e - mean[k]) with gsc = g * 2/n, pushed to the four taps
template <int NK>
__device__ __forceinline__ void scatter_edges(const float4* __restrict__ feats, float* __restrict__ grad_feats,
                                              const int* __restrict__ esrc, int e_begin, int n_e, int img_stride4, int Wf,
                                              int v, int g, const int (*s_rec)[KD][TP], const float4 (*s_wt)[KD][TP],
                                              const float4 (&mean)[KD], const float4 (&gsc)[KD]) {
    for (int e = 0; e < n_e; ++e) {
        const int src = __ldg(esrc + e_begin + e);
        const float4* base = feats + (size_t)src * img_stride4 + g;
        float* gbase = grad_feats + ((size_t)src * img_stride4 + g) * 4;
        int prev = -1;
        float4 a00 = make_float4(0, 0, 0, 0), a01 = a00, a10 = a00, a11 = a00;
        auto flush = [&]() {
            if (prev < 0) return;
            float* p = gbase + (size_t)(prev >> 2) * 32;
            const int dx = (prev & 1) * 32, dy = ((prev >> 1) & 1) * Wf * 32;
            // clamped (out-of-bounds) taps alias a live one with weight 0: their sums are exactly 0 unless aliased
            red_add4(p, a00);
            if (dx) red_add4(p + dx, a01); else red_add4(p, a01);
            if (dy) red_add4(p + dy, a10); else red_add4(p, a10);
            red_add4(p + dy + dx, a11);
            a00 = a01 = a10 = a11 = make_float4(0, 0, 0, 0);
        };
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const int rec = s_rec[e][k][v];
            const float4 wt = s_wt[e][k][v];
            if (rec != prev) {
                flush();
                prev = rec;
            }
            const float4* p = base + (size_t)(rec >> 2) * 8;
            const int dx = (rec & 1) * 8, dy = ((rec >> 1) & 1) * Wf * 8;
            const float4 t00 = ldg4(p), t01 = ldg4(p + dx), t10 = ldg4(p + dy), t11 = ldg4(p + dy + dx);
            float4 x = make_float4(t00.x * wt.x, t00.y * wt.x, t00.z * wt.x, t00.w * wt.x);
            fma4(x, wt.y, t01);
            fma4(x, wt.z, t10);
            fma4(x, wt.w, t11);
            const float4 gx = make_float4(gsc[k].x * (x.x - mean[k].x), gsc[k].y * (x.y - mean[k].y),
                                          gsc[k].z * (x.z - mean[k].z), gsc[k].w * (x.w - mean[k].w));
            fma4(a00, wt.x, gx);
            fma4(a01, wt.y, gx);
            fma4(a10, wt.z, gx);
            fma4(a11, wt.w, gx);
        }
        flush();
    }
}

__global__ void __launch_bounds__(256, 1)
planesweep_var_bwd_kernel(const float4* __restrict__ feats, SampleGeom geom, const float* __restrict__ cams,
                          const int* __restrict__ ref_img, const int* __restrict__ rowptr, const int* __restrict__ esrc,
                          double d0, double d1, int D, int h, int w, int H, int W, int chunks_per_cta,
                          const float* __restrict__ grad_out, float* __restrict__ grad_feats) {
    pdl_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 (*s_wt)[KD][TP] = reinterpret_cast<float4 (*)[KD][TP]>(smem_raw);
    int (*s_rec)[KD][TP] = reinterpret_cast<int (*)[KD][TP]>(smem_raw + sizeof(float4) * EMAX * KD * TP);
    float* s_g = reinterpret_cast<float*>(smem_raw + (sizeof(float4) + sizeof(int)) * EMAX * KD * TP);
    __shared__ float s_P[EMAX][12];

    const int tid = threadIdx.x;
    const int r = blockIdx.y;
    const int P = h * w;
    const int p0 = blockIdx.x * TP;
    const int e0 = rowptr[r], e1 = rowptr[r + 1];
    const float inv_n = 1.f / (float)(e1 - e0);
    const int img_stride4 = geom.Hf * geom.Wf * 8;
    const int pv = tid & 31, pk = tid >> 5;
    const int v = tid >> 3, g = tid & 7;
    const int p = min(p0 + pv, P - 1);
    const float u = linspace_np(0.0, (double)(W - 1), w, p % w);
    const float vv = linspace_np(0.0, (double)(H - 1), h, p / w);

    const int chunk_begin = blockIdx.z * chunks_per_cta;
    const int n_chunks = (D + KD - 1) / KD;
    for (int ch = chunk_begin; ch < min(chunk_begin + chunks_per_cta, n_chunks); ++ch) {
        const int dbase = ch * KD;
        __syncthreads();
        {   // the gradient tile [c][k][pixel] of this chunk, 128-byte rows
            const int warp = tid >> 5, lane = tid & 31;
            const bool ok = p0 + lane < P;
            for (int row = warp; row < 32 * KD; row += 8) {
                const int c = row >> 3, k = row & 7, d = dbase + k;
                s_g[c * CS + k * TP + lane] = (ok && d < D) ? __ldg(grad_out + (((size_t)r * 32 + c) * D + d) * P + p0 + lane) : 0.f;
            }
        }
        const float z = linspace_np(d0, d1, D, min(dbase + pk, D - 1));
        float X0, X1, X2;
        backproject(cams + (size_t)__ldg(ref_img + r) * CAM_STRIDE, (float)((double)u * (double)z),
                    (float)((double)vv * (double)z), z, X0, X1, X2);
        float4 mean[KD], gsc[KD];
#pragma unroll
        for (int k = 0; k < KD; ++k) mean[k] = make_float4(0, 0, 0, 0);
        for (int pass = 0; pass < 2; ++pass) {
            for (int eb = e0; eb < e1; eb += EMAX) {
                const int n_e = min(EMAX, e1 - eb);
                __syncthreads();
                if (tid < n_e * 12) s_P[tid / 12][tid % 12] = __ldg(cams + (size_t)__ldg(esrc + eb + tid / 12) * CAM_STRIDE + CAM_P + tid % 12);
                __syncthreads();
                for (int e = 0; e < n_e; ++e) {
                    int rec;
                    float4 wt;
                    make_record(s_P[e], X0, X1, X2, geom, rec, wt);
                    s_rec[e][pk][pv] = rec;
                    s_wt[e][pk][pv] = wt;
                }
                __syncthreads();
                if (pass == 0)
                    sum_edges<KD>(feats, esrc, eb, n_e, img_stride4, geom.Wf, v, g, s_rec, s_wt, mean);
                else
                    scatter_edges<KD>(feats, grad_feats, esrc, eb, n_e, img_stride4, geom.Wf, v, g, s_rec, s_wt, mean, gsc);
            }
            if (pass == 0) {
#pragma unroll
                for (int k = 0; k < KD; ++k) {
                    mean[k].x *= inv_n; mean[k].y *= inv_n; mean[k].z *= inv_n; mean[k].w *= inv_n;
                    const float* gp = s_g + (4 * g) * CS + k * TP + v;
                    const float f = 2.f * inv_n;
                    gsc[k] = make_float4(gp[0] * f, gp[CS] * f, gp[2 * CS] * f, gp[3 * CS] * f);
                }
            }
        }
    }
}

// point-level backward: grad of the variance features [n_pts, rows, ld] (channels feat_off..+32) w.r.t. the maps
__global__ void __launch_bounds__(256, 1)
points_var_bwd_kernel(const float4* __restrict__ feats, SampleGeom geom, const float* __restrict__ cams,
                      const int* __restrict__ ref_img, const int* __restrict__ rowptr, const int* __restrict__ esrc,
                      const float* __restrict__ depth, int h, int w, int H, int W, int n_side, double offset,
                      const float* __restrict__ grad_feat, int rows_per_point, int feat_stride, int feat_off,
                      float* __restrict__ grad_feats) {
    pdl_wait();
    __shared__ int s_rec[EMAX][KD][TP];
    __shared__ float4 s_wt[EMAX][KD][TP];
    __shared__ float s_P[EMAX][12];
    const int tid = threadIdx.x;
    const int r = blockIdx.y;
    const int P = h * w;
    const int p0 = blockIdx.x * TP;
    const int e0 = rowptr[r], e1 = rowptr[r + 1];
    const float inv_n = 1.f / (float)(e1 - e0);
    const int img_stride4 = geom.Hf * geom.Wf * 8;
    const int n_hyp = 2 * n_side + 1;
    const int pv = tid & 31, pk = tid >> 5;
    const int v = tid >> 3, g = tid & 7;
    const int p = min(p0 + pv, P - 1);
    const float u = linspace_np(0.0, (double)(W - 1), w, p % w);
    const float vv = linspace_np(0.0, (double)(H - 1), h, p / w);
    const float dpt = __ldg(depth + (size_t)r * P + p);
    const float z = __fadd_rn(dpt, (float)((double)(pk - n_side) * offset));
    float X0, X1, X2;
    backproject(cams + (size_t)__ldg(ref_img + r) * CAM_STRIDE, __fmul_rn(u, z), __fmul_rn(vv, z), z, X0, X1, X2);

    float4 mean[KD], gsc[KD];
#pragma unroll
    for (int k = 0; k < KD; ++k) mean[k] = gsc[k] = make_float4(0, 0, 0, 0);
    const bool live = p0 + v < P;
    for (int pass = 0; pass < 2; ++pass) {
        for (int eb = e0; eb < e1; eb += EMAX) {
            const int n_e = min(EMAX, e1 - eb);
            __syncthreads();
            if (tid < n_e * 12) s_P[tid / 12][tid % 12] = __ldg(cams + (size_t)__ldg(esrc + eb + tid / 12) * CAM_STRIDE + CAM_P + tid % 12);
            __syncthreads();
            if (pk < n_hyp) {
                for (int e = 0; e < n_e; ++e) {
                    int rec;
                    float4 wt;
                    make_record(s_P[e], X0, X1, X2, geom, rec, wt);
                    s_rec[e][pk][pv] = rec;
                    s_wt[e][pk][pv] = wt;
                }
            }
            __syncthreads();
            if (pass == 0) {
                if (n_hyp == 1) sum_edges<1>(feats, esrc, eb, n_e, img_stride4, geom.Wf, v, g, s_rec, s_wt, mean);
                else sum_edges<7>(feats, esrc, eb, n_e, img_stride4, geom.Wf, v, g, s_rec, s_wt, mean);
            } else if (live) {
                if (n_hyp == 1) scatter_edges<1>(feats, grad_feats, esrc, eb, n_e, img_stride4, geom.Wf, v, g, s_rec, s_wt, mean, gsc);
                else scatter_edges<7>(feats, grad_feats, esrc, eb, n_e, img_stride4, geom.Wf, v, g, s_rec, s_wt, mean, gsc);
            }
        }
        if (pass == 0) {
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                if (k < n_hyp) {
                    mean[k].x *= inv_n; mean[k].y *= inv_n; mean[k].z *= inv_n; mean[k].w *= inv_n;
                    float4 gv = make_float4(0, 0, 0, 0);
                    if (live)
                        gv = __ldg(reinterpret_cast<const float4*>(grad_feat + (((size_t)r * P + p0 + v) * rows_per_point + k) * feat_stride + feat_off + 4 * g));
                    const float f = 2.f * inv_n;
                    gsc[k] = make_float4(gv.x * f, gv.y * f, gv.z * f, gv.w * f);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------ camera algebra
__device__ void inv3(const double* m, double* o) {
    double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    double det = a * A + b * B + c * C;
    double id = 1.0 / det;
    o[0] = A * id; o[1] = -(b * i - c * h) * id; o[2] = (b * f - c * e) * id;
    o[3] = B * id; o[4] = (a * i - c * g) * id;  o[5] = -(a * f - c * d) * id;
    o[6] = C * id; o[7] = -(a * h - b * g) * id; o[8] = (a * e - b * d) * id;
}
__device__ void load3x3(const float* p, double* o) {
    for (int i = 0; i < 9; ++i) o[i] = (double)p[i];
}

// Per-image camera table: Kinv | P = K [R|t] | R | t  (see CAM_* above).
__global__ void camera_tables_kernel(const float* __restrict__ R, const float* __restrict__ t,
                                     const float* __restrict__ K, int n, float* __restrict__ out) {
    pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* k = K + 9 * i;
    const float* r = R + 9 * i;
    const float* tv = t + 3 * i;
    float* o = out + (size_t)i * CAM_STRIDE;
    if (k[1] == 0.f && k[3] == 0.f && k[6] == 0.f && k[7] == 0.f && k[8] == 1.f) {
        // zero-skew pinhole: what torch.inverse (LAPACK getrf/getri, no pivoting needed) returns
        o[CAM_KINV + 0] = __fdiv_rn(1.f, k[0]); o[CAM_KINV + 1] = 0.f; o[CAM_KINV + 2] = -__fdiv_rn(k[2], k[0]);
        o[CAM_KINV + 3] = 0.f; o[CAM_KINV + 4] = __fdiv_rn(1.f, k[4]); o[CAM_KINV + 5] = -__fdiv_rn(k[5], k[4]);
        o[CAM_KINV + 6] = 0.f; o[CAM_KINV + 7] = 0.f; o[CAM_KINV + 8] = 1.f;
    } else {
        double kd[9], ki[9];
        load3x3(k, kd);
        inv3(kd, ki);
        for (int a = 0; a < 9; ++a) o[CAM_KINV + a] = (float)ki[a];
    }
    // P = bmm(K, cat(R, t)) (mvsnet.py:196-197): for matrices this small torch.bmm takes its
    // plain-loop path, ((k0 r0 + k1 r1) + k2 r2) with every product and sum rounded (no FMA)
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 4; ++b) {
            float r0 = b < 3 ? r[b] : tv[0], r1 = b < 3 ? r[3 + b] : tv[1], r2 = b < 3 ? r[6 + b] : tv[2];
            o[CAM_P + a * 4 + b] = __fadd_rn(__fadd_rn(__fmul_rn(k[a * 3], r0), __fmul_rn(k[a * 3 + 1], r1)),
                                             __fmul_rn(k[a * 3 + 2], r2));
        }
    for (int a = 0; a < 9; ++a) o[CAM_R + a] = r[a];
    for (int a = 0; a < 3; ++a) o[CAM_T + a] = tv[a];
    for (int a = 33; a < CAM_STRIDE; ++a) o[a] = 0.f;
}

// NCHW -> NHWC through a padded 32x32 shared tile (both sides coalesced)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int HW) {
    pdl_wait();
    __shared__ float tile[32][33];
    int n = blockIdx.z;
    int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* s = src + (size_t)n * C * HW;
    float* d = dst + (size_t)n * C * HW;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, hw = hw0 + threadIdx.x;
        if (c < C && hw < HW) tile[i][threadIdx.x] = s[(size_t)c * HW + hw];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int hw = hw0 + i, c = c0 + threadIdx.x;
        if (c < C && hw < HW) d[(size_t)hw * C + c] = tile[threadIdx.x][i];
    }
}

static SampleGeom make_geom(int Hf, int Wf, int H, int W) {
    SampleGeom g;
    g.Hf = Hf; g.Wf = Wf;
    g.wm1 = (float)(W - 1); g.hm1 = (float)(H - 1);
    g.rwm1 = 1.0 / (double)(W - 1); g.rhm1 = 1.0 / (double)(H - 1);
    g.wfm1 = (float)(Wf - 1); g.hfm1 = (float)(Hf - 1);
    return g;
}

}  // namespace dv3d

using namespace dv3d;

// 0 = exact (default: the reference's fp32 operation chain, bit-exact x_var), 1 = fast (tolerance mode, DV3D_WARP=fast)
static std::atomic<int> g_warp_mode{[] {
    const char* e = getenv("DV3D_WARP");
    return (e && (e[0] == 'f' || e[0] == 'F' || e[0] == '1')) ? 1 : 0;
}()};
extern "C" int dv3d_set_warp_mode(int mode) {
    DV3D_REQUIRE(mode == 0 || mode == 1, "set_warp_mode: 0 = exact, 1 = fast; got %d", mode);
    g_warp_mode.store(mode);
    return DV3D_OK;
}
extern "C" int dv3d_get_warp_mode(void) { return g_warp_mode.load(); }

extern "C" int dv3d_nchw_to_nhwc(const float* src, float* dst, int n, int C, int HW, void* stream) {
    DV3D_REQUIRE(src && dst && n >= 0 && C > 0 && HW > 0, "nchw_to_nhwc: bad arguments");
    if (n == 0) return DV3D_OK;
    dim3 grid(cdiv(HW, 32), cdiv(C, 32), n), block(32, 8);
    DV3D_LAUNCH((nchw_to_nhwc_kernel), grid, block, 0, (cudaStream_t)stream, src, dst, C, HW);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_camera_tables(const float* rotmats, const float* tvecs, const float* K, int n_imgs, float* out,
                                  void* stream) {
    DV3D_REQUIRE(rotmats && tvecs && K && out && n_imgs >= 0, "camera_tables: bad arguments");
    if (n_imgs == 0) return DV3D_OK;
    DV3D_LAUNCH((camera_tables_kernel), cdiv(n_imgs, 64), 64, 0, (cudaStream_t)stream, rotmats, tvecs, K, n_imgs, out);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_planesweep_var(const float* feats_nhwc, int n_imgs, int C, int Hf, int Wf, const float* cams,
                                   const int* ref_img, const int* edge_rowptr, const int* edge_src, int n_ref,
                                   double depth_start,
                                   double depth_interval, int D, int h, int w, int H, int W, float* x_var,
                                   void* stream) {
    DV3D_REQUIRE(C == 32, "planesweep_var: C must be 32 (IMG_FEAT_DIM, mv3d/config.py:42), got %d", C);
    DV3D_REQUIRE(feats_nhwc && cams && ref_img && edge_rowptr && edge_src && x_var, "planesweep_var: null pointer");
    DV3D_REQUIRE(n_imgs > 0 && Hf > 1 && Wf > 1 && D > 0 && h > 0 && w > 0 && H > 1 && W > 1 && n_ref >= 0,
                 "planesweep_var: bad shape");
    DV3D_REQUIRE((long long)Hf * Wf < (1 << 29), "planesweep_var: feature map too large for the record encoding");
    if (n_ref == 0) return DV3D_OK;
    const int P = h * w, tiles = cdiv(P, TP), n_chunks = cdiv(D, KD);
    // enough CTAs for ~4 waves of 2 CTAs/SM when the batch is small, whole depth per CTA otherwise
    int dsplit = cdiv(4 * 2 * kNumSMs, (long long)tiles * n_ref);
    dsplit = dsplit < 1 ? 1 : (dsplit > n_chunks ? n_chunks : dsplit);
    const int chunks_per_cta = cdiv(n_chunks, dsplit);
    dim3 grid(tiles, n_ref, cdiv(n_chunks, chunks_per_cta));
    DV3D_REQUIRE(n_ref <= 65535, "planesweep_var: n_ref > 65535");
    double d1 = depth_start + depth_interval * (D - 1);
    const size_t smem = (sizeof(float4) + sizeof(int)) * EMAX * KD * TP + sizeof(float) * 32 * CS;
    if (g_warp_mode == 1) {
        static std::atomic<unsigned long long> fattr{0};
        DV3D_FUNC_SMEM_ONCE(fattr, (planesweep_var_fast_kernel), (int)smem);
        DV3D_LAUNCH((planesweep_var_fast_kernel), grid, 256, smem, (cudaStream_t)stream, reinterpret_cast<const float4*>(feats_nhwc), make_geom(Hf, Wf, H, W), cams, ref_img, edge_rowptr, edge_src, (float)depth_start, (float)depth_interval, D, h, w, H, W, chunks_per_cta, x_var);
        DV3D_LAUNCHED();
        return DV3D_OK;
    }
    static std::atomic<unsigned long long> attr{0};
    DV3D_FUNC_SMEM_ONCE(attr, (planesweep_var_kernel), (int)smem);
    DV3D_LAUNCH((planesweep_var_kernel), grid, 256, smem, (cudaStream_t)stream, reinterpret_cast<const float4*>(feats_nhwc), make_geom(Hf, Wf, H, W), cams, ref_img, edge_rowptr, edge_src, depth_start, d1, D, h, w, H, W, chunks_per_cta, x_var);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_points_var(const float* feats_nhwc, int n_imgs, int C, int Hf, int Wf, const float* cams,
                               const int* ref_img, const int* edge_rowptr, const int* edge_src,
                               const float* depth, int n_ref, int h, int w, int H, int W, int n_side, double offset,
                               float* pts_out, float* feat_out, int rows_per_point, int feat_stride, int feat_off,
                               void* stream) {
    DV3D_REQUIRE(C == 32, "points_var: C must be 32, got %d", C);
    DV3D_REQUIRE(n_side == 0 || n_side == 3, "points_var: n_side must be 0 (point cloud) or 3 (PointFlow), got %d",
                 n_side);
    DV3D_REQUIRE(feats_nhwc && cams && ref_img && edge_rowptr && edge_src && depth && pts_out && feat_out,
                 "points_var: null pointer");
    DV3D_REQUIRE(rows_per_point >= 2 * n_side + 1, "points_var: rows_per_point < number of hypotheses");
    DV3D_REQUIRE(feat_stride % 4 == 0 && feat_off % 4 == 0 && feat_off + C <= feat_stride,
                 "points_var: feat_stride/feat_off must be multiples of 4 with room for C channels");
    DV3D_REQUIRE(n_imgs > 0 && Hf > 1 && Wf > 1 && h > 0 && w > 0 && n_ref >= 0 && n_ref <= 65535, "points_var: bad shape");
    if (n_ref == 0) return DV3D_OK;
    dim3 grid(cdiv(h * w, TP), n_ref);
    if ((long long)grid.x * grid.y > kNumSMs)
        DV3D_LAUNCH((points_var_kernel<2>), grid, 256, 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(feats_nhwc), make_geom(Hf, Wf, H, W), cams, ref_img, edge_rowptr, edge_src, depth, h, w, H, W, n_side, offset, pts_out, feat_out, rows_per_point, feat_stride, feat_off);
    else
        DV3D_LAUNCH((points_var_kernel<1>), grid, 256, 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(feats_nhwc), make_geom(Hf, Wf, H, W), cams, ref_img, edge_rowptr, edge_src, depth, h, w, H, W, n_side, offset, pts_out, feat_out, rows_per_point, feat_stride, feat_off);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_planesweep_var_backward(const float* feats_nhwc, int n_imgs, int C, int Hf, int Wf, const float* cams,
                                            const int* ref_img, const int* edge_rowptr, const int* edge_src, int n_ref,
                                            double depth_start, double depth_interval, int D, int h, int w, int H, int W,
                                            const float* grad_x_var, float* grad_feats_nhwc, void* stream) {
    DV3D_REQUIRE(C == 32, "planesweep_var_backward: C must be 32, got %d", C);
    DV3D_REQUIRE(feats_nhwc && cams && ref_img && edge_rowptr && edge_src && grad_x_var && grad_feats_nhwc,
                 "planesweep_var_backward: null pointer");
    DV3D_REQUIRE(n_imgs > 0 && Hf > 1 && Wf > 1 && D > 0 && h > 0 && w > 0 && H > 1 && W > 1 && n_ref >= 0 && n_ref <= 65535,
                 "planesweep_var_backward: bad shape");
    DV3D_REQUIRE((long long)Hf * Wf < (1 << 29), "planesweep_var_backward: feature map too large for the record encoding");
    if (n_ref == 0) return DV3D_OK;
    const int P = h * w, tiles = cdiv(P, TP), n_chunks = cdiv(D, KD);
    int dsplit = cdiv(4 * kNumSMs, (long long)tiles * n_ref);
    dsplit = dsplit < 1 ? 1 : (dsplit > n_chunks ? n_chunks : dsplit);
    const int chunks_per_cta = cdiv(n_chunks, dsplit);
    dim3 grid(tiles, n_ref, cdiv(n_chunks, chunks_per_cta));
    const double d1 = depth_start + depth_interval * (D - 1);
    const size_t smem = (sizeof(float4) + sizeof(int)) * EMAX * KD * TP + sizeof(float) * 32 * CS;
    static std::atomic<unsigned long long> attr{0};
    DV3D_FUNC_SMEM_ONCE(attr, (planesweep_var_bwd_kernel), (int)smem);
    DV3D_LAUNCH((planesweep_var_bwd_kernel), grid, 256, smem, (cudaStream_t)stream, reinterpret_cast<const float4*>(feats_nhwc), make_geom(Hf, Wf, H, W), cams, ref_img, edge_rowptr, edge_src, depth_start, d1, D, h, w, H, W, chunks_per_cta, grad_x_var, grad_feats_nhwc);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_points_var_backward(const float* feats_nhwc, int n_imgs, int C, int Hf, int Wf, const float* cams,
                                        const int* ref_img, const int* edge_rowptr, const int* edge_src,
                                        const float* depth, int n_ref, int h, int w, int H, int W, int n_side, double offset,
                                        const float* grad_feat, int rows_per_point, int feat_stride, int feat_off,
                                        float* grad_feats_nhwc, void* stream) {
    DV3D_REQUIRE(C == 32, "points_var_backward: C must be 32, got %d", C);
    DV3D_REQUIRE(n_side == 0 || n_side == 3, "points_var_backward: n_side must be 0 or 3, got %d", n_side);
    DV3D_REQUIRE(feats_nhwc && cams && ref_img && edge_rowptr && edge_src && depth && grad_feat && grad_feats_nhwc,
                 "points_var_backward: null pointer");
    DV3D_REQUIRE(rows_per_point >= 2 * n_side + 1 && feat_stride % 4 == 0 && feat_off % 4 == 0 && feat_off + C <= feat_stride &&
                     ((uintptr_t)grad_feat & 15) == 0,
                 "points_var_backward: bad gradient layout");
    DV3D_REQUIRE(n_imgs > 0 && Hf > 1 && Wf > 1 && h > 0 && w > 0 && n_ref >= 0 && n_ref <= 65535, "points_var_backward: bad shape");
    if (n_ref == 0) return DV3D_OK;
    dim3 grid(cdiv(h * w, TP), n_ref);
    DV3D_LAUNCH((points_var_bwd_kernel), grid, 256, 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(feats_nhwc), make_geom(Hf, Wf, H, W), cams, ref_img, edge_rowptr, edge_src, depth, h, w, H, W, n_side, offset, grad_feat, rows_per_point, feat_stride, feat_off, grad_feats_nhwc);
    DV3D_LAUNCHED();
    return DV3D_OK;
}
