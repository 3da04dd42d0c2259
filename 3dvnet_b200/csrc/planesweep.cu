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

#include "common.cuh"

namespace dv3d {

constexpr int TP = 32;     // pixels per CTA
constexpr int KD = 8;      // planes (or hypotheses) per chunk
constexpr int EMAX = 8;    // edges whose records are staged per pass
constexpr int CS = KD * TP + 1;  // padded channel stride of the output tile (bank-conflict free)

struct SampleGeom {
    int Hf, Wf;
    float wm1, hm1;    // W-1, H-1 of the FULL image (mvsnet.py:205-206)
    float wfm1, hfm1;  // Wf-1, Hf-1 (grid_sample align_corners=True un-normalisation)
};

// One (pixel, plane, edge) sample: q = z*a + b, z=|q_z|+1e-8, normalise, un-normalise, split
// into tap base + bilinear weights.  Taps outside the map get weight 0 and a clamped address
// (zero padding per tap, like ATen's grid_sampler_2d).
__device__ __forceinline__ void make_record(float z, float ax, float ay, float az, float bx, float by, float bz,
                                            const SampleGeom& g, int& rec, float4& wt) {
    float qx = fmaf(z, ax, bx), qy = fmaf(z, ay, by), qz = fmaf(z, az, bz);
    float zz = fabsf(qz) + 1e-8f;
    float x = __fdiv_rn(qx, zz), y = __fdiv_rn(qy, zz);
    float gx = __fsub_rn(__fmul_rn(__fdiv_rn(x, g.wm1), 2.f), 1.f);
    float gy = __fsub_rn(__fmul_rn(__fdiv_rn(y, g.hm1), 2.f), 1.f);
    float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), g.wfm1);
    float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), g.hfm1);
    rec = 0;
    wt = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ix > -1.f && ix < (float)g.Wf && iy > -1.f && iy < (float)g.Hf) {
        float fx0 = floorf(ix), fy0 = floorf(iy);
        float tx = ix - fx0, ty = iy - fy0;
        int x0 = (int)fx0, y0 = (int)fy0;
        bool l = x0 >= 0, r = x0 + 1 <= g.Wf - 1, t = y0 >= 0, b = y0 + 1 <= g.Hf - 1;
        float wl = l ? 1.f - tx : 0.f, wr = r ? tx : 0.f;
        float wtp = t ? 1.f - ty : 0.f, wb = b ? ty : 0.f;
        wt = make_float4(wl * wtp, wr * wtp, wl * wb, wr * wb);  // nw, ne, sw, se
        int cx = l ? x0 : 0, cy = t ? y0 : 0;
        int dx = (l && r) ? 1 : 0, dy = (t && b) ? 1 : 0;
        // when the left/top tap is the clamped one its weight is 0 and the right/bottom tap
        // must still address column x0+1 / row y0+1 == cx / cy: dx = dy = 0 does that.
        rec = ((cy * g.Wf + cx) << 2) | (dy << 1) | dx;
    }
}

__device__ __forceinline__ void fma4(float4& acc, float w, const float4& t) {
    acc.x = fmaf(w, t.x, acc.x);
    acc.y = fmaf(w, t.y, acc.y);
    acc.z = fmaf(w, t.z, acc.z);
    acc.w = fmaf(w, t.w, acc.w);
}

// Phase 2 for one staged pass of edges. NK = number of live planes/hypotheses in the chunk.
template <int NK>
__device__ __forceinline__ void consume_edges(const float4* __restrict__ feats, const int* __restrict__ esrc,
                                              int e_begin, int n_e, int img_stride4, int Wf, int v, int g,
                                              const int (*s_rec)[KD][TP], const float4 (*s_wt)[KD][TP],
                                              float4 (&acc_s)[KD], float4 (&acc_q)[KD]) {
    for (int e = 0; e < n_e; ++e) {
        const float4* base = feats + (size_t)__ldg(esrc + e_begin + e) * img_stride4 + g;
        int prev = -1;
        float4 t00 = make_float4(0, 0, 0, 0), t01 = t00, t10 = t00, t11 = t00;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            int rec = s_rec[e][k][v];
            float4 wt = s_wt[e][k][v];
            if (rec != prev) {
                const float4* p = base + (size_t)(rec >> 2) * 8;
                int dx = (rec & 1) * 8, dy = ((rec >> 1) & 1) * Wf * 8;
                t00 = ldg4(p);
                t01 = ldg4(p + dx);
                t10 = ldg4(p + dy);
                t11 = ldg4(p + dy + dx);
                prev = rec;
            }
            float4 val;
            val.x = wt.x * t00.x; val.y = wt.x * t00.y; val.z = wt.x * t00.z; val.w = wt.x * t00.w;
            fma4(val, wt.y, t01);
            fma4(val, wt.z, t10);
            fma4(val, wt.w, t11);
            acc_s[k].x += val.x; acc_s[k].y += val.y; acc_s[k].z += val.z; acc_s[k].w += val.w;
            acc_q[k].x = fmaf(val.x, val.x, acc_q[k].x);
            acc_q[k].y = fmaf(val.y, val.y, acc_q[k].y);
            acc_q[k].z = fmaf(val.z, val.z, acc_q[k].z);
            acc_q[k].w = fmaf(val.w, val.w, acc_q[k].w);
        }
    }
}

__device__ __forceinline__ float var_of(float s, float q, float n) {
    float m = __fdiv_rn(s, n);
    return __fsub_rn(__fdiv_rn(q, n), __fmul_rn(m, m));  // E[x^2] - E[x]^2 (mvsnet.py:216)
}

// numpy.linspace(start, stop, n, dtype=float32): float64 arithmetic, last point == stop
__device__ __forceinline__ float linspace_np(double start, double stop, int n, int i) {
    if (n == 1) return (float)start;
    if (i == n - 1) return (float)stop;
    double step = (stop - start) / (double)(n - 1);
    return (float)__dadd_rn(__dmul_rn((double)i, step), start);
}

__global__ void __launch_bounds__(256, 2)
planesweep_var_kernel(const float4* __restrict__ feats, SampleGeom geom, const float* __restrict__ xform,
                      const int* __restrict__ rowptr, const int* __restrict__ esrc, double d0, double d1, int D,
                      int h, int w, int H, int W, int chunks_per_cta, float* __restrict__ out) {
    // dynamic shared memory (73 KB > the 48 KB static limit): weights | records | output tile
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 (*s_wt)[KD][TP] = reinterpret_cast<float4 (*)[KD][TP]>(smem_raw);
    int (*s_rec)[KD][TP] = reinterpret_cast<int (*)[KD][TP]>(smem_raw + sizeof(float4) * EMAX * KD * TP);
    float* s_out = reinterpret_cast<float*>(smem_raw + (sizeof(float4) + sizeof(int)) * EMAX * KD * TP);

    const int tid = threadIdx.x;
    const int r = blockIdx.y;
    const int P = h * w;
    const int p0 = blockIdx.x * TP;
    const int e0 = rowptr[r], e1 = rowptr[r + 1];
    const float n_edges = (float)(e1 - e0);
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
        float4 acc_s[KD], acc_q[KD];
#pragma unroll
        for (int k = 0; k < KD; ++k) acc_s[k] = acc_q[k] = make_float4(0, 0, 0, 0);
        const float z = linspace_np(d0, d1, D, min(dbase + pk, D - 1));

        for (int eb = e0; eb < e1; eb += EMAX) {
            const int n_e = min(EMAX, e1 - eb);
            if (eb != e0) __syncthreads();  // previous pass fully consumed
            for (int e = 0; e < n_e; ++e) {
                const float* x = xform + (size_t)(eb + e) * 12;
                float ax = fmaf(__ldg(x + 0), u, fmaf(__ldg(x + 1), vv, __ldg(x + 2)));
                float ay = fmaf(__ldg(x + 3), u, fmaf(__ldg(x + 4), vv, __ldg(x + 5)));
                float az = fmaf(__ldg(x + 6), u, fmaf(__ldg(x + 7), vv, __ldg(x + 8)));
                int rec;
                float4 wt;
                make_record(z, ax, ay, az, __ldg(x + 9), __ldg(x + 10), __ldg(x + 11), geom, rec, wt);
                s_rec[e][pk][pv] = rec;
                s_wt[e][pk][pv] = wt;
            }
            __syncthreads();
            consume_edges<KD>(feats, esrc, eb, n_e, img_stride4, geom.Wf, v, g, s_rec, s_wt, acc_s, acc_q);
        }

        // phase 3: variance -> shared tile -> coalesced rows
#pragma unroll
        for (int k = 0; k < KD; ++k) {
            float* o = s_out + (4 * g) * CS + k * TP + v;
            o[0] = var_of(acc_s[k].x, acc_q[k].x, n_edges);
            o[CS] = var_of(acc_s[k].y, acc_q[k].y, n_edges);
            o[2 * CS] = var_of(acc_s[k].z, acc_q[k].z, n_edges);
            o[3 * CS] = var_of(acc_s[k].w, acc_q[k].w, n_edges);
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
        // the next chunk's first __syncthreads (after its records are written) orders these
        // reads of s_out before the next writes; records are rewritten only after every
        // thread has passed the barrier above, i.e. finished consuming them.
    }
}

// Point-level variant: the "planes" of a pixel are its 2n+1 depth hypotheses around the
// current depth estimate (lightningmodel.py:201-205); outputs are point-major.
__global__ void __launch_bounds__(256, 2)
points_var_kernel(const float4* __restrict__ feats, SampleGeom geom, const float* __restrict__ xform,
                  const int* __restrict__ rowptr, const int* __restrict__ esrc, const float* __restrict__ backproj,
                  const float* __restrict__ depth, int h, int w, int H, int W, int n_side, float offset,
                  float* __restrict__ pts_out, float* __restrict__ feat_out, int rows_per_point, int feat_stride,
                  int feat_off) {
    __shared__ int s_rec[EMAX][KD][TP];
    __shared__ float4 s_wt[EMAX][KD][TP];

    const int tid = threadIdx.x;
    const int r = blockIdx.y;
    const int P = h * w;
    const int p0 = blockIdx.x * TP;
    const int e0 = rowptr[r], e1 = rowptr[r + 1];
    const float n_edges = (float)(e1 - e0);
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
    const float z = __fadd_rn(dpt, (float)((double)(pk - n_side) * (double)offset));

    if (live && pk < n_hyp) {  // world point of hypothesis pk
        const float* B = backproj + (size_t)r * 12;
        float rx = fmaf(B[0], u, fmaf(B[1], vv, B[2]));
        float ry = fmaf(B[3], u, fmaf(B[4], vv, B[5]));
        float rz = fmaf(B[6], u, fmaf(B[7], vv, B[8]));
        float* o = pts_out + (((size_t)r * P + p) * n_hyp + pk) * 3;
        o[0] = fmaf(z, rx, B[9]);
        o[1] = fmaf(z, ry, B[10]);
        o[2] = fmaf(z, rz, B[11]);
    }

    float4 acc_s[KD], acc_q[KD];
#pragma unroll
    for (int k = 0; k < KD; ++k) acc_s[k] = acc_q[k] = make_float4(0, 0, 0, 0);

    for (int eb = e0; eb < e1; eb += EMAX) {
        const int n_e = min(EMAX, e1 - eb);
        if (eb != e0) __syncthreads();
        if (pk < n_hyp) {
            for (int e = 0; e < n_e; ++e) {
                const float* x = xform + (size_t)(eb + e) * 12;
                float ax = fmaf(__ldg(x + 0), u, fmaf(__ldg(x + 1), vv, __ldg(x + 2)));
                float ay = fmaf(__ldg(x + 3), u, fmaf(__ldg(x + 4), vv, __ldg(x + 5)));
                float az = fmaf(__ldg(x + 6), u, fmaf(__ldg(x + 7), vv, __ldg(x + 8)));
                int rec;
                float4 wt;
                make_record(z, ax, ay, az, __ldg(x + 9), __ldg(x + 10), __ldg(x + 11), geom, rec, wt);
                s_rec[e][pk][pv] = rec;
                s_wt[e][pk][pv] = wt;
            }
        }
        __syncthreads();
        if (n_hyp == 1)
            consume_edges<1>(feats, esrc, eb, n_e, img_stride4, geom.Wf, v, g, s_rec, s_wt, acc_s, acc_q);
        else
            consume_edges<7>(feats, esrc, eb, n_e, img_stride4, geom.Wf, v, g, s_rec, s_wt, acc_s, acc_q);
    }
    if (p0 + v < P) {
        const int p = p0 + v;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            if (k < n_hyp) {
                float4 o;
                o.x = var_of(acc_s[k].x, acc_q[k].x, n_edges);
                o.y = var_of(acc_s[k].y, acc_q[k].y, n_edges);
                o.z = var_of(acc_s[k].z, acc_q[k].z, n_edges);
                o.w = var_of(acc_s[k].w, acc_q[k].w, n_edges);
                *reinterpret_cast<float4*>(feat_out + (((size_t)r * P + p) * rows_per_point + k) * feat_stride +
                                           feat_off + 4 * g) = o;
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
__device__ void mm3(const double* a, const double* b, double* o) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) o[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}
__device__ void load3x3(const float* p, double* o) {
    for (int i = 0; i < 9; ++i) o[i] = (double)p[i];
}

__global__ void edge_transforms_kernel(const float* __restrict__ R, const float* __restrict__ t,
                                       const float* __restrict__ K, const int* __restrict__ eref,
                                       const int* __restrict__ esrc, int E, float* __restrict__ out) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int r = eref[e], s = esrc[e];
    double Rr[9], Rs[9], Kr[9], Ks[9], Kri[9], RrT[9], A[9], Bm[9], M[9];
    load3x3(R + 9 * r, Rr); load3x3(R + 9 * s, Rs); load3x3(K + 9 * r, Kr); load3x3(K + 9 * s, Ks);
    inv3(Kr, Kri);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) RrT[i * 3 + j] = Rr[j * 3 + i];
    mm3(Rs, RrT, A);    // R_s R_r^T
    mm3(Ks, A, Bm);     // K_s R_s R_r^T
    mm3(Bm, Kri, M);
    double tr[3] = {t[3 * r], t[3 * r + 1], t[3 * r + 2]}, ts[3] = {t[3 * s], t[3 * s + 1], t[3 * s + 2]};
    double tt[3];
    for (int i = 0; i < 3; ++i) tt[i] = ts[i] - (A[i * 3] * tr[0] + A[i * 3 + 1] * tr[1] + A[i * 3 + 2] * tr[2]);
    for (int i = 0; i < 9; ++i) out[e * 12 + i] = (float)M[i];
    for (int i = 0; i < 3; ++i) out[e * 12 + 9 + i] = (float)(Ks[i * 3] * tt[0] + Ks[i * 3 + 1] * tt[1] + Ks[i * 3 + 2] * tt[2]);
}

__global__ void ref_backprojection_kernel(const float* __restrict__ R, const float* __restrict__ t,
                                          const float* __restrict__ K, const int* __restrict__ ref_img, int n,
                                          float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = ref_img[i];
    double Rr[9], Kr[9], Kri[9], RrT[9], B[9];
    load3x3(R + 9 * r, Rr); load3x3(K + 9 * r, Kr);
    inv3(Kr, Kri);
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) RrT[a * 3 + b] = Rr[b * 3 + a];
    mm3(RrT, Kri, B);
    for (int a = 0; a < 9; ++a) out[i * 12 + a] = (float)B[a];
    for (int a = 0; a < 3; ++a)
        out[i * 12 + 9 + a] = (float)(-(RrT[a * 3] * t[3 * r] + RrT[a * 3 + 1] * t[3 * r + 1] + RrT[a * 3 + 2] * t[3 * r + 2]));
}

// NCHW -> NHWC through a padded 32x32 shared tile (both sides coalesced)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int HW) {
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
    g.wfm1 = (float)(Wf - 1); g.hfm1 = (float)(Hf - 1);
    return g;
}

}  // namespace dv3d

using namespace dv3d;

extern "C" int dv3d_nchw_to_nhwc(const float* src, float* dst, int n, int C, int HW, void* stream) {
    DV3D_REQUIRE(src && dst && n >= 0 && C > 0 && HW > 0, "nchw_to_nhwc: bad arguments");
    if (n == 0) return DV3D_OK;
    dim3 grid(cdiv(HW, 32), cdiv(C, 32), n), block(32, 8);
    nchw_to_nhwc_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, dst, C, HW);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_edge_transforms(const float* rotmats, const float* tvecs, const float* K, const int* edge_ref,
                                    const int* edge_src, int n_edges, float* xform_out, void* stream) {
    DV3D_REQUIRE(rotmats && tvecs && K && edge_ref && edge_src && xform_out && n_edges >= 0,
                 "edge_transforms: bad arguments");
    if (n_edges == 0) return DV3D_OK;
    edge_transforms_kernel<<<cdiv(n_edges, 64), 64, 0, (cudaStream_t)stream>>>(rotmats, tvecs, K, edge_ref, edge_src,
                                                                               n_edges, xform_out);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_ref_backprojection(const float* rotmats, const float* tvecs, const float* K, const int* ref_img,
                                       int n_ref, float* out, void* stream) {
    DV3D_REQUIRE(rotmats && tvecs && K && ref_img && out && n_ref >= 0, "ref_backprojection: bad arguments");
    if (n_ref == 0) return DV3D_OK;
    ref_backprojection_kernel<<<cdiv(n_ref, 64), 64, 0, (cudaStream_t)stream>>>(rotmats, tvecs, K, ref_img, n_ref, out);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_planesweep_var(const float* feats_nhwc, int n_imgs, int C, int Hf, int Wf, const float* xform,
                                   const int* edge_rowptr, const int* edge_src, int n_ref, double depth_start,
                                   double depth_interval, int D, int h, int w, int H, int W, float* x_var,
                                   void* stream) {
    DV3D_REQUIRE(C == 32, "planesweep_var: C must be 32 (IMG_FEAT_DIM, mv3d/config.py:42), got %d", C);
    DV3D_REQUIRE(feats_nhwc && xform && edge_rowptr && edge_src && x_var, "planesweep_var: null pointer");
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
    static bool attr_set = false;
    if (!attr_set) {
        DV3D_CUDA(cudaFuncSetAttribute(planesweep_var_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    planesweep_var_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(feats_nhwc), make_geom(Hf, Wf, H, W), xform, edge_rowptr, edge_src,
        depth_start, d1, D, h, w, H, W, chunks_per_cta, x_var);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_points_var(const float* feats_nhwc, int n_imgs, int C, int Hf, int Wf, const float* xform,
                               const int* edge_rowptr, const int* edge_src, const float* backproj,
                               const float* depth, int n_ref, int h, int w, int H, int W, int n_side, float offset,
                               float* pts_out, float* feat_out, int rows_per_point, int feat_stride, int feat_off,
                               void* stream) {
    DV3D_REQUIRE(C == 32, "points_var: C must be 32, got %d", C);
    DV3D_REQUIRE(n_side == 0 || n_side == 3, "points_var: n_side must be 0 (point cloud) or 3 (PointFlow), got %d",
                 n_side);
    DV3D_REQUIRE(feats_nhwc && xform && edge_rowptr && edge_src && backproj && depth && pts_out && feat_out,
                 "points_var: null pointer");
    DV3D_REQUIRE(rows_per_point >= 2 * n_side + 1, "points_var: rows_per_point < number of hypotheses");
    DV3D_REQUIRE(feat_stride % 4 == 0 && feat_off % 4 == 0 && feat_off + C <= feat_stride,
                 "points_var: feat_stride/feat_off must be multiples of 4 with room for C channels");
    DV3D_REQUIRE(n_imgs > 0 && Hf > 1 && Wf > 1 && h > 0 && w > 0 && n_ref >= 0 && n_ref <= 65535, "points_var: bad shape");
    if (n_ref == 0) return DV3D_OK;
    dim3 grid(cdiv(h * w, TP), n_ref);
    points_var_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(feats_nhwc), make_geom(Hf, Wf, H, W), xform, edge_rowptr, edge_src, backproj,
        depth, h, w, H, W, n_side, offset, pts_out, feat_out, rows_per_point, feat_stride, feat_off);
    DV3D_LAUNCHED();
    return DV3D_OK;
}
