// The PointFlow hypothesis decoder (refinement.py:17-25,42-44; lightningmodel.py:238-241) as ONE tcgen05 kernel:
//
//     Conv1d(352->128,k3)+BN+ReLU -> Conv1d(128->128,k3)+BN+ReLU -> Conv1d(128->128,k3)+BN+ReLU
//     -> Conv1d(128->1,k3)+bias -> softmax over the 7 hypotheses -> expected depth offset (+ depth += offset)
//
// A CTA owns a tile of 128 operand rows = 16 points x (7 hypotheses + 1 zero row).  The Conv1d taps never leave
// a point, so the whole stack of a tile runs without touching HBM between layers.
//
//   tap-stationary   out[m] = Y0[m-1] + Y1[m] + Y2[m+1] with Y_t = A W_t: ONE pass over A computes all three taps
//                    as N = 3 x 128 accumulator columns (TMEM columns [0,384), UMMA 128x256x8 + 128x128x8 per K
//                    step); the row shifts are two warp shuffles in the epilogue (a 32-lane TMEM quadrant is 4
//                    whole points, and the row before / after a point is a zero row, so nothing crosses warps).
//                    A is gathered, split and stored ONCE per layer instead of once per tap.
//   K pipeline       chunks of 16 floats (64-byte rows, SWIZZLE_64B).  Layer 1: 3 stages of
//                    [A big | A small | B big | B small] = 8+8+24+24 KB; the 8 producer warps gather the operand
//                    rows 6 chunks ahead in registers (zero rows are never read), split into TF32 big + fp32
//                    remainder and arrive ONCE per warp; the weights come as one bulk copy per stage.
//                    Layers 2, 3: the A operand (128 x 128, big + small = 128 KB) stays in shared memory, written
//                    by the previous layer's epilogue straight from TMEM (BN, ReLU, zero row, split); only the
//                    weights stream through a 2 x 48 KB ring.
//   precision        3xTF32 (A_big B_big + A_small B_big + A_big B_small, fp32-grade) or plain TF32.
//   head             the last epilogue keeps the 128 activations of a row in registers, forms the three tap dot
//                    products of the 128->1 convolution, and 16 threads finish softmax + expectation per point.
//
// Shared memory (bytes from the 1024-aligned base): layer 1 stages [0,192K); layers 2/3: A [0,128K), weight ring
// [128K,224K) (its first use waits for the last layer-1 MMAs of stage 2); head scratch aliases the ring.
#include <math.h>

#include "tc.cuh"

namespace dv3d {

constexpr int DF_BM = 128;            // rows per tile
constexpr int DF_H = 128;             // hidden channels
constexpr int DF_N = 3 * DF_H;        // tap-stationary accumulator columns
constexpr int DF_KC = 16;             // floats per K chunk (64-byte swizzle rows)
constexpr int DF_A_IMG = DF_BM * 64;  // 8 KB: one A chunk image (big or small)
constexpr int DF_B_IMG = DF_N * 64;   // 24 KB: one weight chunk image
constexpr int DF_STAGE1 = 2 * DF_A_IMG + 2 * DF_B_IMG;  // 64 KB
constexpr int DF_STAGES1 = 3;
constexpr int DF_A2_BYTES = (DF_H / DF_KC) * 2 * DF_A_IMG;  // 128 KB
constexpr int DF_RING_OFF = DF_A2_BYTES;
constexpr int DF_NH = DF_N / 2;              // 192 accumulator columns per channel half (3 taps x 64 channels)
constexpr int DF_BH_IMG = DF_NH * 64;        // 12 KB: the weight rows of one channel half of a chunk image
constexpr int DF_RING_STAGE = 2 * DF_BH_IMG;  // 24 KB: big + small rows of one (chunk, half)
constexpr int DF_RING = 4;
constexpr int DF_DATA_BYTES = DF_RING_OFF + DF_RING * DF_RING_STAGE;  // 224 KB
constexpr int DF_PRODUCERS = 256;
constexpr int DF_THREADS = DF_PRODUCERS + 64;
constexpr int DF_PREFETCH = 6;
constexpr int DF_TMEM_COLS = 512;
constexpr size_t DF_SMEM = 1024 + (size_t)DF_DATA_BYTES + 256;
static_assert(DF_STAGES1 * DF_STAGE1 <= DF_DATA_BYTES, "layer-1 stages must fit the data region");
static_assert(DF_SMEM <= 227 * 1024, "shared memory budget");

struct DecoderFusedArgs {
    const float* x;        // [n_pts, 8, ld]
    long long M;           // n_pts * 8
    int ld, K1;            // row pitch, input channels (multiple of 16)
    const float* Wp[3];    // packed by dv3d_decoder_pack_weights
    const float* scale[3]; // folded BatchNorm
    const float* shift[3];
    const float* head_w;   // [1, 128, 3]
    float head_b;
    float lo, hi;          // offsets of the first / last hypothesis
    float* prob_out;       // [n_pts, 7] or nullptr
    float* offset_out;     // [n_pts] or nullptr
    float* depth_accum;    // [n_pts] or nullptr: depth += expected offset (eval-3dvnet.py:99)
    int precision;         // 1 = 3xTF32, 2 = TF32
    long long* timing;     // tools/decoder_phases.py: [n_tiles][8] clock64 stamps of thread 0 / the epilogue, or nullptr
};

// byte offset of 16-byte unit j of row r inside a K-major SWIZZLE_64B chunk image
__device__ __forceinline__ int sw64_off(int r, int j) { return (r >> 3) * 512 + (r & 7) * 64 + ((j ^ ((r & 7) >> 1)) << 4); }

// weight_tkn [3, Cin, 128] -> per 16-row K chunk the shared-memory image of B (row n = half * 192 + tap * 64 +
// channel % 64, K-major SWIZZLE_64B), TF32 big parts then fp32 remainders
__global__ void __launch_bounds__(256)
pack_decoder_weights_kernel(const float* __restrict__ W, int Cin, float* __restrict__ out) {
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3ll * Cin * DF_H) return;
    const int co = (int)(i % DF_H);
    const int k = (int)((i / DF_H) % Cin);
    const int t = (int)(i / ((long long)DF_H * Cin));
    float big, small;
    split_tf32(__ldg(W + i), big, small);
    // row of the B operand: channel half (co >> 6) outermost, then tap, then channel within the half - the two
    // halves are separate N = 192 accumulators whose epilogues overlap the other half's MMAs
    const int n = (co >> 6) * DF_NH + t * 64 + (co & 63), chunk = k / DF_KC, kk = k % DF_KC;
    const size_t img = DF_B_IMG / 4;  // floats per image
    const size_t pos = (size_t)(sw64_off(n, kk >> 2) >> 2) + (kk & 3);
    out[(size_t)chunk * 2 * img + pos] = big;
    out[(size_t)chunk * 2 * img + img + pos] = small;
}

// one K step (8 tf32) of 3xTF32 (or TF32) into ONE half accumulator: D[:, col .. col + 192) += A (big|small) x B rows
__device__ __forceinline__ void df_issue_half(uint32_t tmem_col, uint32_t a_big, uint32_t a_small, uint32_t b_big,
                                              uint32_t b_small, int precision, uint32_t acc) {
    constexpr uint32_t id = umma_idesc_tf32(DF_BM, DF_NH);
    umma_tf32(tmem_col, umma_desc_sw64(a_big), umma_desc_sw64(b_big), id, acc);
    if (precision == 1) {
        umma_tf32(tmem_col, umma_desc_sw64(a_small), umma_desc_sw64(b_big), id, 1);
        umma_tf32(tmem_col, umma_desc_sw64(a_big), umma_desc_sw64(b_small), id, 1);
    }
}

// Epilogue of one channel HALF (64 channels x 3 taps = 192 accumulator columns) of a layer: all 8 warps take part,
// warp w reads TMEM lanes 32 (w & 3) .. +31 (rows) and 32 of the half's 64 channels (two blocks of 16).
struct EpiCtx {
    uint32_t trow;      // TMEM address of the warp's lane quadrant
    int sub;            // which 32 channels of the half
    bool pad_row, first_row, last_rows;
};
__device__ __forceinline__ void epi_load(const EpiCtx& e, int h, int cb, uint32_t (&y0)[16], uint32_t (&y1)[16],
                                         uint32_t (&y2)[16]) {
    const uint32_t col = (uint32_t)(DF_NH * h + 32 * e.sub + 16 * cb);   // tap 0; taps are 64 columns apart
    tmem_ld16_nowait(e.trow + col, y0);
    tmem_ld16_nowait(e.trow + col + 64, y1);
    tmem_ld16_nowait(e.trow + col + 128, y2);
}
// tap shifts -> BN + ReLU for the 16 channels starting at c0
__device__ __forceinline__ void epi_combine(const EpiCtx& e, int c0, const uint32_t (&y0)[16], const uint32_t (&y1)[16],
                                            const uint32_t (&y2)[16], const float* __restrict__ scale,
                                            const float* __restrict__ shift, float (&v)[16]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c0) + j);
        const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + c0) + j);
        const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = 4 * j + i;
            // out[m] = Y0[m-1] + Y1[m] + Y2[m+1]; rows outside the point are zero rows
            float up = __shfl_up_sync(0xffffffffu, __uint_as_float(y0[k]), 1);
            float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(y2[k]), 1);
            if (e.first_row) up = 0.f;
            if (e.last_rows) dn = 0.f;
            float s = (up + __uint_as_float(y1[k])) + dn;
            s = fmaxf(fmaf(s, scv[i], shv[i]), 0.f);
            v[k] = e.pad_row ? 0.f : s;
        }
    }
}

// PERSISTENT: grid = min(tiles, 148) CTAs, CTA b owns tiles b, b + grid, ...  The mbarrier rings keep running
// across tiles (stage / phase from a running chunk index), so nothing is re-initialised between tiles, the first
// operand rows of the next tile are gathered while the head of the previous one finishes, and the per-wave cost of
// bringing up a 225 KB-shared-memory CTA (measured: ~10 us per wave) is paid once per launch.
__global__ void __launch_bounds__(DF_THREADS, 1)
decoder_fused_kernel(const __grid_constant__ DecoderFusedArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + DF_DATA_BYTES);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 18);
    const uint32_t bar0 = smem_u32(s_bar);
    // full[3] empty[3] | ring full[4] empty[4] | accumulator-complete [2 halves] | operand-ready / drained [2 halves]
    const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * 3, bar_rfull = bar0 + 8 * 6, bar_rempty = bar0 + 8 * 10,
                   bar_accum = bar0 + 8 * 14, bar_aready = bar0 + 8 * 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n1 = a.K1 / DF_KC;              // chunks of layer 1
    constexpr int n2 = DF_H / DF_KC;          // chunks of layers 2 and 3
    const bool x3 = a.precision == 1;
    const uint32_t w_bytes = x3 ? 2 * DF_B_IMG : DF_B_IMG;
    const long long n_tiles = (a.M + DF_BM - 1) / DF_BM;

    if (tid == 0) {
        for (int s = 0; s < DF_STAGES1; ++s) {
            mbar_init(bar_full + 8 * s, DF_PRODUCERS / 32 + 1);  // one arrive per producer warp + the copy thread
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int s = 0; s < DF_RING; ++s) {
            mbar_init(bar_rfull + 8 * s, 1);
            mbar_init(bar_rempty + 8 * s, 1);
        }
        for (int hh = 0; hh < 2; ++hh) {
            mbar_init(bar_accum + 8 * hh, 1);
            mbar_init(bar_aready + 8 * hh, DF_PRODUCERS / 32);
        }
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(smem_u32(s_tmem), DF_TMEM_COLS);
    pdl_wait();  // everything above is independent of the previous kernel in the stream
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;

    // j = local tile counter.  Running indices keep every mbarrier's phase consistent across tiles: layer-1 ring
    // chunk g1 = j n1 + it; weight-ring chunk g2 = 4 n2 j + i (layers 2/3: 2 layers x 2 halves x n2 chunks per tile);
    // accumulator-complete and operand-ready barriers of each half complete 3 times per tile (3 j + layer).
    //
    // Layers 2 and 3 run HALF-MAJOR: all K chunks of channel half 0 (N = 192 accumulator columns), then half 1, so
    // the epilogue of half 0 (TMEM -> BN/ReLU -> next operand) overlaps the MMAs of half 1, and the next layer's
    // first chunks (which need only the operand columns half 0 produced) overlap the epilogue of half 1.
    if (warp == 8) {
        // ===================== MMA issuer: the whole warp walks the pipeline (uniform control flow keeps the
        // descriptors in uniform registers), one elected lane issues
        int j = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
            // the last epilogues of the previous tile have drained both half accumulators
            if (j > 0) {
                mbar_wait(bar_aready, (uint32_t)(3 * (j - 1) + 2) & 1u);
                mbar_wait(bar_aready + 8, (uint32_t)(3 * (j - 1) + 2) & 1u);
                tc_fence_after();
            }
            for (int it = 0; it < n1; ++it) {
                const int g = j * n1 + it, st = g % DF_STAGES1;
                mbar_wait(bar_full + 8 * st, (uint32_t)(g / DF_STAGES1) & 1u);
                tc_fence_after();
                const uint32_t a_big = smem_u32(smem + st * DF_STAGE1), a_small = a_big + DF_A_IMG,
                               b_big = a_big + 2 * DF_A_IMG, b_small = b_big + DF_B_IMG;
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < DF_KC / 8; ++kk) {
                        const uint32_t acc = (it | kk) ? 1u : 0u, ko = kk * 32;
                        // chunk-major in layer 1 (the operand streams by): both halves per K step, interleaved
                        df_issue_half(tmem, a_big + ko, a_small + ko, b_big + ko, b_small + ko, a.precision, acc);
                        df_issue_half(tmem + DF_NH, a_big + ko, a_small + ko, b_big + DF_BH_IMG + ko, b_small + DF_BH_IMG + ko,
                                      a.precision, acc);
                    }
                    umma_commit(bar_empty + 8 * st);
                }
                __syncwarp();
            }
            if (elect_one()) {
                umma_commit(bar_accum);
                umma_commit(bar_accum + 8);
            }
            __syncwarp();
            for (int layer = 1; layer < 3; ++layer) {
                for (int hh = 0; hh < 2; ++hh) {
                    for (int c = 0; c < n2; ++c) {
                        if (hh == 0 && (c == 0 || c == n2 / 2)) {
                            // operand chunks [0, n2/2) come from the previous layer's half-0 epilogue (which also
                            // drained accumulator half 0), chunks [n2/2, n2) and accumulator half 1 from half 1's
                            mbar_wait(bar_aready + 8 * (c ? 1 : 0), (uint32_t)(3 * j + layer - 1) & 1u);
                            tc_fence_after();
                        }
                        const int g = ((2 * j + layer - 1) * 2 + hh) * n2 + c, st = g % DF_RING;
                        mbar_wait(bar_rfull + 8 * st, (uint32_t)(g / DF_RING) & 1u);
                        tc_fence_after();
                        const uint32_t a_big = smem_u32(smem + c * 2 * DF_A_IMG), a_small = a_big + DF_A_IMG,
                                       b_big = smem_u32(smem + DF_RING_OFF + st * DF_RING_STAGE), b_small = b_big + DF_BH_IMG;
                        if (elect_one()) {
#pragma unroll
                            for (int kk = 0; kk < DF_KC / 8; ++kk)
                                df_issue_half(tmem + DF_NH * hh, a_big + kk * 32, a_small + kk * 32, b_big + kk * 32,
                                              b_small + kk * 32, a.precision, (c | kk) ? 1u : 0u);
                            umma_commit(bar_rempty + 8 * st);
                        }
                        __syncwarp();
                    }
                    if (elect_one()) umma_commit(bar_accum + 8 * hh);
                    __syncwarp();
                }
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            // ===================== weight copies: one bulk copy per chunk as soon as its stage is free
            int j = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
                // the stages of layer 1 overlap the A operand and the weight ring of layers 2/3: all MMAs of the
                // previous tile must have completed (its third accumulator completion)
                // (a parity wait is only unambiguous once the previous completion has happened: this thread gets here
                // after its last ring copy of the previous tile, which waited for layer-3 MMAs, i.e. past completion
                // 3 (j-1) + 1)
                if (j > 0) mbar_wait(bar_accum + 8, (uint32_t)(3 * (j - 1) + 2) & 1u);   // half 1 of layer 3 finishes last
                int g_last2 = -1;  // last chunk of this tile that lives in physical stage 2
                for (int it = 0; it < n1; ++it) {
                    const int g = j * n1 + it, st = g % DF_STAGES1;
                    if (st == 2) g_last2 = g;
                    mbar_wait(bar_empty + 8 * st, ((uint32_t)(g / DF_STAGES1) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(bar_full + 8 * st, w_bytes);
                    bulk_g2s(smem_u32(smem + st * DF_STAGE1 + 2 * DF_A_IMG), a.Wp[0] + (size_t)it * (2 * DF_B_IMG / 4), w_bytes,
                             bar_full + 8 * st);
                }
                // the ring of layers 2/3 overlaps physical stage 2 (and nothing else that is still live): wait for
                // the MMAs of the LAST chunk that used it, as if one more chunk were to be loaded into it
                if (g_last2 >= 0) mbar_wait(bar_empty + 8 * 2, ((uint32_t)((g_last2 + DF_STAGES1) / DF_STAGES1) & 1u) ^ 1u);
                for (int i = 0; i < 4 * n2; ++i) {
                    const int g = 4 * n2 * j + i, st = g % DF_RING, layer = i / (2 * n2), hh = (i / n2) & 1, c = i % n2;
                    mbar_wait(bar_rempty + 8 * st, ((uint32_t)(g / DF_RING) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(bar_rfull + 8 * st, x3 ? 2 * DF_BH_IMG : DF_BH_IMG);
                    // the rows of half hh are contiguous inside the big and inside the small image of the chunk
                    const float* src = a.Wp[1 + layer] + (size_t)c * (2 * DF_B_IMG / 4) + (size_t)hh * (DF_BH_IMG / 4);
                    const uint32_t dst = smem_u32(smem + DF_RING_OFF + st * DF_RING_STAGE);
                    bulk_g2s(dst, src, DF_BH_IMG, bar_rfull + 8 * st);
                    if (x3) bulk_g2s(dst + DF_BH_IMG, src + DF_B_IMG / 4, DF_BH_IMG, bar_rfull + 8 * st);
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== warps 0..7: layer-1 producers, then the epilogues of all three layers
        const int pj = tid & 3;    // 16-byte unit of the 64-byte chunk row
        const int r0 = tid >> 2;   // rows r0 and r0 + 64
        EpiCtx ec;
        ec.sub = warp >> 2;
        ec.trow = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        ec.pad_row = (lane & 7) == 7;
        ec.first_row = (lane & 7) == 0;
        ec.last_rows = (lane & 7) >= 6;
        const int row = 32 * (warp & 3) + lane;
        // head scratch [2 halves][3 taps][128 rows] at the very end of the data region: only the weight ring's
        // second stage lives there, and it is idle from the last MMA of a tile to layer 2 of the next
        float* s_head = reinterpret_cast<float*>(smem + DF_DATA_BYTES - 6 * DF_BM * sizeof(float));
        int j = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
            const long long m0 = tile * DF_BM;
            long long* stamp = a.timing ? a.timing + (size_t)tile * 8 : nullptr;
            if (stamp && tid == 0) stamp[0] = stamp[1] = clock64();
            {
                const float* src[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int r = r0 + 64 * i;
                    const long long m = m0 + r;
                    // the 8th row of a point is the Conv1d zero padding: never read (and never trusted to be zero)
                    src[i] = (m < a.M && (r & 7) != 7) ? a.x + (size_t)m * a.ld + pj * 4 : nullptr;
                }
                float4 buf[DF_PREFETCH][2];
#pragma unroll
                for (int p = 0; p < DF_PREFETCH; ++p)
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        buf[p][i] = (p < n1 && src[i]) ? __ldg(reinterpret_cast<const float4*>(src[i] + p * DF_KC))
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
                for (int it0 = 0; it0 < n1; it0 += DF_PREFETCH) {
#pragma unroll
                    for (int p = 0; p < DF_PREFETCH; ++p) {
                        const int it = it0 + p;
                        if (it < n1) {
                            const int g = j * n1 + it, st = g % DF_STAGES1;
                            mbar_wait(bar_empty + 8 * st, ((uint32_t)(g / DF_STAGES1) & 1u) ^ 1u);
                            unsigned char* stage = smem + st * DF_STAGE1;
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                const float4 v = buf[p][i];
                                float4 big, small;
                                split_tf32(v.x, big.x, small.x);
                                split_tf32(v.y, big.y, small.y);
                                split_tf32(v.z, big.z, small.z);
                                split_tf32(v.w, big.w, small.w);
                                const int off = sw64_off(r0 + 64 * i, pj);
                                *reinterpret_cast<float4*>(stage + off) = big;
                                if (x3) *reinterpret_cast<float4*>(stage + DF_A_IMG + off) = small;
                            }
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(bar_full + 8 * st);
                            if (it + DF_PREFETCH < n1) {
#pragma unroll
                                for (int i = 0; i < 2; ++i)
                                    if (src[i]) buf[p][i] = __ldg(reinterpret_cast<const float4*>(src[i] + (it + DF_PREFETCH) * DF_KC));
                            }
                        }
                    }
                }
            }
            // ---- epilogues, one channel half at a time (see the MMA issuer): TMEM -> tap shifts -> BN + ReLU ->
            // next layer's A operand (layers 1, 2) or the head's tap dot products (layer 3)
            float dot[3] = {0.f, 0.f, 0.f};  // head: tap dot products of this row over this thread's 2 x 32 channels
            for (int layer = 0; layer < 3; ++layer) {
                const float* scale = a.scale[layer];
                const float* shift = a.shift[layer];
                for (int hh = 0; hh < 2; ++hh) {
                    mbar_wait(bar_accum + 8 * hh, (uint32_t)(3 * j + layer) & 1u);
                    tc_fence_after();
                    if (stamp && tid == 0 && hh == 0) stamp[2 + 2 * layer] = clock64();   // half 0 of this layer complete
                    uint32_t ya[3][16], yb[3][16];
                    epi_load(ec, hh, 0, ya[0], ya[1], ya[2]);
                    tmem_ld_wait();
                    epi_load(ec, hh, 1, yb[0], yb[1], yb[2]);
                    const int c00 = 64 * hh + 32 * ec.sub;   // first of this thread's 32 channels
                    float v[2][16];
                    epi_combine(ec, c00, ya[0], ya[1], ya[2], scale, shift, v[0]);
                    tmem_ld_wait();
                    epi_combine(ec, c00 + 16, yb[0], yb[1], yb[2], scale, shift, v[1]);
                    if (layer < 2) {
                        // The next layer's A operand overwrites THIS layer's in place.  In the half-major layers half 1's
                        // MMAs still read it while half 0's epilogue runs (that overlap is the point), so half 0 keeps
                        // its 32 values in registers and stores them once accumulator half 1 is complete.
                        if (layer > 0 && hh == 0) mbar_wait(bar_accum + 8, (uint32_t)(3 * j + layer) & 1u);
#pragma unroll
                        for (int cb = 0; cb < 2; ++cb) {
                            // chunk (c00 + 16 cb) / 16, K-major SWIZZLE_64B, big + small images
                            unsigned char* img = smem + (size_t)((c00 + 16 * cb) / DF_KC) * 2 * DF_A_IMG;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                float4 big, small;
                                split_tf32(v[cb][4 * q], big.x, small.x);
                                split_tf32(v[cb][4 * q + 1], big.y, small.y);
                                split_tf32(v[cb][4 * q + 2], big.z, small.z);
                                split_tf32(v[cb][4 * q + 3], big.w, small.w);
                                const int off = sw64_off(row, q);
                                *reinterpret_cast<float4*>(img + off) = big;
                                if (x3) *reinterpret_cast<float4*>(img + DF_A_IMG + off) = small;
                            }
                        }
                    } else {
#pragma unroll
                        for (int cb = 0; cb < 2; ++cb) {
                            const float4* w4 = reinterpret_cast<const float4*>(a.head_w + 3 * (c00 + 16 * cb));  // [16 ch][3 taps]
                            float wv[48];
#pragma unroll
                            for (int q = 0; q < 12; ++q) {
                                const float4 t = __ldg(w4 + q);
                                wv[4 * q] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
                            }
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                dot[0] = fmaf(wv[3 * i], v[cb][i], dot[0]);
                                dot[1] = fmaf(wv[3 * i + 1], v[cb][i], dot[1]);
                                dot[2] = fmaf(wv[3 * i + 2], v[cb][i], dot[2]);
                            }
                        }
                    }
                    if (layer < 2) fence_proxy_async_smem();
                    tc_fence_before();
                    __syncwarp();
                    // layers 1, 2: operand chunks of this half stored; always: this half's accumulator drained
                    if (lane == 0) mbar_arrive(bar_aready + 8 * hh);
                    if (stamp && tid == 0 && hh == 1) stamp[3 + 2 * layer] = clock64();   // both halves' epilogues done
                }
            }
            // ---- head: logits of the 128 -> 1 convolution, softmax over the hypotheses, expected offset
#pragma unroll
            for (int t = 0; t < 3; ++t) s_head[(ec.sub * 3 + t) * DF_BM + row] = dot[t];
            asm volatile("bar.sync 1, %0;" ::"n"(DF_PRODUCERS) : "memory");
            if (tid < DF_BM / 8) {
                const long long p = m0 / 8 + tid;
                if (p * 8 < a.M) {
                    const int rb = tid * 8;
                    float logit[7];
#pragma unroll
                    for (int h = 0; h < 7; ++h) {
                        const int r = rb + h;
                        float l = s_head[1 * DF_BM + r] + s_head[4 * DF_BM + r];
                        if (h > 0) l += s_head[0 * DF_BM + r - 1] + s_head[3 * DF_BM + r - 1];
                        if (h < 6) l += s_head[2 * DF_BM + r + 1] + s_head[5 * DF_BM + r + 1];
                        logit[h] = l + a.head_b;
                    }
                    float mx = logit[0];
#pragma unroll
                    for (int h = 1; h < 7; ++h) mx = fmaxf(mx, logit[h]);
                    float e[7], s = 0.f;
#pragma unroll
                    for (int h = 0; h < 7; ++h) {
                        e[h] = expf(logit[h] - mx);
                        s += e[h];
                    }
                    float acc = 0.f;
#pragma unroll
                    for (int h = 0; h < 7; ++h) {
                        const float pr = e[h] / s;
                        if (a.prob_out) a.prob_out[p * 7 + h] = pr;
                        acc += linspace_torch(a.lo, a.hi, 7, h) * pr;
                    }
                    if (a.offset_out) a.offset_out[p] = acc;
                    if (a.depth_accum) a.depth_accum[p] = a.depth_accum[p] + acc;
                }
            }
            // (s_head is rewritten by the next tile's head three accumulator waits later: no second barrier needed)
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, DF_TMEM_COLS);
}

}  // namespace dv3d

using namespace dv3d;

// profiling aid (tools/decoder_phases.py): when set, every launch writes 8 clock64 stamps per tile into this device
// buffer (kernel entry, after the dependency wait, then accumulator-ready / epilogue-done of the three layers)
static std::atomic<long long*> g_decoder_timing{nullptr};
extern "C" int dv3d_decoder_set_timing_buffer(void* device_buffer) {
    g_decoder_timing.store((long long*)device_buffer);
    return DV3D_OK;
}

extern "C" size_t dv3d_decoder_pack_bytes(int Cin) {
    if (Cin <= 0 || Cin % DF_KC) return 0;
    return (size_t)(Cin / DF_KC) * 2 * DF_B_IMG;
}

extern "C" int dv3d_decoder_pack_weights(const float* weight_tkn, int Cin, int Cout, void* packed, void* stream) {
    DV3D_REQUIRE(weight_tkn && packed && Cin > 0 && Cin % DF_KC == 0 && Cout == DF_H,
                 "decoder_pack_weights: need Cin %% 16 == 0 and Cout == 128 (Cin=%d Cout=%d)", Cin, Cout);
    DV3D_LAUNCH((pack_decoder_weights_kernel), cdiv(3ll * Cin * DF_H, 256), 256, 0, (cudaStream_t)stream, weight_tkn, Cin,
                (float*)packed);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_decoder_fused(const float* x, long long n_pts, int rows_per_point, int Cin, int ldx,
                                  const void* const* W_packed, const float* const* scale, const float* const* shift,
                                  int hidden, const float* head_weight, float head_bias, double offset, int precision,
                                  float* prob_out, float* offset_out, float* depth_accum, void* stream) {
    DV3D_REQUIRE(x && W_packed && scale && shift && head_weight && (offset_out || depth_accum || prob_out) && n_pts >= 0,
                 "decoder_fused: bad arguments");
    DV3D_REQUIRE(rows_per_point == 8, "decoder_fused: the operand layout is [n_pts, 8, C] (7 hypotheses + 1 zero row)");
    DV3D_REQUIRE(hidden == DF_H && Cin > 0 && Cin % DF_KC == 0 && ldx >= Cin && ldx % 4 == 0 && ((uintptr_t)x & 15) == 0,
                 "decoder_fused: hidden must be 128, Cin a multiple of 16, rows 16-byte aligned (Cin=%d hidden=%d ld=%d)", Cin,
                 hidden, ldx);
    DV3D_REQUIRE(precision == 1 || precision == 2, "decoder_fused: precision 1 = 3xTF32, 2 = TF32; got %d", precision);
    if (n_pts == 0) return DV3D_OK;
    DecoderFusedArgs a = {};
    a.x = x;
    a.M = n_pts * 8;
    a.ld = ldx;
    a.K1 = Cin;
    for (int l = 0; l < 3; ++l) {
        DV3D_REQUIRE(W_packed[l] && scale[l] && shift[l] && ((uintptr_t)W_packed[l] & 15) == 0, "decoder_fused: layer %d parameters", l);
        a.Wp[l] = (const float*)W_packed[l];
        a.scale[l] = scale[l];
        a.shift[l] = shift[l];
    }
    a.head_w = head_weight;
    a.head_b = head_bias;
    a.lo = (float)(-3.0 * offset);   // linspace(-n*offset, n*offset, 2n+1), n = 3 (lightningmodel.py:238)
    a.hi = (float)(3.0 * offset);
    a.prob_out = prob_out;
    a.offset_out = offset_out;
    a.depth_accum = depth_accum;
    a.precision = precision;
    a.timing = g_decoder_timing.load();
    static std::atomic<unsigned long long> attr{0};
    DV3D_FUNC_SMEM_ONCE(attr, (decoder_fused_kernel), (int)DF_SMEM);
    const int tiles = cdiv(a.M, DF_BM);
    DV3D_LAUNCH((decoder_fused_kernel), tiles < kNumSMs ? tiles : kNumSMs, DF_THREADS, DF_SMEM, (cudaStream_t)stream, a);
    DV3D_LAUNCHED();
    return DV3D_OK;
}
