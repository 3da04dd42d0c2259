// Gather-GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), see gemm.cuh for the
// contraction it implements.
//
//   CTA tile        128 output rows x N (64 | 128) channels, accumulator in TMEM
//                   (128 lanes x N fp32 columns), UMMA shape M128 x N x K8, kind::tf32
//   K pipeline      chunks of 32 floats (one 128-byte swizzle row per operand row), 3 stages.
//                   A (gathered feature rows): 8 producer warps read the rows named by the
//                   slice's row table / shift with 16-byte loads FOUR chunks ahead (the register
//                   ring covers the L2 / HBM latency of a gather), split every value into a
//                   TF32 "big" part and an fp32 remainder, and store both in the canonical
//                   K-major SWIZZLE_128B layout (st.shared + fence.proxy.async).
//                   B (weights): pre-packed once per layer (pack_weights_kernel) into the exact
//                   shared-memory image of every chunk, big and small part, and brought in by one
//                   bulk-copy (cp.async.bulk -> mbarrier complete_tx) per stage, issued by a
//                   dedicated copy thread (warp 9) as soon as the stage is free - the first
//                   copies are in flight while the producers still wait for their first rows.
//   precision       3xTF32: acc += A_big B_big + A_big B_small + A_small B_big  (fp32-grade,
//                   error ~2^-21 per product), or plain TF32 (first term only) as an opt-in.
//   MMA issue       one elected thread of warp 8; tcgen05.commit releases the stage / signals
//                   the epilogue through mbarriers - no __syncthreads in the main loop.
//   epilogue        warps 0-3 move the accumulator tile TMEM -> shared memory (tcgen05.ld, thread =
//                   row); then ALL warps run the epilogue on 16-byte units (thread = row x 4
//                   channels, GroupNorm statistics by shuffles over the 4 lanes of a 16-channel
//                   group): folded BN / bias, per-row GroupNorm, residual, ReLU - residual reads
//                   and output stores are fully coalesced.
//   sparsity        kernel-map slices with no live row in the tile are skipped by producers and
//                   issuer alike (flags computed from the staged row table).
#include "gemm.cuh"
#include <stdlib.h>

#include "tc.cuh"

namespace dv3d {

constexpr int TC_BM = 128;
constexpr int TC_KC = 32;                  // floats per K chunk = 128 bytes per operand row
constexpr int TC_STAGES = 3;
constexpr int TC_PRODUCERS = 256;          // 8 warps
constexpr int TC_THREADS = TC_PRODUCERS + 64;  // + warp 8: MMA issuer, warp 9: weight copies
constexpr int TC_PREFETCH = 4;             // chunks of A rows in flight per producer thread
constexpr int TC_MAX_SPLIT = 9;
constexpr int TC_A_BYTES = TC_BM * 128;    // one operand image (big or small part)

__host__ __device__ constexpr int tc_stage_bytes(int N) { return 2 * TC_A_BYTES + 2 * N * 128; }
__host__ __device__ constexpr size_t tc_smem_bytes(int N) {
    return 1024 /* alignment slack */ + (size_t)TC_STAGES * tc_stage_bytes(N) + sizeof(int) * kMaxSlices * TC_BM + 256;
}

static int g_gemm_precision = 1;  // 1 = 3xTF32, 2 = TF32
static int g_pair_ws = -1;         // pair-major GEMM: 1 = weight-stationary kernel (default), 0 = the general kernels (DV3D_PAIR_WS=0)
static int g_gemm_persistent = 0;  // 0 = automatic (more tiles than SMs), 1 = whenever possible, -1 = never (tests / A-B runs)
// profiling aid (tools/gemm_phases.py): 8 clock64 stamps per CTA of every launch, or nullptr
__device__ long long* g_gemm_stamps = nullptr;

// ------------------------------------------------------------------ weight packing
// W [Ktot, N] row-major -> per 32-row chunk the shared-memory image of B as a K-major
// SWIZZLE_128B operand: row n (output channel) holds k = 0..31 in 16-byte units, unit j
// stored at j ^ (n & 7).  Image of the big parts first, then of the remainders.
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ W, int Ktot, int N, float* __restrict__ out) {
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)Ktot * N) return;
    const int k = (int)(i / N), n = (int)(i - (long long)k * N);
    const int chunk = k >> 5, kk = k & 31;
    float big, small;
    split_tf32(__ldg(W + i), big, small);
    const size_t img = (size_t)N * 32;  // floats per image
    const size_t pos = (size_t)n * 32 + ((((kk >> 2) ^ (n & 7)) << 2) | (kk & 3));
    out[(size_t)chunk * 2 * img + pos] = big;
    out[(size_t)chunk * 2 * img + img + pos] = small;
}

// ------------------------------------------------------------------ the kernel
// Position of a producer in the CTA's sequence of K chunks (live slices of its split range).
struct ChunkCursor {
    int s, kc, nk, wchunk;
    const float* src[4];
};

__device__ __forceinline__ void cursor_enter(ChunkCursor& c, const GemmDesc& d, int s_end, uint32_t active,
                                             const int* s_rows, int r0, int j) {
    while (c.s < s_end && !((active >> c.s) & 1u)) {
        c.wchunk += d.slice[c.s].K / TC_KC;
        ++c.s;
    }
    c.kc = 0;
    if (c.s < s_end) {
        const GemmSlice& sl = d.slice[c.s];
        c.nk = sl.K / TC_KC;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = s_rows[c.s * TC_BM + r0 + 32 * i];
            c.src[i] = row >= 0 ? sl.src + (size_t)row * sl.ld + j * 4 : nullptr;
        }
    }
}
__device__ __forceinline__ void cursor_next(ChunkCursor& c, const GemmDesc& d, int s_end, uint32_t active,
                                            const int* s_rows, int r0, int j) {
    ++c.wchunk;
    if (++c.kc == c.nk) {
        ++c.s;
        cursor_enter(c, d, s_end, active, s_rows, r0, j);
    }
}
__device__ __forceinline__ void cursor_load(const ChunkCursor& c, float4 (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
        v[i] = c.src[i] ? __ldg(reinterpret_cast<const float4*>(c.src[i] + c.kc * TC_KC)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// advance (s, kc, wchunk) over the live slices of a split range
struct ChunkWalk {
    int s, kc, nk, wchunk;
};
__device__ __forceinline__ void walk_enter(ChunkWalk& c, const GemmDesc& d, int s_end, uint32_t active) {
    while (c.s < s_end && !((active >> c.s) & 1u)) {
        c.wchunk += d.slice[c.s].K / TC_KC;
        ++c.s;
    }
    c.kc = 0;
    c.nk = c.s < s_end ? d.slice[c.s].K / TC_KC : 0;
}

// ------------------------------------------------------------------ tile epilogue (shared by both kernels)
struct EpiParams {
    float4 scale, shift, gw, gb;   // this thread's 4-channel column of the per-channel parameters
};
template <int BN>
__device__ __forceinline__ EpiParams load_epi_params(const GemmDesc& d, int tid) {
    const int c0 = (tid % (BN / 4)) * 4;
    const float4 one = make_float4(1.f, 1.f, 1.f, 1.f), zero = make_float4(0.f, 0.f, 0.f, 0.f);
    EpiParams ep;
    ep.scale = d.scale ? __ldg(reinterpret_cast<const float4*>(d.scale + c0)) : one;
    ep.shift = d.shift ? __ldg(reinterpret_cast<const float4*>(d.shift + c0)) : zero;
    ep.gw = d.gn_weight ? __ldg(reinterpret_cast<const float4*>(d.gn_weight + c0)) : one;
    ep.gb = d.gn_weight ? __ldg(reinterpret_cast<const float4*>(d.gn_bias + c0)) : zero;
    return ep;
}

// s_tile: the 128 x BN accumulator tile in shared memory (row pitch BN + 4); NT threads (a multiple of BN / 4) take part
template <int BN, int NT>
__device__ __forceinline__ void tile_epilogue(const GemmDesc& d, const float* s_tile, long long m0, int tid,
                                              const EpiParams& ep, const float4* ws_tile4, int n_split, long long* stamp) {
    constexpr int UNITS = BN / 4, ITEMS = TC_BM * UNITS, TILE_LD = BN + 4;
    // A thread keeps ONE 4-channel column for all its rows (NT is a multiple of UNITS): the per-channel
    // parameters are loaded once, row addresses advance by a constant, and a warp still covers whole rows, so the 4
    // lanes of a 16-channel GroupNorm group stay adjacent.
    // The loop is SPECIALISED: measured with in-kernel clock stamps (tools/gemm_phases.py), the generic loop spent
    // ~480 cycles per row iteration although it issues ~100 instructions - with 2.5 warps per scheduler the fetch
    // bubbles of its ~15 taken branches per iteration (feature tests, reconvergence points) are not hidden.  The three
    // common shapes (plain store for the pair-major partial rows, affine for Linear / Conv1d / Conv2d, GroupNorm for
    // the sparse layers) get straight-line bodies; split partials, peer stores, pooling and zero rows take the
    // general loop.
    {
        constexpr int ROWS_PER_IT = NT / UNITS;
        static_assert(NT % UNITS == 0, "a thread must keep its column");
        const int c0 = (tid % UNITS) * 4, row0 = tid / UNITS;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool gn = d.gn_weight != nullptr;
        const bool relu = d.relu_out != 0;
        float* out0 = d.out + (size_t)m0 * d.out_ld + c0;
        const float* res0 = d.residual ? d.residual + (size_t)m0 * d.res_ld + c0 : nullptr;
        const int rows_live = (int)(d.M - m0 < TC_BM ? d.M - m0 : TC_BM);
        const int zmod = d.zero_row_mod, zphase = zmod ? (int)(m0 % zmod) : 0;
        const float* tile0 = s_tile + c0;
        const int out_ld = d.out_ld, res_ld = d.res_ld;
        if (stamp && tid == 0) stamp[8] = stamp[9] = clock64();   // column parameters loaded
        const bool general = ws_tile4 || d.n_peers || zmod || (d.pool_out && gn);
        auto group_norm = [&](float4& y) {   // same operation order as epilogue4 (gemm.cuh): identical bits
            float sum = (y.x + y.y) + (y.z + y.w);
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            const float mean = sum * (1.f / 16.f);
            const float dx = y.x - mean, dy = y.y - mean, dz = y.z - mean, dw = y.w - mean;
            float q = fmaf(dx, dx, dy * dy) + fmaf(dz, dz, dw * dw);
            q += __shfl_xor_sync(0xffffffffu, q, 1);
            q += __shfl_xor_sync(0xffffffffu, q, 2);
            const float rstd = 1.f / sqrtf(q * (1.f / 16.f) + 1e-5f);
            y.x = fmaf(dx * rstd, ep.gw.x, ep.gb.x);
            y.y = fmaf(dy * rstd, ep.gw.y, ep.gb.y);
            y.z = fmaf(dz * rstd, ep.gw.z, ep.gb.z);
            y.w = fmaf(dw * rstd, ep.gw.w, ep.gb.w);
        };
        if (!general && !gn && !d.scale && !d.shift && !res0 && !relu) {
            // ---- plain store (pair-major partial rows)
#pragma unroll 4
            for (int row = row0; row < TC_BM; row += ROWS_PER_IT) {
                const float4 y = *reinterpret_cast<const float4*>(tile0 + row * TILE_LD);
                if (row < rows_live) *reinterpret_cast<float4*>(out0 + (size_t)row * out_ld) = y;
            }
        } else if (!general && !gn) {
            // ---- y * scale + shift (+ residual) (+ ReLU) (+ fused segment max of the PointNet layers)
            const bool has_scale = d.scale != nullptr, has_shift = d.shift != nullptr, has_pool = d.pool_out != nullptr;
#pragma unroll 2
            for (int row = row0; row < TC_BM; row += ROWS_PER_IT) {
                float4 y = *reinterpret_cast<const float4*>(tile0 + row * TILE_LD);
                if (has_scale) { y.x *= ep.scale.x; y.y *= ep.scale.y; y.z *= ep.scale.z; y.w *= ep.scale.w; }
                if (has_shift) { y.x += ep.shift.x; y.y += ep.shift.y; y.z += ep.shift.z; y.w += ep.shift.w; }
                if (row < rows_live) {
                    if (res0) {
                        const float4 rv = __ldg(reinterpret_cast<const float4*>(res0 + (size_t)row * res_ld));
                        y.x += rv.x; y.y += rv.y; y.z += rv.z; y.w += rv.w;
                    }
                    if (relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
                    *reinterpret_cast<float4*>(out0 + (size_t)row * out_ld) = y;
                    if (has_pool) {
                        float* o = d.pool_out + (size_t)__ldg(d.pool_seg + m0 + row) * out_ld + c0;
                        atomic_max_f32(o, y.x);
                        atomic_max_f32(o + 1, y.y);
                        atomic_max_f32(o + 2, y.z);
                        atomic_max_f32(o + 3, y.w);
                    }
                }
            }
        } else if (!general) {
            // ---- GroupNorm (+ residual) (+ ReLU)
#pragma unroll 2
            for (int row = row0; row < TC_BM; row += ROWS_PER_IT) {
                float4 y = *reinterpret_cast<const float4*>(tile0 + row * TILE_LD);
                if (d.scale) { y.x *= ep.scale.x; y.y *= ep.scale.y; y.z *= ep.scale.z; y.w *= ep.scale.w; }
                if (d.shift) { y.x += ep.shift.x; y.y += ep.shift.y; y.z += ep.shift.z; y.w += ep.shift.w; }
                group_norm(y);
                if (row < rows_live) {
                    if (res0) {
                        const float4 rv = __ldg(reinterpret_cast<const float4*>(res0 + (size_t)row * res_ld));
                        y.x += rv.x; y.y += rv.y; y.z += rv.z; y.w += rv.w;
                    }
                    if (relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
                    *reinterpret_cast<float4*>(out0 + (size_t)row * out_ld) = y;
                }
            }
        } else {
            // ---- general loop: K-split partials, peer stores (symmetric heap), fused segment max, zero rows
            // rows of this tile that go to peer p: [lo, hi) packed as lo | hi << 16 (halo exchange, GemmDesc::peer_rows)
            unsigned peer_iv[kMaxPeers];
#pragma unroll
            for (int p = 0; p < kMaxPeers; ++p) {
                peer_iv[p] = 0u;
                if (p < d.n_peers) {
                    long long lo = 0, hi = TC_BM;
                    if (d.peer_rows) {
                        const long long g0 = d.row_base + m0;
                        lo = (long long)__ldg(d.peer_rows + 2 * p) - g0;
                        hi = (long long)__ldg(d.peer_rows + 2 * p + 1) - g0;
                        lo = lo < 0 ? 0 : (lo > TC_BM ? TC_BM : lo);
                        hi = hi < 0 ? 0 : (hi > TC_BM ? TC_BM : hi);
                    }
                    peer_iv[p] = (unsigned)lo | ((unsigned)hi << 16);
                }
            }
            for (int row = row0; row < TC_BM; row += ROWS_PER_IT) {
                float4 y;
                if (ws_tile4) {
                    const int i = row * UNITS + (tid % UNITS);
                    // partials in split order; every load is issued before the first add
                    float4 v[TC_MAX_SPLIT];
#pragma unroll
                    for (int sp = 0; sp < TC_MAX_SPLIT; ++sp)
                        v[sp] = sp < n_split ? __ldcg(ws_tile4 + (size_t)sp * ITEMS + i) : zero4;
                    y = v[0];
#pragma unroll
                    for (int sp = 1; sp < TC_MAX_SPLIT; ++sp) {
                        y.x += v[sp].x; y.y += v[sp].y; y.z += v[sp].z; y.w += v[sp].w;
                    }
                } else {
                    y = *reinterpret_cast<const float4*>(tile0 + row * TILE_LD);
                }
                if (d.scale) { y.x *= ep.scale.x; y.y *= ep.scale.y; y.z *= ep.scale.z; y.w *= ep.scale.w; }
                if (d.shift) { y.x += ep.shift.x; y.y += ep.shift.y; y.z += ep.shift.z; y.w += ep.shift.w; }
                if (gn) group_norm(y);
                if (row < rows_live) {
                    if (res0) {
                        const float4 rv = __ldg(reinterpret_cast<const float4*>(res0 + (size_t)row * res_ld));
                        y.x += rv.x; y.y += rv.y; y.z += rv.z; y.w += rv.w;
                    }
                    if (relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
                    if (zmod && (zphase + row) % zmod == d.zero_row_val) y = zero4;
                    const size_t at = (size_t)row * out_ld;
                    *reinterpret_cast<float4*>(out0 + at) = y;
#pragma unroll
                    for (int p = 0; p < kMaxPeers; ++p)
                        if ((unsigned)row >= (peer_iv[p] & 0xffffu) && (unsigned)row < (peer_iv[p] >> 16))
                            *reinterpret_cast<float4*>(d.peer_out[p] + (size_t)m0 * out_ld + c0 + at) = y;
                    if (d.pool_out) {
                        float* o = d.pool_out + (size_t)__ldg(d.pool_seg + m0 + row) * out_ld + c0;
                        atomic_max_f32(o, y.x);
                        atomic_max_f32(o + 1, y.y);
                        atomic_max_f32(o + 2, y.z);
                        atomic_max_f32(o + 3, y.w);
                    }
                }
            }
        }
    }
}

// grid = (row tiles, K splits).  With K splits > 1 every CTA contracts a contiguous range of
// slices, parks its raw 128 x N partial in d.split_ws and bumps the tile's counter; the CTA
// that arrives last adds the partials in split order (deterministic) and runs the epilogue.
template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gather_gemm_tc_kernel(const __grid_constant__ GemmDesc d, int precision) {
    const float* __restrict__ Wp = d.Wp;
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B operands need 1024-byte alignment
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int STAGE = tc_stage_bytes(BN);
    constexpr int B_IMG = BN * 128;
    constexpr int TILE_LD = BN + 4;  // padded row pitch of the epilogue tile (conflict-free both ways)
    static_assert((size_t)TC_BM * TILE_LD * sizeof(float) <= (size_t)TC_STAGES * STAGE, "epilogue tile must fit the stages");
    int* s_rows = reinterpret_cast<int*>(smem + TC_STAGES * STAGE);                 // [n_slices][128]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_rows + kMaxSlices * TC_BM);    // full[3] empty[3] accum
    uint32_t* s_misc = reinterpret_cast<uint32_t*>(s_bar + 2 * TC_STAGES + 1);     // [0] tmem base, [1] active, [2] last
    float* s_tile = reinterpret_cast<float*>(smem);                                 // reuses the stages after the last MMA

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long m0 = (long long)blockIdx.x * TC_BM;
    const uint32_t bar_full = smem_u32(s_bar), bar_empty = smem_u32(s_bar + TC_STAGES),
                   bar_accum = smem_u32(s_bar + 2 * TC_STAGES);
    const int n_split = gridDim.y;
    const int per_split = (d.n_slices + n_split - 1) / n_split;
    const int s_begin = min((int)blockIdx.y * per_split, d.n_slices), s_end = min(s_begin + per_split, d.n_slices);

    long long* stamp = g_gemm_stamps ? g_gemm_stamps + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 12 : nullptr;
    if (stamp && tid == 0) stamp[0] = clock64();
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, TC_PRODUCERS / 32 + 1);  // one elected arrive per producer warp + the copy thread's expect_tx
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_accum, 1);
        s_misc[1] = 0;
        fence_mbar_init();
    }
    // 3xTF32: columns [0, BN) collect A_big*B_big, [BN, 2BN) A_big*B_small (one N = 2 BN MMA), [2BN, 3BN) A_small*B_big:
    // two INDEPENDENT accumulator chains, so consecutive MMAs never wait for each other's write-back
    // (tools/microbench/umma_rate.cu: a dependent N=256 -> N=128 pair costs ~370 cycles, an independent one 192)
    constexpr uint32_t TMEM_COLS = BN == 128 ? 512 : 256;
    if (warp == 8) tmem_alloc(smem_u32(s_misc), TMEM_COLS);
    // everything above is independent of the previous kernel in the stream (PDL prologue)
    if (stamp && tid == 0) stamp[1] = clock64();   // prologue (barriers, TMEM) done
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (stamp && tid == 0) stamp[2] = clock64();   // dependency resolved
    // epilogue column parameters of this thread (it keeps one 4-channel column, see tile_epilogue): requested now,
    // consumed after the main loop - their L2 round trip (~480 cycles) disappears behind the gathers
    const EpiParams epi = load_epi_params<BN>(d, tid);

    // Single-slice launches (pair-major sparse convolution, Linear, Conv2d rows...): every producer thread derives
    // the 4 row numbers it gathers itself (the 8 threads of a row write the same value and each reads back its own),
    // so no block-wide staging pass and no second barrier stand between the dependency wait and the first gather.
    const bool fast_rows = d.n_slices == 1 && !d.kmap;
    if (fast_rows) {
        if (tid < TC_PRODUCERS) {
            const GemmSlice& sl = d.slice[0];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = (tid >> 3) + 32 * i;
                const long long m = m0 + r;
                long long rr = -1;
                if (m < d.M) {
                    rr = sl.idx ? (long long)__ldg(sl.idx + m * sl.idx_stride) : m + sl.shift;
                    if (!sl.idx && rr >= d.n_src_rows) rr = -1;
                }
                s_rows[r] = rr < 0 ? -1 : (int)rr;
            }
        }
    } else {
    // stage the row table of the tile and flag the slices that have at least one live row
    if (tid < TC_PRODUCERS) {
        unsigned my_active = 0;
        const int ns = s_end - s_begin;
        for (int i = tid; i < ns * TC_BM; i += TC_PRODUCERS) {
            int s, r;
            long long rr = -1;
            if (d.kmap) {  // kernel map [M][n_slices]: consecutive threads read consecutive entries
                r = i / ns;
                s = s_begin + (i - r * ns);
                if (m0 + r < d.M) rr = (long long)__ldg(d.kmap + (m0 + r) * d.n_slices + s);
            } else {
                s = s_begin + (i >> 7);
                r = i & (TC_BM - 1);
                const GemmSlice& sl = d.slice[s];
                const long long m = m0 + r;
                if (m < d.M) {
                    rr = sl.idx ? (long long)__ldg(sl.idx + m * sl.idx_stride) : m + sl.shift;
                    if (!sl.idx && rr >= d.n_src_rows) rr = -1;
                }
            }
            if (rr < 0) rr = -1;
            s_rows[s * TC_BM + r] = (int)rr;
            if (rr >= 0) my_active |= 1u << s;
        }
        my_active = __reduce_or_sync(0xffffffffu, my_active);
        if (lane == 0 && my_active) atomicOr(&s_misc[1], my_active);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    }
    if (stamp && tid == 0) stamp[3] = clock64();   // row table staged
    const uint32_t tmem_base = s_misc[0];
    const uint32_t active = fast_rows ? 1u : s_misc[1];
    int n_it = 0;
    for (int s = s_begin; s < s_end; ++s)
        if ((active >> s) & 1u) n_it += d.slice[s].K / TC_KC;

    if (tid < TC_PRODUCERS) {
        // ===================== producers: gather A TC_PREFETCH chunks ahead, split, store swizzled
        const int j = tid & 7;    // 16-byte unit within the 128-byte row
        const int r0 = tid >> 3;  // rows r0 + 32 i
        ChunkCursor ld;
        ld.s = s_begin;
        ld.wchunk = 0;
        cursor_enter(ld, d, s_end, active, s_rows, r0, j);
        float4 buf[TC_PREFETCH][4];
#pragma unroll
        for (int p = 0; p < TC_PREFETCH; ++p) {
            if (p < n_it) {
                cursor_load(ld, buf[p]);
                cursor_next(ld, d, s_end, active, s_rows, r0, j);
            }
        }
        auto store_chunk = [&](int it, const float4 (&v)[4]) {
            const int st = it % TC_STAGES;
            const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
            mbar_wait(bar_empty + 8 * st, ph ^ 1u);
            unsigned char* stage = smem + st * STAGE;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 a = v[i];
                if (d.relu_in) {
                    a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
                }
                float4 big, small;
                split_tf32(a.x, big.x, small.x);
                split_tf32(a.y, big.y, small.y);
                split_tf32(a.z, big.z, small.z);
                split_tf32(a.w, big.w, small.w);
                const int r = r0 + 32 * i;
                const int off = r * 128 + ((j ^ (r & 7)) << 4);
                *reinterpret_cast<float4*>(stage + off) = big;
                *reinterpret_cast<float4*>(stage + TC_A_BYTES + off) = small;
            }
            // every lane makes its stores visible to the async proxy, then ONE arrive per warp (256 serialised
            // arrivals per stage were ~1000 cycles of a 4-chunk tile)
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * st);
        };
        for (int it0 = 0; it0 < n_it; it0 += TC_PREFETCH) {
#pragma unroll
            for (int p = 0; p < TC_PREFETCH; ++p) {
                const int it = it0 + p;
                if (it < n_it) {
                    store_chunk(it, buf[p]);
                    if (it + TC_PREFETCH < n_it) {
                        cursor_load(ld, buf[p]);
                        cursor_next(ld, d, s_end, active, s_rows, r0, j);
                    }
                }
            }
        }
    } else if (warp == 8) {
        // ===================== MMA issuer: the whole warp walks the pipeline (uniform control flow keeps the
        // descriptors in uniform registers, no per-MMA R2UR waterfall), one elected lane issues
        constexpr uint32_t idesc = umma_idesc_tf32(TC_BM, BN);
        // B_big and B_small are adjacent in the stage: together they are ONE K-major operand of 2 BN rows, so
        // A_big is read from shared memory once for both of its products (20 KB instead of 24 KB per K step)
        constexpr uint32_t idesc_pair = umma_idesc_tf32(TC_BM, 2 * BN);
        for (int it = 0; it < n_it; ++it) {
            const int st = it % TC_STAGES;
            const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
            mbar_wait(bar_full + 8 * st, ph);
            tc_fence_after();
            const uint32_t a_big = smem_u32(smem + st * STAGE), a_small = a_big + TC_A_BYTES,
                           b_big = a_big + 2 * TC_A_BYTES;
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < TC_KC / 8; ++kk) {
                    const uint32_t ko = kk * 32;  // 8 tf32 = 32 bytes inside the swizzle row
                    const uint32_t acc = (it | kk) ? 1u : 0u;
                    if (precision == 1) {
                        umma_tf32(tmem_base, umma_desc_sw128(a_big + ko), umma_desc_sw128(b_big + ko), idesc_pair, acc);
                        umma_tf32(tmem_base + 2 * BN, umma_desc_sw128(a_small + ko), umma_desc_sw128(b_big + ko), idesc, acc);
                    } else {
                        umma_tf32(tmem_base, umma_desc_sw128(a_big + ko), umma_desc_sw128(b_big + ko), idesc, acc);
                    }
                }
                umma_commit(bar_empty + 8 * st);  // frees the stage once these MMAs have read it
            }
            __syncwarp();
        }
        if (n_it > 0 && elect_one()) umma_commit(bar_accum);
        __syncwarp();
    } else {
        if (lane == 0 && n_it > 0) {
            // ===================== weight copies: one bulk copy per chunk as soon as its stage is free
            ChunkWalk w;
            w.s = s_begin;
            w.wchunk = d.tile_wslice ? __ldg(d.tile_wslice + blockIdx.x) * (d.slice[0].K / TC_KC) : 0;
            for (int s = 0; s < s_begin; ++s) w.wchunk += d.slice[s].K / TC_KC;
            walk_enter(w, d, s_end, active);
            for (int it = 0; it < n_it; ++it) {
                const int st = it % TC_STAGES;
                const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
                mbar_wait(bar_empty + 8 * st, ph ^ 1u);
                mbar_arrive_expect_tx(bar_full + 8 * st, 2 * B_IMG);
                bulk_g2s(smem_u32(smem + st * STAGE + 2 * TC_A_BYTES), Wp + (size_t)w.wchunk * (2 * BN * 32), 2 * B_IMG,
                         bar_full + 8 * st);
                ++w.wchunk;
                if (++w.kc == w.nk) {
                    ++w.s;
                    walk_enter(w, d, s_end, active);
                }
            }
        }
        __syncwarp();
    }

    // ===================== accumulator tile TMEM -> shared memory: all 8 producer warps, thread = row = TMEM lane
    // (warp w may touch lanes 32 (w % 4) .. +31: warps 0-3 drain the first half of the columns, warps 4-7 the second)
    if (warp < 8) {
        if (stamp && tid == 0) stamp[4] = clock64();   // this producer's last chunk stored
        if (n_it > 0) mbar_wait(bar_accum, 0);  // all MMAs done: the stages are free to be overwritten
        tc_fence_after();
        if (stamp && tid == 0) stamp[5] = clock64();   // accumulator complete
        const int row = (warp & 3) * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const int c_begin = (warp >> 2) * (BN / 2);
#pragma unroll 1
        for (int c0 = c_begin; c0 < c_begin + BN / 2; c0 += 16) {
            float y[16];
            if (n_it > 0) {
                uint32_t ry[16], rz[16], rw[16];   // the three loads are in flight together, one wait
                tmem_ld16_nowait(trow + (uint32_t)c0, ry);
                if (precision == 1) {
                    tmem_ld16_nowait(trow + (uint32_t)(BN + c0), rz);
                    tmem_ld16_nowait(trow + (uint32_t)(2 * BN + c0), rw);
                }
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) y[i] = __uint_as_float(ry[i]);
                if (precision == 1) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) y[i] += __uint_as_float(rz[i]) + __uint_as_float(rw[i]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) y[i] = 0.f;
            }
            float4* tp = reinterpret_cast<float4*>(s_tile + row * TILE_LD + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) tp[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (stamp && tid == 0) stamp[6] = clock64();   // tile in shared memory
    if (warp == 8) tmem_dealloc(tmem_base, TMEM_COLS);

    // ===================== epilogue on 16-byte units, all warps
    constexpr int UNITS = BN / 4;          // per row
    constexpr int ITEMS = TC_BM * UNITS;   // per tile; a multiple of 32, so warps stay converged
    const float4* ws_tile4 = nullptr;
    if (n_split > 1) {
        float4* ws_base = reinterpret_cast<float4*>(d.split_ws) + (size_t)blockIdx.x * n_split * ITEMS;  // [n_split][ITEMS]
        float4* mine = ws_base + (size_t)blockIdx.y * ITEMS;
        for (int i = tid; i < ITEMS; i += TC_THREADS)
            mine[i] = *reinterpret_cast<const float4*>(s_tile + (i / UNITS) * TILE_LD + (i % UNITS) * 4);
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const int prev = atomicAdd(d.split_counters + blockIdx.x, 1);
            const bool last = prev == n_split - 1;
            if (last) d.split_counters[blockIdx.x] = 0;  // ready for the next launch on this stream
            s_misc[2] = last ? 1u : 0u;
        }
        __syncthreads();
        if (!s_misc[2]) return;
        __threadfence();
        ws_tile4 = ws_base;
    }
    tile_epilogue<BN, TC_THREADS>(d, s_tile, m0, tid, epi, ws_tile4, n_split, stamp);
    if (stamp && tid == 0) stamp[7] = clock64();
}

// ------------------------------------------------------------------ persistent variant (single-slice launches)
// Pair-major sparse convolutions, Linear layers and Conv2d rows are ONE slice with K = 64..256: a tile is 2..8 K chunks,
// i.e. ~0.4-1.5 us of MMA inside a ~6.5 us chain of dependent latencies (row numbers -> gather -> split/store -> MMA ->
// TMEM drain -> epilogue, tools/gemm_phases.py).  With more tiles than SMs that chain used to be paid once per tile and
// per wave.  Here grid = min(tiles, 148) CTAs loop over their tiles: the mbarrier rings and the weight copies keep
// running across tiles, the accumulator tile gets its own shared-memory region (so the weight copies of the next tile
// never wait for this tile's epilogue), and the producers request the NEXT tile's row numbers and first chunks right
// after storing the last chunk of this one - that round trip overlaps the MMA wait, the drain and the epilogue.
template <int BN>
__host__ __device__ constexpr int tcp_stages() { return BN == 128 ? 2 : 3; }
template <int BN>
__host__ __device__ constexpr size_t tcp_smem_bytes() {
    return 1024 + (size_t)tcp_stages<BN>() * tc_stage_bytes(BN) + (size_t)TC_BM * (BN + 4) * sizeof(float) + 512 +
           2 * TC_PRODUCERS * 4 * sizeof(int);   // + two buffers of row numbers (4 per producer thread)
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gather_gemm_tc_persistent_kernel(const __grid_constant__ GemmDesc d, int precision) {
    constexpr int PS = tcp_stages<BN>();
    constexpr int STAGE = tc_stage_bytes(BN), B_IMG = BN * 128, TILE_LD = BN + 4;
    constexpr uint32_t TMEM_COLS = BN == 128 ? 512 : 256;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* s_tile = reinterpret_cast<float*>(smem + PS * STAGE);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_tile + TC_BM * TILE_LD);   // full[PS] empty[PS] accum tmem_free
    uint32_t* s_misc = reinterpret_cast<uint32_t*>(s_bar + 2 * PS + 2);
    int* s_rows = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(s_bar) + 512);   // [2][TC_PRODUCERS][4]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_full = smem_u32(s_bar), bar_empty = smem_u32(s_bar + PS), bar_accum = smem_u32(s_bar + 2 * PS),
                   bar_free = smem_u32(s_bar + 2 * PS + 1);
    const GemmSlice& sl = d.slice[0];
    const int nk = sl.K / TC_KC;
    const long long n_tiles = (d.M + TC_BM - 1) / TC_BM;
    const float* __restrict__ Wp = d.Wp;

    if (tid == 0) {
        for (int s = 0; s < PS; ++s) {
            mbar_init(bar_full + 8 * s, TC_PRODUCERS / 32 + 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_accum, 1);
        mbar_init(bar_free, 1);
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(smem_u32(s_misc), TMEM_COLS);
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_misc[0];

    if (warp == 8) {
        // ===================== MMA issuer (uniform control flow, one elected lane)
        constexpr uint32_t idesc = umma_idesc_tf32(TC_BM, BN), idesc_pair = umma_idesc_tf32(TC_BM, 2 * BN);
        int j = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
            if (j > 0) {   // the previous tile has left TMEM
                mbar_wait(bar_free, (uint32_t)(j - 1) & 1u);
                tc_fence_after();
            }
            for (int it = 0; it < nk; ++it) {
                const int g = j * nk + it, st = g % PS;
                mbar_wait(bar_full + 8 * st, (uint32_t)(g / PS) & 1u);
                tc_fence_after();
                const uint32_t a_big = smem_u32(smem + st * STAGE), a_small = a_big + TC_A_BYTES, b_big = a_big + 2 * TC_A_BYTES;
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < TC_KC / 8; ++kk) {
                        const uint32_t ko = kk * 32, acc = (it | kk) ? 1u : 0u;
                        if (precision == 1) {
                            umma_tf32(tmem_base, umma_desc_sw128(a_big + ko), umma_desc_sw128(b_big + ko), idesc_pair, acc);
                            umma_tf32(tmem_base + 2 * BN, umma_desc_sw128(a_small + ko), umma_desc_sw128(b_big + ko), idesc, acc);
                        } else {
                            umma_tf32(tmem_base, umma_desc_sw128(a_big + ko), umma_desc_sw128(b_big + ko), idesc, acc);
                        }
                    }
                    umma_commit(bar_empty + 8 * st);
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(bar_accum);
            __syncwarp();
        }
    } else if (warp == 9) {
        if (lane == 0) {
            // ===================== weight copies: run ahead across tiles, bounded by the stage ring only
            int j = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
                const size_t wchunk0 = d.tile_wslice ? (size_t)__ldg(d.tile_wslice + tile) * nk : 0;
                for (int it = 0; it < nk; ++it) {
                    const int g = j * nk + it, st = g % PS;
                    mbar_wait(bar_empty + 8 * st, ((uint32_t)(g / PS) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(bar_full + 8 * st, 2 * B_IMG);
                    bulk_g2s(smem_u32(smem + st * STAGE + 2 * TC_A_BYTES), Wp + (wchunk0 + it) * (2 * BN * 32), 2 * B_IMG,
                             bar_full + 8 * st);
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== producers + epilogue (warps 0..7)
        const EpiParams epi = load_epi_params<BN>(d, tid);
        const int pj = tid & 7, r0 = tid >> 3;   // 16-byte unit of the 128-byte chunk row; rows r0 + 32 i
        const float* src[4];
        float4 buf[TC_PREFETCH][4];
        // Two requests run ahead of the tile being stored: the ROW NUMBERS of the tile after next and the first chunks of
        // the next tile (whose row numbers arrived a tile ago).  Measured at 13.7 k pair tiles (tools/
        // gemm_persistent_phases.py): issued together, the dependent pair (row number -> row address -> chunk load)
        // stalled every producer thread for 8.3 k of the 20.2 k cycles of a tile period.
        // The row numbers travel global -> shared memory by cp.async (no register, nothing that could stall behind the
        // load): with 168 registers per thread the compiler kept them on the stack, and its store of a freshly loaded
        // value stalled the in-order warp for a full memory round trip, twice per tile.
        auto request_rows = [&](long long tile, int which) {
            const long long m0 = tile * TC_BM;
            int* mine = s_rows + ((size_t)which * TC_PRODUCERS + tid) * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long m = m0 + r0 + 32 * i;
                if (sl.idx && m < d.M) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(mine + i)), "l"(sl.idx + m * sl.idx_stride)
                                 : "memory");
                } else {
                    const long long rr = m + sl.shift;
                    mine[i] = (!sl.idx && m < d.M && rr >= 0 && rr < d.n_src_rows) ? (int)rr : -1;
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto request_chunks = [&](int which) {   // first chunks of the tile whose row numbers were requested a tile ago
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            const int4 rows = *reinterpret_cast<const int4*>(s_rows + ((size_t)which * TC_PRODUCERS + tid) * 4);
            const int rr[4] = {rows.x, rows.y, rows.z, rows.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) src[i] = rr[i] >= 0 ? sl.src + (size_t)(unsigned)rr[i] * sl.ld + pj * 4 : nullptr;
#pragma unroll
            for (int p = 0; p < TC_PREFETCH; ++p)
                if (p < nk) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        buf[p][i] = src[i] ? __ldcg(reinterpret_cast<const float4*>(src[i] + p * TC_KC)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
        };
        if ((long long)blockIdx.x < n_tiles) {
            request_rows(blockIdx.x, 0);
            request_chunks(0);
            if ((long long)blockIdx.x + gridDim.x < n_tiles) request_rows((long long)blockIdx.x + gridDim.x, 1);
        }
        // profiling aid (tools/gemm_phases.py --persistent): the phases of this CTA's 8th tile (steady state)
        long long* const stamps = g_gemm_stamps ? g_gemm_stamps + (size_t)blockIdx.x * 12 : nullptr;
        int j = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
            const long long m0 = tile * TC_BM;
            long long* const stamp = (stamps && tid == 0 && (j == 7 || j == 8)) ? stamps + (j - 7) * 6 : nullptr;
            if (stamp) stamp[0] = clock64();
            for (int it0 = 0; it0 < nk; it0 += TC_PREFETCH) {
#pragma unroll
                for (int p = 0; p < TC_PREFETCH; ++p) {
                    const int it = it0 + p;
                    if (it < nk) {
                        const int g = j * nk + it, st = g % PS;
                        mbar_wait(bar_empty + 8 * st, ((uint32_t)(g / PS) & 1u) ^ 1u);
                        unsigned char* stage = smem + st * STAGE;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float4 a = buf[p][i];
                            if (d.relu_in) {
                                a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
                            }
                            float4 big, small;
                            split_tf32(a.x, big.x, small.x);
                            split_tf32(a.y, big.y, small.y);
                            split_tf32(a.z, big.z, small.z);
                            split_tf32(a.w, big.w, small.w);
                            const int r = r0 + 32 * i;
                            const int off = r * 128 + ((pj ^ (r & 7)) << 4);
                            *reinterpret_cast<float4*>(stage + off) = big;
                            *reinterpret_cast<float4*>(stage + TC_A_BYTES + off) = small;
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_full + 8 * st);
                        if (it + TC_PREFETCH < nk) {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                buf[p][i] = src[i] ? __ldcg(reinterpret_cast<const float4*>(src[i] + (it + TC_PREFETCH) * TC_KC))
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                }
            }
            // the next tile's rows and first chunks travel while this tile finishes
            if (stamp) stamp[1] = clock64();   // chunks stored
            if (tile + gridDim.x < n_tiles) {
                request_chunks((j + 1) & 1);
                if (tile + 2 * (long long)gridDim.x < n_tiles) request_rows(tile + 2 * (long long)gridDim.x, j & 1);
            }
            if (stamp) stamp[2] = clock64();   // next tile requested

            mbar_wait(bar_accum, (uint32_t)j & 1u);
            tc_fence_after();
            if (stamp) stamp[3] = clock64();   // accumulator complete
            // every warp is done reading the previous tile from s_tile
            asm volatile("bar.sync 1, %0;" ::"n"(TC_PRODUCERS) : "memory");
            {
                const int row = (warp & 3) * 32 + lane;
                const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
                const int c_begin = (warp >> 2) * (BN / 2);
#pragma unroll 1
                for (int c0 = c_begin; c0 < c_begin + BN / 2; c0 += 16) {
                    uint32_t ry[16], rz[16], rw[16];
                    tmem_ld16_nowait(trow + (uint32_t)c0, ry);
                    if (precision == 1) {
                        tmem_ld16_nowait(trow + (uint32_t)(BN + c0), rz);
                        tmem_ld16_nowait(trow + (uint32_t)(2 * BN + c0), rw);
                    }
                    tmem_ld_wait();
                    float y[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) y[i] = __uint_as_float(ry[i]);
                    if (precision == 1) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) y[i] += __uint_as_float(rz[i]) + __uint_as_float(rw[i]);
                    }
                    float4* tp = reinterpret_cast<float4*>(s_tile + row * TILE_LD + c0);
#pragma unroll
                    for (int i = 0; i < 4; ++i) tp[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
                }
            }
            tc_fence_before();
            asm volatile("bar.sync 1, %0;" ::"n"(TC_PRODUCERS) : "memory");
            if (tid == 0) mbar_arrive(bar_free);   // TMEM may be overwritten by the next tile's first MMA
            if (stamp) stamp[4] = clock64();   // tile in shared memory
            tile_epilogue<BN, TC_PRODUCERS>(d, s_tile, m0, tid, epi, nullptr, 1, nullptr);
            if (stamp) stamp[5] = clock64();   // rows stored
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, TMEM_COLS);
}

static bool pair_ws_enabled() {
    if (g_pair_ws < 0) {
        const char* e = getenv("DV3D_PAIR_WS");
        g_pair_ws = (e && e[0] == '0') ? 0 : 1;
    }
    return g_pair_ws != 0;
}

// ------------------------------------------------------------------ pair-major GEMM, weight stationary
// The pair-major sparse convolution (sparse_pairs.cu) is a plain GEMM per 128-pair tile: gathered rows [128, Cin] x the
// weight block of the tile's kernel offset [Cin, Cout] -> partial rows P.  Tiles are sorted by offset, so a CTA that takes
// a CONTIGUOUS range of tiles changes its weight block once or twice per launch: the whole block (big + small parts,
// up to 128 KB) stays in shared memory and only the gathered rows stream through the stage ring.  Measured on the
// level-1 layers of the 64-view scene (13.7 k tiles, tools/gemm_persistent_phases.py): with the weights streamed per K
// chunk through a two-stage ring, the third and fourth chunk of every tile waited for a weight copy that could only
// be issued once its stage was free - 128 KB of weights per 64 KB of rows, 15.3 k cycles per tile for 3.1 k of MMA.
// The accumulator goes TMEM -> registers -> global (the rows of P are contiguous), no shared-memory tile.
template <int BN, int NK>
__host__ __device__ constexpr size_t pair_ws_smem_bytes() {
    return 1024 + 3 * (size_t)(2 * TC_A_BYTES) + (size_t)NK * 2 * BN * 128 + 512 + 2 * TC_BM * sizeof(int);
}
static_assert(pair_ws_smem_bytes<128, 4>() <= 232448, "shared memory budget of the weight-stationary pair kernel");

template <int BN, int NK>
__global__ void __launch_bounds__(TC_THREADS, 1)
pair_gemm_ws_kernel(const __grid_constant__ GemmDesc d, int precision) {
    constexpr int PS = 3;   // three of a tile's four K chunks are stored before the first MMA has to finish
    constexpr int A_STAGE = 2 * TC_A_BYTES, B_CHUNK = 2 * BN * 128;
    constexpr uint32_t TMEM_COLS = BN == 128 ? 512 : 256;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* s_w = smem + PS * A_STAGE;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_w + NK * B_CHUNK);   // full[PS] empty[PS] accum free wfull[NK]
    uint32_t* s_misc = reinterpret_cast<uint32_t*>(s_bar + 2 * PS + 2 + NK);
    volatile int* s_done = reinterpret_cast<volatile int*>(s_misc + 2);      // tiles whose MMAs have completed
    int* s_rows = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(s_bar) + 512);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_full = smem_u32(s_bar), bar_empty = smem_u32(s_bar + PS), bar_accum = smem_u32(s_bar + 2 * PS),
                   bar_free = smem_u32(s_bar + 2 * PS + 1), bar_w = smem_u32(s_bar + 2 * PS + 2);
    const GemmSlice& sl = d.slice[0];
    const long long n_tiles = (d.M + TC_BM - 1) / TC_BM;
    const long long t0 = n_tiles * blockIdx.x / gridDim.x, t1 = n_tiles * (blockIdx.x + 1) / gridDim.x;
    const float* __restrict__ Wp = d.Wp;

    if (tid == 0) {
        for (int s = 0; s < PS; ++s) {
            mbar_init(bar_full + 8 * s, TC_PRODUCERS / 32);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_accum, 1);
        mbar_init(bar_free, 1);
        for (int k = 0; k < NK; ++k) mbar_init(bar_w + 8 * k, 1);
        *s_done = 0;
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(smem_u32(s_misc), TMEM_COLS);
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_misc[0];

    if (warp == 8) {
        // ===================== MMA issuer
        constexpr uint32_t idesc = umma_idesc_tf32(TC_BM, BN), idesc_pair = umma_idesc_tf32(TC_BM, 2 * BN);
        int cur = -1, n_w = 0, j = 0;
        for (long long tile = t0; tile < t1; ++tile, ++j) {
            const int ws = __ldg(d.tile_wslice + tile);
            const bool changed = ws != cur;
            if (changed) cur = ws, ++n_w;
            if (j > 0) {   // the previous tile has left TMEM
                mbar_wait(bar_free, (uint32_t)(j - 1) & 1u);
                tc_fence_after();
            }
            for (int it = 0; it < NK; ++it) {
                const int g = j * NK + it, st = g % PS;
                if (changed) mbar_wait(bar_w + 8 * it, (uint32_t)(n_w - 1) & 1u);
                mbar_wait(bar_full + 8 * st, (uint32_t)(g / PS) & 1u);
                tc_fence_after();
                const uint32_t a_big = smem_u32(smem + st * A_STAGE), a_small = a_big + TC_A_BYTES;
                const uint32_t b_big = smem_u32(s_w + it * B_CHUNK);
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < TC_KC / 8; ++kk) {
                        const uint32_t ko = kk * 32, acc = (it | kk) ? 1u : 0u;
                        if (precision == 1) {
                            umma_tf32(tmem_base, umma_desc_sw128(a_big + ko), umma_desc_sw128(b_big + ko), idesc_pair, acc);
                            umma_tf32(tmem_base + 2 * BN, umma_desc_sw128(a_small + ko), umma_desc_sw128(b_big + ko), idesc, acc);
                        } else {
                            umma_tf32(tmem_base, umma_desc_sw128(a_big + ko), umma_desc_sw128(b_big + ko), idesc, acc);
                        }
                    }
                    umma_commit(bar_empty + 8 * st);
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(bar_accum);
            __syncwarp();
        }
    } else if (warp == 9) {
        if (lane == 0) {
            // ===================== weight block: (re)loaded when the kernel offset changes, once the MMAs that read the
            // old block have completed (s_done is advanced by the epilogue warps behind their accumulator wait)
            int cur = -1, j = 0;
            for (long long tile = t0; tile < t1; ++tile, ++j) {
                const int ws = __ldg(d.tile_wslice + tile);
                if (ws == cur) continue;
                cur = ws;
                const long long t_start = clock64();
                while (*s_done < j) {
                    if (clock64() - t_start > 4000000000ll) __trap();
                }
                fence_proxy_async_smem();
                for (int it = 0; it < NK; ++it) {
                    mbar_arrive_expect_tx(bar_w + 8 * it, B_CHUNK);
                    bulk_g2s(smem_u32(s_w + it * B_CHUNK), Wp + ((size_t)ws * NK + it) * (2 * BN * 32), B_CHUNK, bar_w + 8 * it);
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== producers + epilogue (warps 0..7)
        const int pj = tid & 7, r0 = tid >> 3;   // 16-byte unit of the 128-byte chunk row; rows r0 + 32 i
        float4 buf[NK][4];
        // row numbers of a tile by cp.async (see the persistent kernel), one int per row: the first of the 8 threads of a
        // row fetches it; readers wait for their own copies and meet at a barrier before they read
        auto request_rows = [&](long long tile, int which) {
            if (pj == 0) {
                const long long m0 = tile * TC_BM;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = r0 + 32 * i;
                    const long long m = m0 + r;
                    if (m < d.M)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(s_rows + which * TC_BM + r)),
                                     "l"(sl.idx + m * sl.idx_stride)
                                     : "memory");
                    else
                        s_rows[which * TC_BM + r] = -1;
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto request_chunks = [&](int which) {   // every chunk of the tile whose row numbers were requested a tile ago
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("bar.sync 1, %0;" ::"n"(TC_PRODUCERS) : "memory");
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rr = s_rows[which * TC_BM + r0 + 32 * i];
                const float* src = rr >= 0 ? sl.src + (size_t)(unsigned)rr * sl.ld + pj * 4 : nullptr;
#pragma unroll
                for (int p = 0; p < NK; ++p)
                    buf[p][i] = src ? __ldg(reinterpret_cast<const float4*>(src + p * TC_KC)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        if (t0 < t1) {
            request_rows(t0, 0);
            request_chunks(0);
            if (t0 + 1 < t1) request_rows(t0 + 1, 1);
        }
        long long* const stamps = g_gemm_stamps ? g_gemm_stamps + (size_t)blockIdx.x * 12 : nullptr;
        int j = 0;
        for (long long tile = t0; tile < t1; ++tile, ++j) {
            const long long m0 = tile * TC_BM;
            long long* const stamp = (stamps && tid == 0 && (j == 7 || j == 8)) ? stamps + (j - 7) * 6 : nullptr;
            if (stamp) stamp[0] = clock64();
#pragma unroll
            for (int it = 0; it < NK; ++it) {
                const int g = j * NK + it, st = g % PS;
                mbar_wait(bar_empty + 8 * st, ((uint32_t)(g / PS) & 1u) ^ 1u);
                unsigned char* stage = smem + st * A_STAGE;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 a = buf[it][i];
                    float4 big, small;
                    split_tf32(a.x, big.x, small.x);
                    split_tf32(a.y, big.y, small.y);
                    split_tf32(a.z, big.z, small.z);
                    split_tf32(a.w, big.w, small.w);
                    const int r = r0 + 32 * i;
                    const int off = r * 128 + ((pj ^ (r & 7)) << 4);
                    *reinterpret_cast<float4*>(stage + off) = big;
                    *reinterpret_cast<float4*>(stage + TC_A_BYTES + off) = small;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full + 8 * st);
            }
            // the next tile's chunks and the row numbers of the tile after it travel while this tile finishes
            if (stamp) stamp[1] = clock64();   // chunks stored
            if (tile + 1 < t1) {
                request_chunks((j + 1) & 1);
                if (tile + 2 < t1) request_rows(tile + 2, j & 1);
            }
            if (stamp) stamp[2] = clock64();   // next tile requested
            mbar_wait(bar_accum, (uint32_t)j & 1u);
            tc_fence_after();
            if (tid == 0) *s_done = j + 1;
            if (stamp) stamp[3] = stamp[4] = clock64();   // accumulator complete (no shared-memory tile in this kernel)
            {
                // TMEM -> registers -> the (now idle) stage ring as a [128, BN] tile, 16-byte chunks XOR-swizzled by the row
                // so that a thread writing its row and a warp reading one row both hit distinct banks ...
                constexpr int CPR = BN / 4;   // 16-byte chunks per row
                const int row = (warp & 3) * 32 + lane;
                const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
                const int c_begin = (warp >> 2) * (BN / 2);
                unsigned char* srow = smem + (size_t)row * (BN * 4);
#pragma unroll 1
                for (int c0 = c_begin; c0 < c_begin + BN / 2; c0 += 16) {
                    uint32_t ry[16], rz[16], rw[16];
                    tmem_ld16_nowait(trow + (uint32_t)c0, ry);
                    if (precision == 1) {
                        tmem_ld16_nowait(trow + (uint32_t)(BN + c0), rz);
                        tmem_ld16_nowait(trow + (uint32_t)(2 * BN + c0), rw);
                    }
                    tmem_ld_wait();
                    float y[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) y[i] = __uint_as_float(ry[i]);
                    if (precision == 1) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) y[i] += __uint_as_float(rz[i]) + __uint_as_float(rw[i]);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<float4*>(srow + ((((c0 >> 2) + i) ^ (row & (CPR - 1))) << 4)) =
                            make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
                }
                tc_fence_before();
                asm volatile("bar.sync 1, %0;" ::"n"(TC_PRODUCERS) : "memory");
                if (tid == 0) mbar_arrive(bar_free);   // TMEM may be overwritten by the next tile's first MMA
                // ... and out as whole rows: a warp instruction stores 512 contiguous bytes
                constexpr int RPI = 32 / CPR;   // rows per warp instruction (1 or 2)
                const int ch = lane % CPR, sub = lane / CPR;
#pragma unroll 4
                for (int rr = warp * RPI + sub; rr < TC_BM; rr += 8 * RPI) {
                    const float4 v = *reinterpret_cast<const float4*>(smem + (size_t)rr * (BN * 4) + ((ch ^ (rr & (CPR - 1))) << 4));
                    if (m0 + rr < d.M) *reinterpret_cast<float4*>(d.out + (size_t)(m0 + rr) * d.out_ld + ch * 4) = v;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(TC_PRODUCERS) : "memory");   // the ring is free for the next tile's rows
            if (stamp) stamp[5] = clock64();   // rows stored
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, int NK>
static int launch_pair_ws(const GemmDesc& d, int tiles, cudaStream_t st) {
    static std::atomic<unsigned long long> attr{0};
    constexpr size_t smem_bytes = pair_ws_smem_bytes<BN, NK>();
    DV3D_FUNC_SMEM_ONCE(attr, (pair_gemm_ws_kernel<BN, NK>), (int)smem_bytes);
    DV3D_LAUNCH((pair_gemm_ws_kernel<BN, NK>), tiles < kNumSMs ? tiles : kNumSMs, TC_THREADS, smem_bytes, st, d, g_gemm_precision);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

int validate_gather_gemm(const GemmDesc& d, int k_multiple);

// K splits for a launch: fill the 148 SMs when the row tiles alone cannot
int gather_gemm_tc_splits(long long M, int n_slices) {
    const int tiles = cdiv(M, TC_BM);
    if (tiles <= 0 || n_slices < 6) return 1;
    int split = kNumSMs / tiles;
    if (split > n_slices / 3) split = n_slices / 3;
    if (split > TC_MAX_SPLIT) split = TC_MAX_SPLIT;
    return split < 1 ? 1 : split;
}

int launch_gather_gemm_tc(const GemmDesc& d_in, cudaStream_t st) {
    GemmDesc d = d_in;
    symm_attach(d);
    const float* Wp = d.Wp;
    int rc = validate_gather_gemm(d, TC_KC);
    if (rc) return rc;
    DV3D_REQUIRE(Wp && ((uintptr_t)Wp & 15) == 0, "gather_gemm_tc: packed weights must be 16-byte aligned");
    DV3D_REQUIRE(!d.tile_wslice || (d.n_slices == 1 && !d.kmap), "gather_gemm_tc: tile_wslice needs exactly one slice");
    DV3D_REQUIRE(d.out_ld % 4 == 0 && ((uintptr_t)d.out & 15) == 0, "gather_gemm_tc: output must be 16-byte aligned");
    if (d.M == 0) return DV3D_OK;
    const int tiles = cdiv(d.M, TC_BM);
    int split = 1;
    if (d.split_ws && d.split_counters && !d.tile_wslice) {
        split = d.split_hint > 1 ? (d.split_hint < d.n_slices ? d.split_hint : d.n_slices) : gather_gemm_tc_splits(d.M, d.n_slices);
        if (split > TC_MAX_SPLIT) split = TC_MAX_SPLIT;
        const size_t need = (size_t)tiles * split * TC_BM * d.N * sizeof(float);
        if (d.split_ws_bytes < need) split = 1;  // tiles * split <= 148 partials fit a dv3d_sparse_conv_workspace_bytes buffer
    }
    // pair-major sparse convolution (plain store of the partial rows): weight-stationary kernel
    if (d.tile_wslice && d.slice[0].idx && !d.scale && !d.shift && !d.gn_weight && !d.residual && !d.relu_in && !d.relu_out &&
        !d.zero_row_mod && !d.pool_out && d.n_peers == 0 && pair_ws_enabled()) {
        const int nk = d.slice[0].K / TC_KC;
        if (d.N == 128 && nk == 4) return launch_pair_ws<128, 4>(d, tiles, st);
        if (d.N == 128 && nk == 2) return launch_pair_ws<128, 2>(d, tiles, st);
        if (d.N == 64 && nk == 4) return launch_pair_ws<64, 4>(d, tiles, st);
        if (d.N == 64 && nk == 2) return launch_pair_ws<64, 2>(d, tiles, st);
    }
    // persistent variant: one slice, no kernel map, no K split, more tiles than SMs (or forced for tests)
    const bool persist_ok = d.n_slices == 1 && !d.kmap && split == 1 && d.M + (d.slice[0].shift > 0 ? d.slice[0].shift : 0) < (1ll << 31);
    if (persist_ok && (g_gemm_persistent > 0 || (g_gemm_persistent == 0 && tiles > kNumSMs))) {
        const int grid_p = tiles < kNumSMs ? tiles : kNumSMs;
        if (d.N == 128) {
            static std::atomic<unsigned long long> attr{0};
            DV3D_FUNC_SMEM_ONCE(attr, (gather_gemm_tc_persistent_kernel<128>), (int)tcp_smem_bytes<128>());
            DV3D_LAUNCH((gather_gemm_tc_persistent_kernel<128>), grid_p, TC_THREADS, tcp_smem_bytes<128>(), st, d, g_gemm_precision);
        } else {
            static std::atomic<unsigned long long> attr{0};
            DV3D_FUNC_SMEM_ONCE(attr, (gather_gemm_tc_persistent_kernel<64>), (int)tcp_smem_bytes<64>());
            DV3D_LAUNCH((gather_gemm_tc_persistent_kernel<64>), grid_p, TC_THREADS, tcp_smem_bytes<64>(), st, d, g_gemm_precision);
        }
        DV3D_LAUNCHED();
        return DV3D_OK;
    }
    dim3 grid(tiles, split);
    if (d.N == 128) {
        static std::atomic<unsigned long long> attr{0};
        DV3D_FUNC_SMEM_ONCE(attr, (gather_gemm_tc_kernel<128>), (int)tc_smem_bytes(128));
        DV3D_LAUNCH((gather_gemm_tc_kernel<128>), grid, TC_THREADS, tc_smem_bytes(128), st, d, g_gemm_precision);
    } else {
        static std::atomic<unsigned long long> attr{0};
        DV3D_FUNC_SMEM_ONCE(attr, (gather_gemm_tc_kernel<64>), (int)tc_smem_bytes(64));
        DV3D_LAUNCH((gather_gemm_tc_kernel<64>), grid, TC_THREADS, tc_smem_bytes(64), st, d, g_gemm_precision);
    }
    DV3D_LAUNCHED();
    return DV3D_OK;
}

}  // namespace dv3d

using namespace dv3d;

extern "C" size_t dv3d_gemm_pack_bytes(int Ktot, int N) {
    if (Ktot <= 0 || N <= 0 || Ktot % TC_KC) return 0;
    return (size_t)Ktot * N * 2 * sizeof(float);
}

extern "C" int dv3d_gemm_pack_weights(const float* W, int Ktot, int N, void* packed, void* stream) {
    DV3D_REQUIRE(W && packed && Ktot > 0 && Ktot % TC_KC == 0 && (N == 64 || N == 128),
                 "gemm_pack_weights: need K %% 32 == 0 and N in {64,128} (K=%d N=%d)", Ktot, N);
    DV3D_LAUNCH((pack_weights_kernel), cdiv((long long)Ktot * N, 256), 256, 0, (cudaStream_t)stream, W, Ktot, N, (float*)packed);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_set_gemm_precision(int mode) {
    DV3D_REQUIRE(mode == 1 || mode == 2, "set_gemm_precision: 1 = 3xTF32 (fp32-grade), 2 = TF32; got %d", mode);
    g_gemm_precision = mode;
    return DV3D_OK;
}
extern "C" int dv3d_get_gemm_precision(void) { return g_gemm_precision; }

extern "C" int dv3d_set_gemm_persistent(int mode) {
    DV3D_REQUIRE(mode >= -1 && mode <= 1, "set_gemm_persistent: -1 = never, 0 = automatic, 1 = whenever possible; got %d", mode);
    g_gemm_persistent = mode;
    return DV3D_OK;
}

extern "C" int dv3d_set_pair_gemm_mode(int weight_stationary) {
    DV3D_REQUIRE(weight_stationary == 0 || weight_stationary == 1, "set_pair_gemm_mode: 0 or 1");
    g_pair_ws = weight_stationary;
    return DV3D_OK;
}

extern "C" int dv3d_gemm_set_timing_buffer(void* device_buffer) {
    long long* p = (long long*)device_buffer;
    DV3D_CUDA(cudaMemcpyToSymbol(g_gemm_stamps, &p, sizeof(p)));
    return DV3D_OK;
}
