// Pair-major sparse convolution: gather -> GEMM -> ordered scatter.
//
// The output-stationary kernel (sparse.cu -> gemm_tc.cu) contracts, for every 128-row output
// tile, ALL 27 kernel offsets: right when most (row, offset) pairs exist.  On the scenes of the
// benchmark (one reference view, 4 cm voxels) they do not - 4 % of the pairs exist on the
// finest level, 10 % on the second, 38 % on the third - and nine tenths of the tensor-core work
// multiplies zero rows.  Here the rows of the GEMM are the EXISTING (output row, offset) pairs,
// grouped by offset (what MinkowskiEngine's in/out maps are, scenemodeling.py:27-39):
//
//   plan   (once per kernel map)  for every offset k the live pairs in ascending output row,
//          padded to whole 128-row tiles: pair_in[slot] = input row, pair_slot[m][k] = slot,
//          tile_k[t] = offset of tile t.  Three launches for ALL the kernel maps of a scene
//          (per-row-block counts, scan over row blocks, rank + fill), each parallel over row blocks.
//   GEMM   P[slot, :] = feat[pair_in[slot], :] @ W[k(tile)]      gather_gemm_tc_kernel with one
//          slice, the weight block chosen per tile (GemmDesc::tile_wslice), no epilogue
//   reduce out[m, :] = epilogue( sum over k ascending of P[pair_slot[m][k], :] )   one warp-row
//          pass; the order of the sum is fixed, so results are bit-reproducible (the reference's
//          ME path is a sequence of per-offset GEMM + scatter-add in the same offset order)
#include "gemm.cuh"

namespace dv3d {

constexpr int PP_TILE = 128;
// plan header (ints): counts[27] | tile_base[28] | n_tiles | n_pairs | pad -> 64 ints
constexpr int PP_HDR = 64;
constexpr int PP_COUNTS = 0, PP_TBASE = 27, PP_NTILES = 55, PP_NPAIRS = 56;

struct PlanView {
    int* hdr;        // [64]
    int* tile_k;     // [tile_cap]
    int* pair_in;    // [tile_cap * 128]
    int* pair_slot;  // [n_out * 27]
    int* block_cnt;  // [row blocks][27] live rows per (row block of 256, offset)
    int* block_off;  // [row blocks][27] exclusive scan over row blocks
    long long tile_cap;
};

static inline long long plan_tile_cap(long long n_out) { return (n_out * 27 + PP_TILE - 1) / PP_TILE + 27; }

static inline long long plan_row_blocks(long long n_out) { return (n_out + 255) / 256; }

static inline size_t plan_bytes(long long n_out) {
    const long long cap = plan_tile_cap(n_out);
    return sizeof(int) * (size_t)(PP_HDR + align_up(cap, 64) + cap * PP_TILE + n_out * 27 + 2 * plan_row_blocks(n_out) * 27);
}

static inline PlanView plan_view(void* plan, long long n_out) {
    PlanView v;
    v.tile_cap = plan_tile_cap(n_out);
    v.hdr = (int*)plan;
    v.tile_k = v.hdr + PP_HDR;
    v.pair_in = v.tile_k + align_up(v.tile_cap, 64);
    v.pair_slot = v.pair_in + v.tile_cap * PP_TILE;
    v.block_cnt = v.pair_slot + n_out * 27;
    v.block_off = v.block_cnt + plan_row_blocks(n_out) * 27;
    return v;
}

// the kernel maps of a scene are planned together: blockIdx.y selects the map
constexpr int PP_MAX_MAPS = 16;
constexpr int PP_RB = 256;  // rows per row block
struct PlanBatch {
    const int* nbr[PP_MAX_MAPS];
    long long n_out[PP_MAX_MAPS];
    PlanView pv[PP_MAX_MAPS];
};

// Ranking the live rows of every offset in ascending row order is a scan over rows; it is done
// in two levels so that every step is parallel over row blocks of 256:
//   pair_block_count   CTA (row block, map): live rows per offset inside the block
//   pair_block_scan    CTA (offset, map): exclusive scan of the block counts -> block offsets, totals
//   pair_fill          CTA (row block, map): rank inside the block (ballots) + block offset -> slots
__global__ void __launch_bounds__(PP_RB)
pair_block_count_kernel(const __grid_constant__ PlanBatch pb) {
    pdl_wait();
    const int j = blockIdx.y;
    const long long n_out = pb.n_out[j], r0 = (long long)blockIdx.x * PP_RB;
    if (r0 >= n_out) return;
    const int* __restrict__ nbr = pb.nbr[j];
    __shared__ int s_cnt[27];
    if (threadIdx.x < 27) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const long long e_end = (n_out - r0 < PP_RB ? n_out - r0 : PP_RB) * 27;
    for (long long e = threadIdx.x; e < e_end; e += PP_RB)  // coalesced over the block's [rows][27] entries
        if (__ldg(nbr + r0 * 27 + e) >= 0) atomicAdd(&s_cnt[(int)(e % 27)], 1);
    __syncthreads();
    if (threadIdx.x < 27) pb.pv[j].block_cnt[(size_t)blockIdx.x * 27 + threadIdx.x] = s_cnt[threadIdx.x];
}

__global__ void __launch_bounds__(256)
pair_block_scan_kernel(const __grid_constant__ PlanBatch pb) {
    pdl_wait();
    const int k = blockIdx.x, j = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const PlanView pv = pb.pv[j];
    const int n_rb = (int)((pb.n_out[j] + PP_RB - 1) / PP_RB);
    __shared__ int s_w[8];
    __shared__ int s_run;
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n_rb; b0 += 256) {
        const int b = b0 + tid;
        const int c = b < n_rb ? pv.block_cnt[(size_t)b * 27 + k] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        int before = s_run;
        for (int w = 0; w < warp; ++w) before += s_w[w];
        if (b < n_rb) pv.block_off[(size_t)b * 27 + k] = before + incl - c;
        __syncthreads();
        if (tid == 0) {
            int t = s_run;
            for (int w = 0; w < 8; ++w) t += s_w[w];
            s_run = t;
        }
        __syncthreads();
    }
    if (tid == 0) pv.hdr[PP_COUNTS + k] = s_run;
}

__global__ void __launch_bounds__(PP_RB)
pair_fill_kernel(const __grid_constant__ PlanBatch pb) {
    pdl_wait();
    const int j = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n_out = pb.n_out[j], r0 = (long long)blockIdx.x * PP_RB;
    if (r0 >= n_out && blockIdx.x != 0) return;
    const int* __restrict__ nbr = pb.nbr[j];
    const PlanView pv = pb.pv[j];
    __shared__ int s_slot0[27];      // first slot of offset k for this row block
    __shared__ int s_wcnt[8][27];    // live rows per (warp, offset)
    __shared__ int s_tiles[28];      // tile base per offset (exclusive prefix), [27] = total
    if (tid == 0) {
        int base = 0, pairs = 0;
        for (int k = 0; k < 27; ++k) {
            const int c = pv.hdr[PP_COUNTS + k];
            s_tiles[k] = base;
            base += (c + PP_TILE - 1) / PP_TILE;
            pairs += c;
        }
        s_tiles[27] = base;
        if (blockIdx.x == 0) {
            for (int k = 0; k <= 27; ++k) pv.hdr[PP_TBASE + k] = s_tiles[k];
            pv.hdr[PP_NTILES] = base;
            pv.hdr[PP_NPAIRS] = pairs;
        }
    }
    __syncthreads();
    if (tid < 27) s_slot0[tid] = s_tiles[tid] * PP_TILE + (r0 < n_out ? pv.block_off[(size_t)blockIdx.x * 27 + tid] : 0);
    if (blockIdx.x == 0) {
        // tile -> offset table and the padding slots of every offset's last tile
        for (int k = 0; k < 27; ++k) {
            const int t0 = s_tiles[k], t1 = s_tiles[k + 1], c = pv.hdr[PP_COUNTS + k];
            for (int t = t0 + tid; t < t1; t += PP_RB) pv.tile_k[t] = k;
            for (int q = c + tid; q < (t1 - t0) * PP_TILE; q += PP_RB) pv.pair_in[t0 * PP_TILE + q] = -1;
        }
    }
    const long long m = r0 + tid;
    const bool in_range = m < n_out;
    int row[27];
    unsigned bal[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        row[k] = in_range ? __ldg(nbr + m * 27 + k) : -1;
        bal[k] = __ballot_sync(0xffffffffu, row[k] >= 0);
        if (lane == 0) s_wcnt[warp][k] = __popc(bal[k]);
    }
    __syncthreads();
    if (!in_range) return;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        int slot = -1;
        if (row[k] >= 0) {
            int before = 0;
            for (int w = 0; w < warp; ++w) before += s_wcnt[w][k];
            slot = s_slot0[k] + before + __popc(bal[k] & ((1u << lane) - 1u));
            pv.pair_in[slot] = row[k];
        }
        pv.pair_slot[m * 27 + k] = slot;
    }
}

// 16-byte read-only load that the compiler may not sink next to its use: the loads of a batch are
// issued back to back (memory-level parallelism), the ordered adds follow
__device__ __forceinline__ float4 ldg4_issue(const float4* p, bool pred) {
    // predicated inside the asm block (no branch): the nine loads of a batch stay in one basic block
    float4 v;
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.s32 q, %5, 0;\n\t"
        "mov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\tmov.f32 %2, 0f00000000;\n\tmov.f32 %3, 0f00000000;\n\t"
        "@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
        : "l"(p), "r"((int)pred));
    return v;
}

// out[m] = epilogue(sum_k P[pair_slot[m][k]]): thread = (row, 4 channels); a row's UNITS = N/4 lanes are
// consecutive, so the GroupNorm shuffles of epilogue4 stay inside the row.  The row's 27 slots are
// loaded once (one or two per lane) and broadcast by shuffles; the partial rows are fetched in
// three batches of nine independent loads and added in ascending offset order.
template <int N>
__global__ void __launch_bounds__(256)
pair_reduce_kernel(const __grid_constant__ GemmDesc d, const float* __restrict__ P, const int* __restrict__ pair_slot) {
    pdl_wait();
    constexpr int UNITS = N / 4, ROWS = 256 / UNITS;
    const int u = threadIdx.x % UNITS;
    const long long m = (long long)blockIdx.x * ROWS + threadIdx.x / UNITS;
    const bool live = m < d.M;
    const int s_a = (live && u < 27) ? __ldg(pair_slot + m * 27 + u) : -1;
    const int s_b = (live && UNITS < 27 && u + UNITS < 27) ? __ldg(pair_slot + m * 27 + u + UNITS) : -1;
    float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k0 = 0; k0 < 27; k0 += 9) {
        float4 v[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const int k = k0 + j;
            const int slot = __shfl_sync(0xffffffffu, k < UNITS ? s_a : s_b, k % UNITS, UNITS);
            v[j] = ldg4_issue(reinterpret_cast<const float4*>(P + (size_t)(slot < 0 ? 0 : slot) * N) + u, slot >= 0);
        }
#pragma unroll
        for (int j = 0; j < 9; ++j) {  // offsets in ascending order; absent pairs add +0
            y.x += v[j].x; y.y += v[j].y; y.z += v[j].z; y.w += v[j].w;
        }
    }
    epilogue4(d, y, m, 4 * u, live, false);
}

}  // namespace dv3d

using namespace dv3d;

extern "C" size_t dv3d_pair_plan_bytes(long long n_out) { return n_out < 0 ? 0 : plan_bytes(n_out); }

extern "C" int dv3d_pair_plan_build(const int* const* nbrs, const long long* n_outs, void* const* plans,
                                    const size_t* plan_bytes_given, int n_maps, void* stream) {
    DV3D_REQUIRE(nbrs && n_outs && plans && plan_bytes_given && n_maps >= 0 && n_maps <= PP_MAX_MAPS,
                 "pair_plan_build: bad arguments (at most %d maps per call)", PP_MAX_MAPS);
    if (n_maps == 0) return DV3D_OK;
    PlanBatch pb = {};
    for (int i = 0; i < n_maps; ++i) {
        DV3D_REQUIRE(nbrs[i] && plans[i] && n_outs[i] >= 0 && ((uintptr_t)plans[i] & 15) == 0, "pair_plan_build: bad map %d", i);
        DV3D_REQUIRE(plan_bytes_given[i] >= plan_bytes(n_outs[i]),
                     "pair_plan_build: plan buffer %d smaller than dv3d_pair_plan_bytes(n_out)", i);
        DV3D_REQUIRE(n_outs[i] * 27 < (1ll << 31) - 27 * PP_TILE, "pair_plan_build: level too large for 32-bit pair slots");
        pb.nbr[i] = nbrs[i];
        pb.n_out[i] = n_outs[i];
        pb.pv[i] = plan_view(plans[i], n_outs[i]);
    }
    cudaStream_t st = (cudaStream_t)stream;
    long long max_rb = 1;
    for (int i = 0; i < n_maps; ++i)
        if (plan_row_blocks(n_outs[i]) > max_rb) max_rb = plan_row_blocks(n_outs[i]);
    DV3D_REQUIRE(max_rb <= 0x7fffffff, "pair_plan_build: level too large");
    DV3D_LAUNCH((pair_block_count_kernel), dim3((unsigned)max_rb, n_maps), PP_RB, 0, st, pb);
    DV3D_LAUNCHED();
    DV3D_LAUNCH((pair_block_scan_kernel), dim3(27, n_maps), 256, 0, st, pb);
    DV3D_LAUNCHED();
    DV3D_LAUNCH((pair_fill_kernel), dim3((unsigned)max_rb, n_maps), PP_RB, 0, st, pb);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

extern "C" int dv3d_pair_plan_counts(const void* const* plans, int n_plans, long long* n_tiles_host, long long* n_pairs_host,
                                     void* stream) {
    DV3D_REQUIRE(plans && n_plans >= 0 && n_plans <= 64 && n_tiles_host, "pair_plan_counts: bad arguments");
    int tmp[64][2];
    cudaStream_t st = (cudaStream_t)stream;
    for (int i0 = 0; i0 < n_plans; i0 += 16) {   // one mailbox message per 16 plans
        ReadItem items[16];
        const int n = n_plans - i0 < 16 ? n_plans - i0 : 16;
        for (int i = 0; i < n; ++i) items[i] = ReadItem{(const int*)plans[i0 + i] + PP_NTILES, tmp[i0 + i], 2 * (int)sizeof(int)};
        int rc = read_back(items, n, st);
        if (rc) return rc;
    }
    for (int i = 0; i < n_plans; ++i) {
        n_tiles_host[i] = tmp[i][0];
        if (n_pairs_host) n_pairs_host[i] = tmp[i][1];
    }
    return DV3D_OK;
}

extern "C" int dv3d_sparse_conv_prefers_pairs(long long n_out, long long n_tiles) {
    // tile-chunks of tensor-core work: pair-major n_tiles vs output-stationary 27 per 128 rows
    // (the pair path adds the reduce pass and a second launch, so it must win clearly)
    const long long dense = 27 * ((n_out + PP_TILE - 1) / PP_TILE);
    if (n_tiles <= 0) return 0;
    // measured with the weight-stationary pair kernel (tools/sparse_conv_paths.py, 128 channels): at 0.60 of the dense
    // work the pair path loses on a 54.8 k-row level (324 vs 277 us: its partial rows make a round trip through HBM) but
    // wins on 6.9 k and 13.7 k rows (46 vs 66, 89 vs 99 us), where the output-stationary kernel cannot fill the SMs
    if (dense <= 3000) return 10 * n_tiles <= 7 * dense;
    return 10 * n_tiles <= 6 * dense;
}

extern "C" size_t dv3d_sparse_conv_pairs_workspace_bytes(long long n_tiles, int Cout) {
    return n_tiles < 0 || Cout <= 0 ? 0 : (size_t)n_tiles * PP_TILE * Cout * sizeof(float);
}

extern "C" int dv3d_sparse_conv_pairs(const float* feat, long long n_in, int Cin, const void* plan, long long n_tiles,
                                      long long n_out, const void* W_packed, int Cout, const float* gn_weight,
                                      const float* gn_bias, const float* residual, int relu, void* workspace,
                                      size_t workspace_bytes, float* out, void* stream) {
    DV3D_REQUIRE(feat && plan && W_packed && out && n_in >= 0 && n_out >= 0 && n_tiles >= 0, "sparse_conv_pairs: bad arguments");
    DV3D_REQUIRE(Cout == 64 || Cout == 128, "sparse_conv_pairs: Cout must be 64 or 128, got %d", Cout);
    DV3D_REQUIRE(n_tiles <= plan_tile_cap(n_out), "sparse_conv_pairs: n_tiles exceeds the plan's capacity");
    DV3D_REQUIRE(workspace_bytes >= dv3d_sparse_conv_pairs_workspace_bytes(n_tiles, Cout) && (workspace || n_tiles == 0),
                 "sparse_conv_pairs: workspace smaller than dv3d_sparse_conv_pairs_workspace_bytes");
    if (n_out == 0) return DV3D_OK;
    PlanView pv = plan_view(const_cast<void*>(plan), n_out);
    cudaStream_t st = (cudaStream_t)stream;
    float* P = (float*)workspace;
    if (n_tiles > 0) {
        GemmDesc g = {};
        g.n_slices = 1;
        g.slice[0] = GemmSlice{feat, pv.pair_in, 1, 0, Cin, Cin};
        g.tile_wslice = pv.tile_k;
        g.Wp = (const float*)W_packed;
        g.M = n_tiles * PP_TILE;
        g.n_src_rows = n_in;
        g.N = Cout;
        g.out = P;
        g.out_ld = Cout;
        int rc = launch_gather_gemm_tc(g, st);
        if (rc) return rc;
    }
    GemmDesc e = {};
    e.M = n_out;
    e.N = Cout;
    e.gn_weight = gn_weight;
    e.gn_bias = gn_bias;
    e.residual = residual;
    e.res_ld = Cout;
    e.relu_out = relu;
    e.out = out;
    e.out_ld = Cout;
    symm_attach(e);
    DV3D_REQUIRE(!gn_weight || gn_bias, "sparse_conv_pairs: GroupNorm needs weight and bias");
    if (Cout == 128)
        DV3D_LAUNCH((pair_reduce_kernel<128>), cdiv(n_out, 8), 256, 0, st, e, P, pv.pair_slot);
    else
        DV3D_LAUNCH((pair_reduce_kernel<64>), cdiv(n_out, 16), 256, 0, st, e, P, pv.pair_slot);
    DV3D_LAUNCHED();
    return DV3D_OK;
}
