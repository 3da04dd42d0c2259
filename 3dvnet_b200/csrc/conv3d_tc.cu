// First CostRegNet layer (mvsnet.py:137, ConvBnReLU3D(32, 8)) on tcgen05.
//
// Cout = 8 is too narrow for an output-stationary implicit GEMM (an MMA costs the same cycles for N = 8 as for N = 64),
// so the roles are turned: for one output plane z and a patch of 8 x 16 INPUT voxels (one 128-row UMMA tile, one-voxel
// halo included)
//     T[(ry, rx), (ky, kx, co)] = sum over (kz, ci)  in[z + kz - 1, Y0 + ry, X0 + rx, ci] * W[co, ci, kz, ky, kx]
// is a GEMM with K = 3 x 32 (the three planes z-1, z, z+1, 32 channels each = one 128-byte swizzle row per voxel) and
// N = 9 x 8 = 72 (padded to 80), and the layer's output is the in-plane gather
//     out[z, y0 + iy, x0 + ix, co] = sum over (ky, kx)  T[(iy + ky, ix + kx), (ky, kx, co)]
// of the 6 x 14 interior voxels, summed in a fixed order (deterministic) from a shared-memory copy of T.
// A CTA slides along z: every input plane is loaded, split into tf32 big + small parts and stored K-major / SWIZZLE_128B
// once, and serves three output planes from a ring of four slots.  3xTF32: A_big x [W_big; W_small] as one N = 160 MMA
// plus A_small x W_big (N = 80) per K step of 8, 240 TMEM columns per output plane, double buffered: the drain / gather /
// store of plane z runs beside the MMAs of plane z+1.  (Measured, tools/conv3d_phases.py: the 24 MMAs of a plane take
// ~2300 cycles whether issued as two or as four accumulation chains - they stream 186 KB of operands through the
// 128 B/clk shared-memory port - so the columns are better spent on the second buffer.)
// Warps 0-3 produce planes, warps 4-11 drain TMEM / gather / store (two per TMEM lane quarter), warp 12 issues the MMAs.
#include <stdlib.h>

#include <atomic>

#include "tc.cuh"

namespace dv3d {

constexpr int CT_PY = 8, CT_PX = 16;          // input patch (rows x cols) = 128 GEMM rows
constexpr int CT_IY = CT_PY - 2, CT_IX = CT_PX - 2;   // interior = outputs per patch and plane
constexpr int CT_ZS_MAX = 16;                 // output planes per work item: chosen per launch (launch_conv3d_c32_c8_tc)
constexpr int CT_NSLOT = 4;                   // ring of input planes
constexpr int CT_CIN = 32, CT_COUT = 8;
constexpr int CT_NV = 9 * CT_COUT;            // 72 useful columns
constexpr int CT_N = 80;                      // padded to a multiple of 16
constexpr int CT_ACC = 3 * CT_N;              // TMEM columns of one chain pair: [big x big | big x small | small x big]
constexpr int CT_EPI_WARPS = 8, CT_EPI_THREADS = 32 * CT_EPI_WARPS;
constexpr int CT_MMA_WARP = 4 + CT_EPI_WARPS;
constexpr int CT_THREADS = 32 * (CT_MMA_WARP + 1);
constexpr int CT_STAGE_LD = 128;              // T is kept column-major [72][128]: lanes = consecutive rows both ways
constexpr uint32_t CT_A_HALF = 128 * 128;     // one [128 x 32] fp32 tile
constexpr uint32_t CT_A_SLOT = 2 * CT_A_HALF;
constexpr uint32_t CT_B_HALF = CT_N * 128;
constexpr uint32_t CT_B_DZ = 2 * CT_B_HALF;
constexpr uint32_t CT_OFF_B = CT_NSLOT * CT_A_SLOT;
constexpr uint32_t CT_OFF_STAGE = CT_OFF_B + 3 * CT_B_DZ;
constexpr uint32_t CT_OFF_BAR = CT_OFF_STAGE + CT_NV * CT_STAGE_LD * 4;
constexpr uint32_t CT_SMEM = CT_OFF_BAR + 128 + 1024;   // + alignment slack
static_assert(CT_SMEM <= 232448, "shared memory budget");
static_assert(2 * CT_ACC <= 512, "TMEM budget");

struct Conv3dTcArgs {
    const float* x;       // [n, 32, D, H, W]
    const float* weight;  // [8, 32, 3, 3, 3]
    const float* scale;   // [8] folded BatchNorm
    const float* shift;
    float* y;             // [n, 8, D, H, W]
    int n, D, H, W;
    int npx, npy, nz, zs;   // patches in x / y, z segments, output planes per segment
    long long n_items;
    long long* timing;    // profiling aid: 64 clock64 stamps per CTA for its first work item (tools/conv3d_phases.py)
};

__global__ void __launch_bounds__(CT_THREADS, 1)
conv3d_c32_c8_tc_kernel(const Conv3dTcArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* stage = reinterpret_cast<float*>(smem + CT_OFF_STAGE);
    const uint32_t bars = smem_u32(smem + CT_OFF_BAR);
    const uint32_t bar_full = bars, bar_empty = bars + 8 * CT_NSLOT, bar_acc_full = bars + 16 * CT_NSLOT,
                   bar_acc_empty = bar_acc_full + 16;   // two TMEM buffers each
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + CT_OFF_BAR + 16 * CT_NSLOT + 32);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- weights -> K-major tf32 big / small images (constant input: before the dependency wait)
    for (int e = tid; e < 3 * CT_N * CT_CIN; e += CT_THREADS) {
        const int ci = e % CT_CIN, nrow = (e / CT_CIN) % CT_N, kz = e / (CT_CIN * CT_N);
        float big = 0.f, small = 0.f;
        if (nrow < CT_NV) {
            const int co = nrow % CT_COUT, kyx = nrow / CT_COUT;
            split_tf32(__ldg(a.weight + ((size_t)(co * CT_CIN + ci) * 27 + kz * 9 + kyx)), big, small);
        }
        const uint32_t off = (uint32_t)(nrow >> 3) * 1024u + (uint32_t)(nrow & 7) * 128u +
                             ((uint32_t)((ci >> 2) ^ (nrow & 7)) << 4) + (uint32_t)(ci & 3) * 4u;
        unsigned char* b = smem + CT_OFF_B + kz * CT_B_DZ;
        *reinterpret_cast<float*>(b + off) = big;
        *reinterpret_cast<float*>(b + CT_B_HALF + off) = small;
    }
    if (tid == 0) {
        for (int s = 0; s < CT_NSLOT; ++s) {
            mbar_init(bar_full + 8 * s, 4);     // one elected arrive per producer warp
            mbar_init(bar_empty + 8 * s, 1);    // tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_acc_full + 8 * b, 1);
            mbar_init(bar_acc_empty + 8 * b, CT_EPI_THREADS);
        }
        fence_mbar_init();
    }
    if (warp == CT_MMA_WARP) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    long long* stamp = a.timing ? a.timing + 64 * blockIdx.x : nullptr;
    if (stamp && tid == 0) stamp[0] = clock64();

    const long long vol = (long long)a.D * a.H * a.W, plane = (long long)a.H * a.W;
    int n_my = 0;   // items this CTA has started (plane / output counters follow from it)

    for (long long item = blockIdx.x; item < a.n_items; item += gridDim.x, ++n_my) {
        const int px = (int)(item % a.npx), py = (int)((item / a.npx) % a.npy);
        const int zs = (int)((item / ((long long)a.npx * a.npy)) % a.nz), nb = (int)(item / ((long long)a.npx * a.npy * a.nz));
        const int ZS = a.zs;
        const int z0 = zs * ZS, Y0 = py * CT_IY - 1, X0 = px * CT_IX - 1;
        const int pc0 = n_my * (ZS + 2), oc0 = n_my * ZS;   // running plane / output-plane counters of this CTA

        if (warp < 4) {
            // ---------------------------------------------------------------- producers: one input plane per step
            const int rx = lane & 15, half = lane >> 4, gx = X0 + rx;
            for (int p = 0; p < ZS + 2; ++p) {
                const int pc = pc0 + p, slot = pc % CT_NSLOT;
                const int gz = z0 - 1 + p;
                float v[2][16];
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int gy = Y0 + warp * 2 + rr;
                    const bool ok = gz >= 0 && gz < a.D && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
                    const float* src = a.x + ((long long)nb * CT_CIN + half * 16) * vol + (long long)gz * plane + (long long)gy * a.W + gx;
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[rr][i] = ok ? __ldg(src + i * vol) : 0.f;
                }
                mbar_wait(bar_empty + 8 * slot, ((pc / CT_NSLOT) & 1) ^ 1);
                unsigned char* A = smem + slot * CT_A_SLOT;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int r = (warp * 2 + rr) * CT_PX + rx;
                    const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4 big, small;
                        split_tf32(v[rr][4 * q + 0], big.x, small.x);
                        split_tf32(v[rr][4 * q + 1], big.y, small.y);
                        split_tf32(v[rr][4 * q + 2], big.z, small.z);
                        split_tf32(v[rr][4 * q + 3], big.w, small.w);
                        const uint32_t off = row_off + ((uint32_t)((half * 4 + q) ^ (r & 7)) << 4);
                        *reinterpret_cast<float4*>(A + off) = big;
                        *reinterpret_cast<float4*>(A + CT_A_HALF + off) = small;
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full + 8 * slot);
            }
        } else if (warp == CT_MMA_WARP) {
            // ---------------------------------------------------------------- MMA issue
            const uint32_t idesc_pair = umma_idesc_tf32(128, 2 * CT_N), idesc = umma_idesc_tf32(128, CT_N);
            for (int zi = 0; zi < ZS; ++zi) {
                for (int p = (zi == 0 ? 0 : 2); p < 3; ++p) {   // planes zi .. zi+2; the older two were waited for before
                    const int pc = pc0 + zi + p;
                    mbar_wait(bar_full + 8 * (pc % CT_NSLOT), (pc / CT_NSLOT) & 1);
                }
                const int oc = oc0 + zi, buf = oc & 1;
                mbar_wait(bar_acc_empty + 8 * buf, ((oc >> 1) & 1) ^ 1);
                tc_fence_after();
                if (elect_one()) {
                    if (stamp && n_my == 0 && zi < 8) stamp[1 + zi * 5] = clock64();
                    const uint32_t d = tmem_base + buf * CT_ACC;
#pragma unroll
                    for (int kz = 0; kz < 3; ++kz) {
                        const int pc = pc0 + zi + kz;
                        const uint32_t a_big = smem_u32(smem + (pc % CT_NSLOT) * CT_A_SLOT), a_small = a_big + CT_A_HALF;
                        const uint32_t b_big = smem_u32(smem + CT_OFF_B + kz * CT_B_DZ);
#pragma unroll
                        for (int kk = 0; kk < CT_CIN / 8; ++kk) {
                            const uint32_t ko = kk * 32;   // 8 tf32 = 32 bytes inside the swizzle row
                            const uint32_t acc = (kz | kk) ? 1u : 0u;
                            umma_tf32(d, umma_desc_sw128(a_big + ko), umma_desc_sw128(b_big + ko), idesc_pair, acc);
                            umma_tf32(d + 2 * CT_N, umma_desc_sw128(a_small + ko), umma_desc_sw128(b_big + ko), idesc, acc);
                        }
                    }
                    umma_commit(bar_empty + 8 * ((pc0 + zi) % CT_NSLOT));   // plane zi is not read again
                    if (zi == ZS - 1) {
                        umma_commit(bar_empty + 8 * ((pc0 + zi + 1) % CT_NSLOT));
                        umma_commit(bar_empty + 8 * ((pc0 + zi + 2) % CT_NSLOT));
                    }
                    umma_commit(bar_acc_full + 8 * buf);
                    if (stamp && n_my == 0 && zi < 8) stamp[2 + zi * 5] = clock64();
                }
                __syncwarp();
            }
        } else {
            // ---------------------------------------------------------------- epilogue: TMEM -> T in smem -> gather -> y
            const int q = warp & 3, hw = (warp - 4) >> 2, et = tid - 128;   // TMEM lane quarter, column half, thread 0..255
            const int r = q * 32 + lane;
            for (int zi = 0; zi < ZS; ++zi) {
                const int oc = oc0 + zi, buf = oc & 1, gz = z0 + zi;
                mbar_wait(bar_acc_full + 8 * buf, (oc >> 1) & 1);
                tc_fence_after();
                if (stamp && n_my == 0 && et == 0 && zi < 8) stamp[3 + zi * 5] = clock64();
                const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + buf * CT_ACC;
                // this warp's column blocks of T: 0..2 or 3..4 (16 columns each; 72..79 are padding)
                for (int b = hw ? 3 : 0; b < (hw ? 5 : 3); ++b) {
                    uint32_t u[3][16];
#pragma unroll
                    for (int t = 0; t < 3; ++t) tmem_ld16_nowait(trow + t * CT_N + b * 16, u[t]);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (b * 16 + i < CT_NV)   // the two correction terms first
                            stage[(b * 16 + i) * CT_STAGE_LD + r] =
                                __uint_as_float(u[0][i]) + (__uint_as_float(u[1][i]) + __uint_as_float(u[2][i]));
                }
                tc_fence_before();
                mbar_arrive(bar_acc_empty + 8 * buf);            // this TMEM buffer is free for the plane after next
                if (stamp && n_my == 0 && et == 0 && zi < 8) stamp[4 + zi * 5] = clock64();
                asm volatile("bar.sync 1, %0;" ::"n"(CT_EPI_THREADS) : "memory");   // T complete
                if (gz < a.D) {
                    // 16 lanes per output row (14 active): a warp reads two patch rows, whose T rows are 16 apart -
                    // consecutive lanes hit consecutive banks and the two halves never collide
                    for (int slot = et; slot < 16 * CT_IY * CT_COUT; slot += CT_EPI_THREADS) {
                        const int ix = slot & 15, iy = (slot >> 4) % CT_IY, co = (slot >> 4) / CT_IY;
                        const int gy = py * CT_IY + iy, gx = px * CT_IX + ix;
                        if (ix >= CT_IX || gy >= a.H || gx >= a.W) continue;
                        float s = 0.f;
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx)
                                s += stage[((ky * 3 + kx) * CT_COUT + co) * CT_STAGE_LD + (iy + ky) * CT_PX + ix + kx];
                        const float v = fmaxf(fmaf(s, __ldg(a.scale + co), __ldg(a.shift + co)), 0.f);
                        a.y[((long long)nb * CT_COUT + co) * vol + (long long)gz * plane + (long long)gy * a.W + gx] = v;
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(CT_EPI_THREADS) : "memory");   // T consumed: the next plane may overwrite it
                if (stamp && n_my == 0 && et == 0 && zi < 8) stamp[5 + zi * 5] = clock64();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == CT_MMA_WARP) tmem_dealloc(tmem_base, 512);
}

// 0: tcgen05 kernel when the layer has its shape (default), 1: always the CUDA-core kernels (verification / A-B)
static std::atomic<int> g_conv3d_mode{-1};
static std::atomic<long long*> g_conv3d_timing{nullptr};

int conv3d_tc_mode() {
    int m = g_conv3d_mode.load();
    if (m < 0) {
        const char* e = getenv("DV3D_CONV3D");
        m = (e && (e[0] == 'f' || e[0] == '1')) ? 1 : 0;   // DV3D_CONV3D=ffma
        g_conv3d_mode.store(m);
    }
    return m;
}

int launch_conv3d_c32_c8_tc(const float* x, int n, int D, int H, int W, const float* weight, const float* scale,
                            const float* shift, float* y, cudaStream_t st) {
    Conv3dTcArgs a = {};
    a.x = x, a.weight = weight, a.scale = scale, a.shift = shift, a.y = y;
    a.n = n, a.D = D, a.H = H, a.W = W;
    a.npx = cdiv(W, CT_IX), a.npy = cdiv(H, CT_IY);
    // output planes per work item: the fewest plane steps on the busiest CTA, counting the two extra input planes
    // every item loads and a prologue of about one plane step per item
    const long long patches = (long long)n * a.npy * a.npx;
    long long best_cost = -1;
    for (int zs = 2; zs <= CT_ZS_MAX && zs <= (D > 2 ? D : 2); ++zs) {
        const long long items = patches * cdiv(D, zs), rounds = (items + kNumSMs - 1) / kNumSMs;
        const long long cost = rounds * (4 * zs + 3);   // in quarter plane steps
        if (best_cost < 0 || cost < best_cost) best_cost = cost, a.zs = zs;
    }
    a.nz = cdiv(D, a.zs);
    a.n_items = patches * a.nz;
    a.timing = g_conv3d_timing.load();
    static std::atomic<unsigned long long> attr{0};
    DV3D_FUNC_SMEM_ONCE(attr, (conv3d_c32_c8_tc_kernel), (int)CT_SMEM);
    const int grid = (int)(a.n_items < kNumSMs ? a.n_items : kNumSMs);
    DV3D_LAUNCH((conv3d_c32_c8_tc_kernel), grid, CT_THREADS, CT_SMEM, st, a);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

}  // namespace dv3d

extern "C" int dv3d_set_conv3d_mode(int mode) {
    DV3D_REQUIRE(mode == 0 || mode == 1, "set_conv3d_mode: 0 = tcgen05 first layer, 1 = CUDA-core kernels only; got %d", mode);
    dv3d::g_conv3d_mode.store(mode);
    return DV3D_OK;
}
extern "C" int dv3d_get_conv3d_mode(void) { return dv3d::conv3d_tc_mode(); }
extern "C" int dv3d_conv3d_set_timing_buffer(void* device_buffer) {
    dv3d::g_conv3d_timing.store((long long*)device_buffer);
    return DV3D_OK;
}
