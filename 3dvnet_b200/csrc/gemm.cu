// Gather-GEMM, fp32 CUDA-core implementation (see gemm.cuh).  64 x N output tile per CTA,
// 256 threads, 4 x (N/16) register tile per thread, K staged through shared memory in
// chunks of 16.  Slices none of whose rows exist in the tile are skipped (sparse kernel maps
// are ~70 % empty), which is where the generalised sparse convolution saves its FLOPs.
#include "gemm.cuh"

namespace dv3d {

constexpr int BM = 64, KC = 16, AS_LD = BM + 4;

template <int BN>
__global__ void __launch_bounds__(256)
gather_gemm_f32_kernel(const __grid_constant__ GemmDesc d) {
    pdl_wait();
    __shared__ __align__(16) float As[KC][AS_LD];
    __shared__ __align__(16) float Ws[KC][BN];
    __shared__ int s_row[BM];
    constexpr int NV = BN / 64;  // float4 column groups per thread

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.x * BM;

    float acc[4][4 * NV];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4 * NV; ++j) acc[i][j] = 0.f;

    int wofs = 0;
    for (int s = 0; s < d.n_slices; ++s) {
        const GemmSlice& sl = d.slice[s];
        int r = -1;
        if (tid < BM) {
            long long m = m0 + tid;
            if (m < d.M) {
                // table rows are trusted (-1 = absent); shifted rows are clipped to the source extent
                long long rr = sl.idx ? (long long)__ldg(sl.idx + m * sl.idx_stride) : m + sl.shift;
                if (rr >= 0 && (sl.idx || rr < d.n_src_rows)) r = (int)rr;
            }
        }
        __syncthreads();  // previous slice's readers of s_row are done
        if (tid < BM) s_row[tid] = r;
        if (!__syncthreads_or(r >= 0)) {
            wofs += sl.K;
            continue;
        }
        const int arow = tid >> 2, akq = tid & 3;
        const int ar = s_row[arow];
        const float* ap = sl.src + (size_t)(ar < 0 ? 0 : ar) * sl.ld + akq * 4;
        for (int kc = 0; kc < sl.K; kc += KC) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ar >= 0) a = __ldg(reinterpret_cast<const float4*>(ap + kc));
            if (d.relu_in) {
                a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
            }
            float4 w[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                int e = tid + i * 256;
                int k = e / (BN / 4), n4 = e % (BN / 4);
                w[i] = __ldg(reinterpret_cast<const float4*>(d.W + (size_t)(wofs + kc + k) * d.N) + n4);
            }
            __syncthreads();  // previous chunk consumed
            As[akq * 4 + 0][arow] = a.x;
            As[akq * 4 + 1][arow] = a.y;
            As[akq * 4 + 2][arow] = a.z;
            As[akq * 4 + 3][arow] = a.w;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                int e = tid + i * 256;
                int k = e / (BN / 4), n4 = e % (BN / 4);
                *reinterpret_cast<float4*>(&Ws[k][n4 * 4]) = w[i];
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    float4 bv = *reinterpret_cast<const float4*>(&Ws[k][v * 64 + tx * 4]);
                    float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][v * 4 + j] = fmaf(a4[i], b4[j], acc[i][v * 4 + j]);
                }
            }
        }
        wofs += sl.K;
    }

    // ------------------------------------------------------------------ epilogue
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        // GroupNorm needs every lane of the 4-lane group in the shuffles: no early exit
        const bool live = m < d.M;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int c0 = v * 64 + tx * 4;
            float y[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float a = acc[i][v * 4 + j];
                if (d.scale) a *= __ldg(d.scale + c0 + j);
                if (d.shift) a += __ldg(d.shift + c0 + j);
                y[j] = a;
            }
            if (d.gn_weight) {
                float s = (y[0] + y[1]) + (y[2] + y[3]);
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                const float mean = s * (1.f / 16.f);
                float q = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) q = fmaf(y[j] - mean, y[j] - mean, q);
                q += __shfl_xor_sync(0xffffffffu, q, 1);
                q += __shfl_xor_sync(0xffffffffu, q, 2);
                const float rstd = 1.f / sqrtf(q * (1.f / 16.f) + 1e-5f);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    y[j] = fmaf((y[j] - mean) * rstd, __ldg(d.gn_weight + c0 + j), __ldg(d.gn_bias + c0 + j));
            }
            if (live) {
                if (d.residual) {
                    float4 rv = __ldg(reinterpret_cast<const float4*>(d.residual + (size_t)m * d.res_ld + c0));
                    y[0] += rv.x; y[1] += rv.y; y[2] += rv.z; y[3] += rv.w;
                }
                if (d.relu_out) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) y[j] = fmaxf(y[j], 0.f);
                }
                if (d.zero_row_mod && (int)(m % d.zero_row_mod) == d.zero_row_val) y[0] = y[1] = y[2] = y[3] = 0.f;
                *reinterpret_cast<float4*>(d.out + (size_t)m * d.out_ld + c0) = make_float4(y[0], y[1], y[2], y[3]);
            }
        }
    }
}

int validate_gather_gemm(const GemmDesc& d, int k_multiple) {
    DV3D_REQUIRE(d.N == 64 || d.N == 128, "gather_gemm: N must be 64 or 128, got %d", d.N);
    DV3D_REQUIRE(d.n_slices >= 1 && d.n_slices <= kMaxSlices, "gather_gemm: bad slice count %d", d.n_slices);
    DV3D_REQUIRE((d.W || d.Wp) && d.out && d.M >= 0 && d.out_ld >= d.N && d.out_ld % 4 == 0, "gather_gemm: bad output");
    DV3D_REQUIRE(!d.gn_weight || d.gn_bias, "gather_gemm: GroupNorm needs weight and bias");
    DV3D_REQUIRE(!d.residual || d.res_ld % 4 == 0, "gather_gemm: residual pitch must be a multiple of 4");
    for (int s = 0; s < d.n_slices; ++s) {
        DV3D_REQUIRE(d.slice[s].src && d.slice[s].K > 0 && d.slice[s].K % k_multiple == 0 && d.slice[s].ld % 4 == 0 &&
                         d.slice[s].ld >= d.slice[s].K && ((uintptr_t)d.slice[s].src & 15) == 0,
                     "gather_gemm: slice %d needs K %% %d == 0, ld %% 4 == 0 and a 16-byte aligned source (K=%d ld=%d)",
                     s, k_multiple, d.slice[s].K, d.slice[s].ld);
    }
    return DV3D_OK;
}

int launch_gather_gemm(const GemmDesc& d, cudaStream_t st) {
    if (d.Wp) return launch_gather_gemm_tc(d, st);
    int rc = validate_gather_gemm(d, KC);
    if (rc) return rc;
    if (d.M == 0) return DV3D_OK;
    const int grid = cdiv(d.M, BM);
    if (d.N == 128)
        DV3D_LAUNCH((gather_gemm_f32_kernel<128>), grid, 256, 0, st, d);
    else
        DV3D_LAUNCH((gather_gemm_f32_kernel<64>), grid, 256, 0, st, d);
    DV3D_LAUNCHED();
    return DV3D_OK;
}

}  // namespace dv3d
