// Microbenchmark: issue rate of tcgen05.mma (cta_group::1, M = 128) on sm_100a for the shapes the decoder / gather
// GEMM kernels use.  Operands are whatever shared memory / TMEM holds (values do not matter for timing).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate tools/microbench/umma_rate.cu && ./umma_rate
// Prints cycles per MMA (clock64 around ITER back-to-back MMAs + commit + wait), for one CTA alone on the GPU and
// with 148 CTAs running the same loop.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    return (uint64_t)((a >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_sw64(uint32_t a) {
    return (uint64_t)((a >> 4) & 0x3FFF) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
__host__ __device__ constexpr uint32_t idesc(int fmt, int M, int N) {  // fmt 2 = tf32, 1 = bf16, 0 = f16
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int KIND>  // 0 tf32, 1 f16-kind
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts_tf32(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(id), "r"(acc) : "memory");
}

// mode: 0 tf32 SS N=256 | 1 tf32 SS N=128 | 2 tf32 SS N=256+N=128 pair (count as 2) | 3 bf16 SS N=256 | 4 bf16 SS N=128
//       5 tf32 TS N=256 | 6 tf32 TS N=128 | 7 tf32 SS N=256, SW64 operands | 8 tf32 SS N=64
__global__ void __launch_bounds__(128, 1) bench(int mode, int iters, long long* out) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const uint32_t b = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((float*)smem)[i] = 0.001f * (i & 255);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tslot;
    if (threadIdx.x == 0) {
        const uint32_t sa = smem_u32(smem);  // 3 stages of 48 KB: A 16 KB, B 32 KB
        long long t0 = clock64();
        int n = 0;
        for (int it = 0; it < iters; ++it) {
            const uint32_t st = sa + (it % 3) * 48 * 1024;
            const uint32_t A = st, B = st + 16 * 1024;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint32_t ko = kk * 32;
                switch (mode) {
                    case 0: mma_ss<0>(tm, desc_sw128(A + ko), desc_sw128(B + ko), idesc(2, 128, 256), 1); n += 1; break;
                    case 1: mma_ss<0>(tm, desc_sw128(A + ko), desc_sw128(B + ko), idesc(2, 128, 128), 1); n += 1; break;
                    case 2:
                        mma_ss<0>(tm, desc_sw128(A + ko), desc_sw128(B + ko), idesc(2, 128, 256), 1);
                        mma_ss<0>(tm + 256, desc_sw128(A + ko), desc_sw128(B + 16384 + ko), idesc(2, 128, 128), 1);
                        n += 2;
                        break;
                    case 3: mma_ss<1>(tm, desc_sw128(A + ko), desc_sw128(B + ko), idesc(1, 128, 256), 1); n += 1; break;
                    case 4: mma_ss<1>(tm, desc_sw128(A + ko), desc_sw128(B + ko), idesc(1, 128, 128), 1); n += 1; break;
                    case 5: mma_ts_tf32(tm, tm + 384 + kk * 8, desc_sw128(B + ko), idesc(2, 128, 256), 1); n += 1; break;
                    case 6: mma_ts_tf32(tm, tm + 384 + kk * 8, desc_sw128(B + ko), idesc(2, 128, 128), 1); n += 1; break;
                    case 7: mma_ss<0>(tm, desc_sw64(A + (kk & 1) * 32 + (kk >> 1) * 8192), desc_sw64(B + (kk & 1) * 32 + (kk >> 1) * 16384), idesc(2, 128, 256), 1); n += 1; break;
                    case 8: mma_ss<0>(tm, desc_sw128(A + ko), desc_sw128(B + ko), idesc(2, 128, 64), 1); n += 1; break;
                    case 9: mma_ss<0>(tm, desc_sw128(A + ko), desc_sw64(B + (kk & 1) * 32 + (kk >> 1) * 16384), idesc(2, 128, 256), 1); n += 1; break;
                    case 10: mma_ss<0>(tm, desc_sw64(A + (kk & 1) * 32 + (kk >> 1) * 8192), desc_sw128(B + ko), idesc(2, 128, 256), 1); n += 1; break;
                    case 11:
                        mma_ss<0>(tm, desc_sw128(A + ko), desc_sw128(B + ko), idesc(2, 128, 256), 1);
                        mma_ss<0>(tm + 256, desc_sw128(A + ko), desc_sw128(B + ko), idesc(2, 128, 256), 1);
                        n += 2;
                        break;
                    case 12:
                        mma_ss<0>(tm, desc_sw128(A + ko), desc_sw128(B + ko), idesc(2, 128, 128), 1);
                        mma_ss<0>(tm + 128, desc_sw128(A + ko), desc_sw128(B + 16384 + ko), idesc(2, 128, 128), 1);
                        mma_ss<0>(tm + 256, desc_sw128(A + ko), desc_sw128(B + ko), idesc(2, 128, 128), 1);
                        n += 3;
                        break;
                    case 13:
                        mma_ss<0>(tm, desc_sw64(A + (kk & 1) * 32 + (kk >> 1) * 8192), desc_sw64(B + (kk & 1) * 32), idesc(2, 128, 256), 1);
                        mma_ss<0>(tm + 256, desc_sw64(A + (kk & 1) * 32 + (kk >> 1) * 8192), desc_sw64(B + 16384 + (kk & 1) * 32), idesc(2, 128, 128), 1);
                        n += 2;
                        break;
                    case 14:
                        mma_ss<1>(tm, desc_sw128(A + ko), desc_sw128(B + ko), idesc(1, 128, 256), 1);
                        mma_ss<1>(tm + 256, desc_sw128(A + ko), desc_sw128(B + ko), idesc(1, 128, 256), 1);
                        n += 2;
                        break;
                    case 15:
                        mma_ss<0>(tm, desc_sw128(A + ko), desc_sw128(B + ko), idesc(2, 128, 192), 1);
                        mma_ss<0>(tm + 192, desc_sw128(A + ko), desc_sw128(B + ko), idesc(2, 128, 192), 1);
                        n += 2;
                        break;
                    case 16:   // 3xTF32 k-step as the decoder issues it, SW128: (256,128) x 3 terms
                        for (int t = 0; t < 3; ++t) {
                            mma_ss<0>(tm, desc_sw128(A + ko + (t == 1 ? 8192 : 0)), desc_sw128(B + ko), idesc(2, 128, 256), 1);
                            mma_ss<0>(tm + 256, desc_sw128(A + ko + (t == 1 ? 8192 : 0)), desc_sw128(B + 16384 + ko), idesc(2, 128, 128), 1);
                        }
                        n += 6;
                        break;
                    case 17:   // same work as three N=128 chains (one per tap), terms innermost
                        for (int t = 0; t < 3; ++t) {
                            mma_ss<0>(tm, desc_sw128(A + ko), desc_sw128(B + ko), idesc(2, 128, 128), 1);
                            mma_ss<0>(tm + 128, desc_sw128(A + ko), desc_sw128(B + 8192 + ko), idesc(2, 128, 128), 1);
                            mma_ss<0>(tm + 256, desc_sw128(A + ko), desc_sw128(B + 16384 + ko), idesc(2, 128, 128), 1);
                        }
                        n += 9;
                        break;
                    case 18: mma_ss<1>(tm, desc_sw128(A + ko), desc_sw128(B + ko), idesc(0, 128, 256), 1); n += 1; break;
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b) : "memory");
        long long t1 = clock64();
        if (blockIdx.x == 0) {
            out[0] = t1 - t0;
            out[1] = n;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

// Tight issue loop: descriptors are built once, every MMA is "add a constant to the low word + issue"; pattern =
// the decoder's 3xTF32 K step: (N=256 -> cols 0, N=128 -> cols 256) x 3 terms, 4 (SW128) or 2 (SW64) K steps per stage.
template <int SW>
__global__ void __launch_bounds__(128, 1) bench_tight(int iters, long long* out) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const uint32_t b = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 192 * 1024 / 4; i += 128) ((float*)smem)[i] = 0.001f * (i & 255);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tslot;
    if (threadIdx.x == 0) {
        const uint32_t sa = smem_u32(smem);
        constexpr uint32_t id256 = idesc(2, 128, 256), id128 = idesc(2, 128, 128);
        long long t0 = clock64();
        int n = 0;
        for (int it = 0; it < iters; ++it) {
            const uint32_t st = sa + (it & 1) * 96 * 1024;
            // stage: A big 16K | A small 16K | B big 32K (rows 0..255 then 256..383 at +rows*pitch) | B small 32K
            const uint64_t a_big = SW == 128 ? desc_sw128(st) : desc_sw64(st);
            const uint64_t a_small = SW == 128 ? desc_sw128(st + 16384) : desc_sw64(st + 16384);
            const uint64_t b_big = SW == 128 ? desc_sw128(st + 32768) : desc_sw64(st + 32768);
            const uint64_t b_small = SW == 128 ? desc_sw128(st + 40960) : desc_sw64(st + 40960);
            constexpr uint64_t hi = SW == 128 ? (256 * 128) >> 4 : (256 * 64) >> 4;  // rows 256.. of a B image
#pragma unroll
            for (int kk = 0; kk < (SW == 128 ? 4 : 2); ++kk) {
                const uint64_t ko = kk * 2;  // 32 bytes >> 4
                mma_ss<0>(tm, a_big + ko, b_big + ko, id256, 1);
                mma_ss<0>(tm + 256, a_big + ko, b_big + hi + ko, id128, 1);
                mma_ss<0>(tm, a_small + ko, b_big + ko, id256, 1);
                mma_ss<0>(tm + 256, a_small + ko, b_big + hi + ko, id128, 1);
                mma_ss<0>(tm, a_big + ko, b_small + ko, id256, 1);
                mma_ss<0>(tm + 256, a_big + ko, b_small + hi + ko, id128, 1);
                n += 6;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b) : "memory");
        long long t1 = clock64();
        if (blockIdx.x == 0) {
            out[0] = t1 - t0;
            out[1] = n;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
    long long* out;
    cudaMalloc(&out, 16);
    const int smem = 1024 + 160 * 1024;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[] = {"tf32 SS N=256", "tf32 SS N=128", "tf32 SS N=256 + N=128 (same A)", "bf16 SS N=256", "bf16 SS N=128",
                           "tf32 TS N=256 (A in TMEM)", "tf32 TS N=128 (A in TMEM)", "tf32 SS N=256 SW64", "tf32 SS N=64",
                           "tf32 N=256 A sw128 B sw64", "tf32 N=256 A sw64 B sw128", "tf32 2 chains N=256", "tf32 3 chains N=128",
                           "tf32 (256+128) SW64 [decoder now]", "bf16 2 chains N=256", "tf32 2 chains N=192",
                           "3xTF32 k-step (256,128)x3 SW128", "3xTF32 k-step 3 chains N=128 x3", "fp16 SS N=256"};
    for (int grid : {148}) {
        for (int mode = 0; mode < 19; ++mode) {
            long long h[2] = {0, 0};
            for (int rep = 0; rep < 2; ++rep) {
                bench<<<grid, 128, smem>>>(mode, 2000, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) {
                    printf("mode %d: %s\n", mode, cudaGetErrorString(e));
                    return 1;
                }
            }
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("grid %3d  %-34s %8.1f cycles / MMA  (%lld MMAs)\n", grid, names[mode], (double)h[0] / (double)h[1], h[1]);
        }
    }
    const int smem2 = 1024 + 192 * 1024;
    cudaFuncSetAttribute(bench_tight<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
    cudaFuncSetAttribute(bench_tight<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
    for (int sw : {128, 64}) {
        long long h[2] = {0, 0};
        for (int rep = 0; rep < 2; ++rep) {
            if (sw == 128) bench_tight<128><<<148, 128, smem2>>>(2000, out);
            else bench_tight<64><<<148, 128, smem2>>>(4000, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("tight sw%d: %s\n", sw, cudaGetErrorString(e));
                return 1;
            }
        }
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("grid 148  tight issue, 3xTF32 K step, SW%-3d     %8.1f cycles / MMA = %.0f cycles / K step (nominal 576)\n", sw,
               (double)h[0] / (double)h[1], 6.0 * (double)h[0] / (double)h[1]);
    }
    return 0;
}
