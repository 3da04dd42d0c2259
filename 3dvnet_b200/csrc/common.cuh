// Shared helpers for lib3dvnet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dv3d.h"

namespace dv3d {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define DV3D_REQUIRE(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            dv3d::set_error(__VA_ARGS__);       \
            return DV3D_EINVAL;                 \
        }                                       \
    } while (0)

#define DV3D_CUDA(expr)                                                                        \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            dv3d::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return DV3D_ECUDA;                                                                 \
        }                                                                                      \
    } while (0)

// after every kernel launch: counts it and surfaces launch-configuration errors
#define DV3D_LAUNCHED()                      \
    do {                                     \
        dv3d::count_launch();                \
        DV3D_CUDA(cudaGetLastError());       \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

}  // namespace dv3d
