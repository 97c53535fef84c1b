// flowcompare_b200 -- shared device/host helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/flowcompare_b200.h"

#define FC_CUDA_OK(expr)                                   \
    do {                                                   \
        cudaError_t _e = (expr);                           \
        if (_e != cudaSuccess) { fc_set_last_cuda_error((int)_e, __FILE__, __LINE__); return FC_ERR_CUDA; } \
    } while (0)

#define FC_LAUNCH_OK()                                     \
    do {                                                   \
        cudaError_t _e = cudaGetLastError();               \
        if (_e != cudaSuccess) { fc_set_last_cuda_error((int)_e, __FILE__, __LINE__); return FC_ERR_LAUNCH; } \
    } while (0)

#define FC_REQUIRE(cond)                                   \
    do { if (!(cond)) return FC_ERR_INVALID_ARG; } while (0)

void fc_set_last_cuda_error(int code, const char* file, int line);

static inline int fc_round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline long long fc_round_up_ll(long long x, long long m) { return (x + m - 1) / m * m; }

// ----------------------------------------------------------------------------- device math
__device__ __forceinline__ float fc_gelu_erf(float x) {
    // exact GELU (torch.nn.GELU default): 0.5*x*(1+erf(x/sqrt(2)))
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float fc_leaky_relu02(float x) { return x > 0.f ? x : 0.2f * x; }

__device__ __forceinline__ float fc_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float fc_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
