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

// ----------------------------------------------------------------------------- per-class event profiler
// bench.py turns this on for ONE instrumented step: every launcher brackets its kernel with two CUDA
// events on the launch stream; fc_profile_end() sums the elapsed time, algorithmic flops and bytes per
// kernel class.  Off (the default) it costs one relaxed load per launch.
enum { FC_CLS_GEMM_FFMA = 0, FC_CLS_GEMM_TC = 1, FC_CLS_ATTENTION = 2, FC_CLS_KNN = 3, FC_CLS_EDGECONV = 4,
       FC_CLS_OTHER = 5, FC_N_CLASSES = 6 };
bool fc_prof_enabled();
void fc_prof_open(int cls, double flops, double bytes, cudaStream_t s);
void fc_prof_close(cudaStream_t s);
struct FcProfScope {
    cudaStream_t s; bool on;
    FcProfScope(int cls, double flops, double bytes, cudaStream_t st) : s(st), on(fc_prof_enabled()) {
        if (on) fc_prof_open(cls, flops, bytes, s);
    }
    ~FcProfScope() { if (on) fc_prof_close(s); }
};
