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
// Branch-free GELU for the tensor-core GEMM's epilogue (the epilogue is issue-bound: erff() costs ~30 instructions
// with both of its branches executed in a divergent warp, this costs 12).
//   gelu(x) = relu(x) - |x| * (0.5*erfc(|x|/sqrt2)),   0.5*erfc(u/sqrt2) ~= 2^(u*r(u) - 1),  r = degree-7 fit on [0, 6.6]
// (beyond 6.6 the term is < 1e-10; the 1/sqrt2 and the 0.5 are folded into the fit and the exponent).  Max absolute
// error of gelu vs the exact form: 6e-8 (fp32 evaluation), i.e. below one ulp of the O(1) activations; unlike
// 0.5*x*(1+erf) it has no cancellation for negative x.
__device__ __forceinline__ float fc_gelu_erf_fast(float x) {
    const float u = fminf(fabsf(x), 6.6f);
    float r = -2.107304487601092e-06f;
    r = fmaf(r, u, 3.082007880290574e-05f);
    r = fmaf(r, u, -0.00014644312103818846f);
    r = fmaf(r, u, -0.0002300897957034123f);
    r = fmaf(r, u, 0.007180221614911477f);
    r = fmaf(r, u, -0.052572273431757154f);
    r = fmaf(r, u, -0.4591854841684466f);
    r = fmaf(r, u, -1.1511073739593443f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(r, u, -1.0f)));
    return fmaf(-fabsf(x), e, fmaxf(x, 0.f));
}
// Two values at a time with Blackwell's packed fp32 instructions (FFMA2: one issue slot for two fmas).  Every component goes
// through the same operation sequence as fc_gelu_erf_fast, so the results are bit-identical to it.
__device__ __forceinline__ float2 fc_gelu_erf_fast2(float2 x) {
    const float2 na = make_float2(-fabsf(x.x), -fabsf(x.y));
    const float2 u = make_float2(fminf(fabsf(x.x), 6.6f), fminf(fabsf(x.y), 6.6f));
    float2 r = make_float2(-2.107304487601092e-06f, -2.107304487601092e-06f);
    r = __ffma2_rn(r, u, make_float2(3.082007880290574e-05f, 3.082007880290574e-05f));
    r = __ffma2_rn(r, u, make_float2(-0.00014644312103818846f, -0.00014644312103818846f));
    r = __ffma2_rn(r, u, make_float2(-0.0002300897957034123f, -0.0002300897957034123f));
    r = __ffma2_rn(r, u, make_float2(0.007180221614911477f, 0.007180221614911477f));
    r = __ffma2_rn(r, u, make_float2(-0.052572273431757154f, -0.052572273431757154f));
    r = __ffma2_rn(r, u, make_float2(-0.4591854841684466f, -0.4591854841684466f));
    r = __ffma2_rn(r, u, make_float2(-1.1511073739593443f, -1.1511073739593443f));
    const float2 t = __ffma2_rn(r, u, make_float2(-1.0f, -1.0f));
    float2 e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(t.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(t.y));
    return __ffma2_rn(na, e, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
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
int fc_prof_open(int cls, double flops, double bytes, cudaStream_t s, long long tag = 0);   // returns the record's id
void fc_prof_close(int id, cudaStream_t s);
struct FcProfScope {
    cudaStream_t s; bool on; int id = -1;
    FcProfScope(int cls, double flops, double bytes, cudaStream_t st, long long tag = 0) : s(st), on(fc_prof_enabled()) {
        if (on) id = fc_prof_open(cls, flops, bytes, s, tag);
    }
    ~FcProfScope() { if (on) fc_prof_close(id, s); }
};
