// Unit peaks of the box, measured with CUDA events (the denominators DESIGN.md quotes next to MEASURED_PEAKS.json):
//   fp32 FFMA and packed FFMA2 rate (all SMs, 8 independent chains per thread),
//   tcgen05.mma kind::f16 / kind::tf32 dense rate with the A operand in tensor memory (TS form, M = 128, N = 96 / 256).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/peaks scripts/micro/peaks.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024, 1) ffma_kernel(float* out, int iters, float a, float b) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;
}
__global__ void __launch_bounds__(1024, 1) ffma2_kernel(float* out, int iters, float a, float b) {
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __ffma2_rn(v[i], a2, b2);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i].x + v[i].y;
    if (s == 123.456f) out[0] = s;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    uint64_t d = 0;
    d |= (uint64_t)((a & 0x3FFFF) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
    return d;
}
// kind: 0 = f16 (K = 16 per instruction), 1 = tf32 (K = 8)
__global__ void __launch_bounds__(128, 1) mma_kernel(int N, int kind, int iters, int n_acc = 2, int ss = 0) {
    extern __shared__ unsigned char raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    unsigned char* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t fmt = kind == 0 ? 0u : 2u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t db = desc_sw128(smem_u32(sm));
        // operands of the four instructions of an unrolled iteration, computed once: the loop is nothing but the MMAs
        uint32_t dd[4], aa[4]; uint64_t bb[4], sa[4];
        for (int j = 0; j < 4; ++j) {
            dd[j] = tmem + (uint32_t)((j % n_acc) * N); aa[j] = tmem + 448 + (uint32_t)(j * 8);
            bb[j] = db + (uint64_t)(j * 2); sa[j] = desc_sw128(smem_u32(sm + 16384)) + (uint64_t)(j * 2);
        }
        for (int i = 0; i < iters; i += 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (ss)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(dd[j]), "l"(sa[j]), "l"(bb[j]), "r"(idesc), "r"(1u) : "memory");
                else if (kind == 0)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(dd[j]), "r"(aa[j]), "l"(bb[j]), "r"(idesc), "r"(1u) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(dd[j]), "r"(aa[j]), "l"(bb[j]), "r"(idesc), "r"(1u) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <class F> static float time_ms(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0.f; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}
int main() {
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    float* d; cudaMalloc(&d, 16);
    {
        const int iters = 4000;
        const double fl = 2.0 * 64 * iters * 1024.0 * 2 * nsm;   // 2 CTAs of 1024 threads per SM
        float ms = time_ms([&] { ffma_kernel<<<2 * nsm, 1024>>>(d, iters, 1.0000001f, 1e-9f); }, 10);
        printf("fp32 FFMA  : %.1f TFLOP/s (%d SMs, %.3f ms)\n", fl / ms / 1e9, nsm, ms);
        ms = time_ms([&] { ffma2_kernel<<<2 * nsm, 1024>>>(d, iters, 1.0000001f, 1e-9f); }, 10);
        printf("fp32 FFMA2 : %.1f TFLOP/s (packed, 2 fma per instruction)\n", 2 * fl / ms / 1e9);
    }
    cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    // cycles per kind::f16 instruction against N and the number of accumulators the instructions rotate through
    // (consecutive MMAs into the SAME accumulator are dependent)
    {
        int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
        for (int ss = 0; ss < 2; ++ss)
            for (int N : {64, 96, 128, 192, 256})
                for (int n_acc : {1, 2, 4}) {
                    if (n_acc * N > 448) continue;
                    const int iters = 50000;
                    float ms = time_ms([&] { mma_kernel<<<nsm, 128, 40 * 1024>>>(N, 0, iters, n_acc, ss); }, 2);
                    cudaError_t e = cudaDeviceSynchronize();
                    printf("kind::f16 %s M=128 N=%3d accumulators=%d : %.1f cycles per MMA at %d MHz (N/2 = %d), %.0f TFLOP/s %s\n", ss ? "SS" : "TS", N, n_acc,
                           ms * 1e-3 * khz * 1e3 / iters, khz / 1000, N / 2, 2.0 * 128 * N * 16 * iters * nsm / ms / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
                }
    }
    for (int kind = 0; kind < 2; ++kind)
        for (int N : {96, 256}) {
            const int iters = 200000;
            const double fl = 2.0 * 128 * N * (kind == 0 ? 16 : 8) * (double)iters * nsm;
            float ms = time_ms([&] { mma_kernel<<<nsm, 128, 40 * 1024>>>(N, kind, iters, N <= 192 ? 2 : 1, 0); }, 3);
            cudaError_t e = cudaDeviceSynchronize();
            printf("tcgen05.mma kind::%s TS M=128 N=%3d : %.1f TFLOP/s dense (%.1f ms) %s\n", kind == 0 ? "f16 " : "tf32", N, fl / ms / 1e9, ms,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
