// Microbenchmark: cycles per tcgen05.mma.kind::tf32 (M=128, N variable), A from shared memory (SS) or TMEM (TS).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/mma_bench scripts/micro/mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    uint64_t d = 0;
    d |= (uint64_t)((a & 0x3FFFF) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
    return d;
}
__global__ void __launch_bounds__(128, 1) bench(int N, int mode, int iters, int n_acc, long long* out) {
    extern __shared__ unsigned char raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    unsigned char* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
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
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = desc_sw128(smem_u32(sm)), db = desc_sw128(smem_u32(sm + 16384));
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem + (uint32_t)((i % n_acc) * N);
            const uint64_t adv = (uint64_t)((i & 3) * 2);
            if (mode == 0) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(d), "l"(da + adv), "l"(db + adv), "r"(idesc), "r"(1u) : "memory");
            } else {
                const uint32_t a_t = tmem + 448 + (uint32_t)((i & 3) * 8);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                             ::"r"(d), "r"(a_t), "l"(db + adv), "r"(idesc), "r"(1u) : "memory");
            }
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 2000;
    for (int grid : {1, 148})
        for (int mode = 0; mode < 2; ++mode)
            for (int N : {64, 96, 128, 176, 192, 256})
                for (int n_acc : {1, 2}) {
                    if (n_acc * N > 448) continue;
                    bench<<<grid, 128, 60 * 1024>>>(N, mode, iters, n_acc, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                    printf("grid %3d %s N=%3d acc=%d : issue %.1f cyc/mma, complete %.1f cyc/mma (floor N/2 = %d) %s\n", grid, mode ? "TS" : "SS", N, n_acc,
                           (double)h[0] / iters, (double)h[1] / iters, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
                }
    return 0;
}
