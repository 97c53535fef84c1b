// Single-head cross attention on the tensor cores: flash style, 3xTF32 error-compensated warp-level MMA
// (mma.sync.m16n8k8.tf32, fp32 accumulate).  Replaces reference models/perceiver.py:108-115 like attention.cu,
// which stays as the exact-fp32 path; this one is used with precision = tf32x3.
//
// CTA = 64 queries of one cloud (4 warps x 16 rows); 64-key tiles of K and V are double-buffered in shared
// memory with cp.async.  S = Q K^T and O += P V are both done as  hi*hi + lo*hi + hi*lo  on TF32 splits made in
// registers, so scores and outputs keep fp32 accuracy (tests: <= 1e-5 abs on softmax outputs of |scores| ~ 8).
// P never leaves registers: the accumulator fragment of S is re-used as the A fragment of P V by permuting the
// key order inside each 8-key step (k-slot t <-> key 2t, slot t+4 <-> key 2t+1) and reading V with the same
// permutation -- no shuffles, no shared-memory round trip.
// (A tcgen05/TMEM version is the next step; the GEMMs, 3/4 of the step, went first.)
#include <atomic>
#include "common.cuh"
#include "gemm.cuh"

namespace {

constexpr int FD = 64;      // head dim
constexpr int FQ = 64;      // queries per CTA
constexpr int FK = 64;      // keys per tile
constexpr int FLD = 68;     // padded row (floats): conflict-free fragment reads for both K and V
constexpr int FTHREADS = 128;

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(FTHREADS, 3)
cross_attention_mma_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ kv, int ldkv,
                           float* __restrict__ out, int ldo, int N, int Nc, float scale) {
    extern __shared__ __align__(16) float fsm[];
    float* Ks = fsm;                          // [2][FK][FLD]
    float* Vs = fsm + 2 * FK * FLD;           // [2][FK][FLD]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.y;
    const int q0 = blockIdx.x * FQ + warp * 16;
    const float* qb = q + ((size_t)b * N) * ldq;
    const float* kvb = kv + ((size_t)b * Nc) * ldkv;

    auto load_tile = [&](int kt, int buf) {
        const int k0 = kt * FK;
#pragma unroll
        for (int i = 0; i < (FK * 2 * FD / 4) / FTHREADS; ++i) {
            const int idx = tid + i * FTHREADS;
            const int row = idx >> 5, c4 = idx & 31;
            const bool ok = k0 + row < Nc;
            const float* src = kvb + (size_t)(ok ? k0 + row : 0) * ldkv + c4 * 4;
            float* dst = (c4 < 16 ? Ks + (buf * FK + row) * FLD + c4 * 4 : Vs + (buf * FK + row) * FLD + (c4 - 16) * 4);
            cp_async16(dst, src, ok ? 16 : 0);     // zero fill beyond the cloud
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int ntiles = (Nc + FK - 1) / FK;
    load_tile(0, 0);

    // Q fragments (A operand, rows g / g+8 of this warp's 16), split once
    uint32_t qh[8][4], ql[8][4];
    {
        const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const float a0 = r0 < N ? qb[(size_t)r0 * ldq + 8 * ks + t] : 0.f;
            const float a1 = r1 < N ? qb[(size_t)r1 * ldq + 8 * ks + t] : 0.f;
            const float a2 = r0 < N ? qb[(size_t)r0 * ldq + 8 * ks + t + 4] : 0.f;
            const float a3 = r1 < N ? qb[(size_t)r1 * ldq + 8 * ks + t + 4] : 0.f;
            split_tf32(a0, qh[ks][0], ql[ks][0]); split_tf32(a1, qh[ks][1], ql[ks][1]);
            split_tf32(a2, qh[ks][2], ql[ks][2]); split_tf32(a3, qh[ks][3], ql[ks][3]);
        }
    }

    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // rows g and g+8 (l: this thread's partial sum)

    for (int kt = 0; kt < ntiles; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < ntiles) { load_tile(kt + 1, buf ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const float* Kt = Ks + buf * FK * FLD;
        const float* Vt = Vs + buf * FK * FLD;

        // ---- S = Q K^T : n-tile j = keys 8j..8j+7, k-step ks = dims 8ks..8ks+7
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
            const float* krow = Kt + (8 * j + g) * FLD + t;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                uint32_t bh0, bl0, bh1, bl1;
                split_tf32(krow[8 * ks], bh0, bl0);
                split_tf32(krow[8 * ks + 4], bh1, bl1);
                mma_tf32(s[j], ql[ks], bh0, bh1);
                mma_tf32(s[j], qh[ks], bl0, bl1);
                mma_tf32(s[j], qh[ks], bh0, bh1);
            }
        }
        // ---- online softmax (rows g: s[j][0..1]; rows g+8: s[j][2..3]; columns k0 + 8j + 2t + {0,1})
        const int kbase = kt * FK + 2 * t;
        float mt0 = -INFINITY, mt1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool v0 = kbase + 8 * j < Nc, v1 = kbase + 8 * j + 1 < Nc;
            s[j][0] = v0 ? s[j][0] * scale : -INFINITY; s[j][1] = v1 ? s[j][1] * scale : -INFINITY;
            s[j][2] = v0 ? s[j][2] * scale : -INFINITY; s[j][3] = v1 ? s[j][3] * scale : -INFINITY;
            mt0 = fmaxf(mt0, fmaxf(s[j][0], s[j][1])); mt1 = fmaxf(mt1, fmaxf(s[j][2], s[j][3]));
        }
        mt0 = fmaxf(mt0, __shfl_xor_sync(0xffffffffu, mt0, 1)); mt0 = fmaxf(mt0, __shfl_xor_sync(0xffffffffu, mt0, 2));
        mt1 = fmaxf(mt1, __shfl_xor_sync(0xffffffffu, mt1, 1)); mt1 = fmaxf(mt1, __shfl_xor_sync(0xffffffffu, mt1, 2));
        const float mn0 = fmaxf(m0, mt0), mn1 = fmaxf(m1, mt1);
        const float al0 = expf(m0 - mn0), al1 = expf(m1 - mn1);   // exp(-inf) = 0 on the first tile
        m0 = mn0; m1 = mn1;
        float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = expf(s[j][0] - mn0); s[j][1] = expf(s[j][1] - mn0);
            s[j][2] = expf(s[j][2] - mn1); s[j][3] = expf(s[j][3] - mn1);
            ps0 += s[j][0] + s[j][1]; ps1 += s[j][2] + s[j][3];
        }
        l0 = l0 * al0 + ps0; l1 = l1 * al1 + ps1;
#pragma unroll
        for (int j = 0; j < 8; ++j) { o[j][0] *= al0; o[j][1] *= al0; o[j][2] *= al1; o[j][3] *= al1; }

        // ---- O += P V : k-step ks = keys 8ks..8ks+7 (slot t <-> key 8ks+2t, slot t+4 <-> key 8ks+2t+1), n-tile j = dims 8j..8j+7
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            uint32_t ph[4], pl[4];
            split_tf32(s[ks][0], ph[0], pl[0]);   // a0: row g,   slot t   = key 2t
            split_tf32(s[ks][2], ph[1], pl[1]);   // a1: row g+8, slot t
            split_tf32(s[ks][1], ph[2], pl[2]);   // a2: row g,   slot t+4 = key 2t+1
            split_tf32(s[ks][3], ph[3], pl[3]);   // a3: row g+8, slot t+4
            const float* v0 = Vt + (8 * ks + 2 * t) * FLD + g;
            const float* v1 = v0 + FLD;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t bh0, bl0, bh1, bl1;
                split_tf32(v0[8 * j], bh0, bl0);
                split_tf32(v1[8 * j], bh1, bl1);
                mma_tf32(o[j], pl, bh0, bh1);
                mma_tf32(o[j], ph, bl0, bl1);
                mma_tf32(o[j], ph, bh0, bh1);
            }
        }
        __syncthreads();   // everyone is done with `buf` before the next iteration's prefetch overwrites it
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (r0 < N) *reinterpret_cast<float2*>(out + ((size_t)b * N + r0) * ldo + 8 * j + 2 * t) = make_float2(o[j][0] * i0, o[j][1] * i0);
        if (r1 < N) *reinterpret_cast<float2*>(out + ((size_t)b * N + r1) * ldo + 8 * j + 2 * t) = make_float2(o[j][2] * i1, o[j][3] * i1);
    }
}

}  // namespace

int fc_launch_cross_attention_mma(const float* q, int ldq, const float* kv, int ldkv, float* out, int ldo,
                                  int B, int N, int Nc, int d, float scale, cudaStream_t stream) {
    FC_REQUIRE(q && kv && out && B > 0 && N > 0 && Nc > 0);
    if (d != FD) return FC_ERR_UNSUPPORTED;
    FC_REQUIRE((ldkv & 3) == 0 && (ldo & 1) == 0 && ldkv >= 2 * FD);
    FC_REQUIRE(((reinterpret_cast<uintptr_t>(kv) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0));
    FC_REQUIRE(B <= 65535);
    const int smem = 4 * FK * FLD * (int)sizeof(float);
    int dev = 0;
    FC_CUDA_OK(cudaGetDevice(&dev));
    static std::atomic<uint64_t> configured{0};     // the dynamic shared-memory opt-in is per device
    if (!(configured.load(std::memory_order_acquire) & (1ull << (dev & 63)))) {
        FC_CUDA_OK(cudaFuncSetAttribute(cross_attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    dim3 grid((N + FQ - 1) / FQ, B);
    FcProfScope prof(FC_CLS_ATTENTION, 4.0 * B * (double)N * Nc * d, 4.0 * B * ((double)N * d * 2 + (double)Nc * d * 2), stream);
    cross_attention_mma_kernel<<<grid, FTHREADS, smem, stream>>>(q, ldq, kv, ldkv, out, ldo, N, Nc, scale);
    fc_count_launch();
    FC_LAUNCH_OK();
    return FC_OK;
}

extern "C" __attribute__((visibility("default"))) int fc_cross_attention_tf32x3(const float* q, int ldq, const float* kv, int ldkv,
                                                                                 float* out, int ldo, int B, int N, int Nc, int d,
                                                                                 float scale, fc_stream_t stream) {
    return fc_launch_cross_attention_mma(q, ldq, kv, ldkv, out, ldo, B, N, Nc, d, scale, (cudaStream_t)stream);
}
