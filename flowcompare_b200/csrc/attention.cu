// Single-head cross attention, flash style (online softmax over key tiles), fp32.
//
// Replaces reference models/perceiver.py:108-115 (`AttentionMine.forward`), which materialises the
// [B, N, Nc] score matrix (5 MB per cloud per layer, x116 layers).  Here a CTA owns 64 queries of one
// cloud, streams 64-key tiles of K and V through shared memory and keeps the running max / sum /
// output accumulator in registers; nothing of size N*Nc touches HBM.
#include <atomic>
#include "common.cuh"
#include "gemm.cuh"

namespace {

constexpr int AD = 64;       // head dim (reference cross_dim_head * cross_heads = 64, single head)
constexpr int AQ = 64;       // queries per CTA
constexpr int AKT = 64;      // keys per tile
constexpr int ATHREADS = 256;
constexpr int ALD = AQ + 4;  // padded leading dim of the transposed tiles

struct AttnSmem {
    float Qt[AD][ALD];   // Qt[d][q]  (pre-scaled by `scale`)
    float Kt[AD][ALD];   // Kt[d][key]
    float V[AKT][AD];    // V[key][d]
    float Pt[AKT][ALD];  // Pt[key][q]
};

__global__ void __launch_bounds__(ATHREADS) cross_attention_kernel(const float* __restrict__ q, int ldq,
                                                                   const float* __restrict__ kv, int ldkv,
                                                                   float* __restrict__ out, int ldo,
                                                                   int N, int Nc, float scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AttnSmem& sm = *reinterpret_cast<AttnSmem*>(smem_raw);

    const int tid = threadIdx.x;
    const int tk = tid & 15;   // column group: keys tk*4.. (S) / dims tk*4.. (O)
    const int tq = tid >> 4;   // row group: queries tq*4..
    const int b = blockIdx.y;
    const int q0 = blockIdx.x * AQ;
    const float* qb = q + ((size_t)b * N) * ldq;
    const float* kvb = kv + ((size_t)b * Nc) * ldkv;

    // stage Q transposed; the reference multiplies the scores by `scale` after the matmul -- scaling
    // the dot product afterwards keeps the same rounding, so Q is stored unscaled.
    for (int e = tid; e < AQ * (AD / 4); e += ATHREADS) {
        const int r = e / (AD / 4), c4 = (e % (AD / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < N) v = *reinterpret_cast<const float4*>(qb + (size_t)(q0 + r) * ldq + c4);
        sm.Qt[c4 + 0][r] = v.x; sm.Qt[c4 + 1][r] = v.y; sm.Qt[c4 + 2][r] = v.z; sm.Qt[c4 + 3][r] = v.w;
    }

    float o[4][4];
    float mrow[4], lrow[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        mrow[i] = -INFINITY; lrow[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    }

    for (int k0 = 0; k0 < Nc; k0 += AKT) {
        __syncthreads();  // previous tile fully consumed (also covers the Q staging on the first pass)
        for (int e = tid; e < AKT * (2 * AD / 4); e += ATHREADS) {
            const int r = e / (2 * AD / 4), c4 = (e % (2 * AD / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + r < Nc) v = *reinterpret_cast<const float4*>(kvb + (size_t)(k0 + r) * ldkv + c4);
            if (c4 < AD) {
                sm.Kt[c4 + 0][r] = v.x; sm.Kt[c4 + 1][r] = v.y; sm.Kt[c4 + 2][r] = v.z; sm.Kt[c4 + 3][r] = v.w;
            } else {
                *reinterpret_cast<float4*>(&sm.V[r][c4 - AD]) = v;
            }
        }
        __syncthreads();

        // S = Q K^T for rows tq*4.., keys tk*4..
        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 16
        for (int d = 0; d < AD; ++d) {
            const float4 a = *reinterpret_cast<const float4*>(&sm.Qt[d][tq * 4]);
            const float4 kk = *reinterpret_cast<const float4*>(&sm.Kt[d][tk * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], bv[j], s[i][j]);
        }
        // online softmax
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mt = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s[i][j] = (k0 + tk * 4 + j < Nc) ? s[i][j] * scale : -INFINITY;
                mt = fmaxf(mt, s[i][j]);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, off));
            const float mnew = fmaxf(mrow[i], mt);
            const float alpha = expf(mrow[i] - mnew);  // exp(-inf) = 0 on the first tile
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float p = expf(s[i][j] - mnew);
                s[i][j] = p;
                ps += p;
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
            lrow[i] = lrow[i] * alpha + ps;
            mrow[i] = mnew;
#pragma unroll
            for (int j = 0; j < 4; ++j) o[i][j] *= alpha;
        }
        // P (transposed) to shared
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(&sm.Pt[tk * 4 + j][tq * 4]) = make_float4(s[0][j], s[1][j], s[2][j], s[3][j]);
        __syncthreads();
        // O += P V for rows tq*4.., dims tk*4..
#pragma unroll 16
        for (int kk = 0; kk < AKT; ++kk) {
            const float4 p = *reinterpret_cast<const float4*>(&sm.Pt[kk][tq * 4]);
            const float4 v = *reinterpret_cast<const float4*>(&sm.V[kk][tk * 4]);
            const float pv[4] = {p.x, p.y, p.z, p.w};
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) o[i][j] = fmaf(pv[i], vv[j], o[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = q0 + tq * 4 + i;
        if (r >= N) continue;
        const float inv = 1.0f / lrow[i];
        float* dst = out + ((size_t)b * N + r) * ldo + tk * 4;
        *reinterpret_cast<float4*>(dst) = make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
    }
}

}  // namespace

int fc_launch_cross_attention(const float* q, int ldq, const float* kv, int ldkv, float* out, int ldo,
                              int B, int N, int Nc, int d, float scale, cudaStream_t stream) {
    FC_REQUIRE(q && kv && out && B > 0 && N > 0 && Nc > 0);
    if (d != AD) return FC_ERR_UNSUPPORTED;
    FC_REQUIRE((ldq & 3) == 0 && (ldkv & 3) == 0 && (ldo & 3) == 0 && ldkv >= 2 * AD);
    FC_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(kv) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
    FC_REQUIRE(B <= 65535);
    int dev = 0;
    FC_CUDA_OK(cudaGetDevice(&dev));
    static std::atomic<uint64_t> configured{0};     // the dynamic shared-memory opt-in is per device
    if (!(configured.load(std::memory_order_acquire) & (1ull << (dev & 63)))) {
        FC_CUDA_OK(cudaFuncSetAttribute(cross_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(AttnSmem)));
        configured.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    dim3 grid((N + AQ - 1) / AQ, B);
    FcProfScope prof(FC_CLS_ATTENTION, 4.0 * B * (double)N * Nc * d, 4.0 * B * ((double)N * d * 2 + (double)Nc * d * 2), stream);
    cross_attention_kernel<<<grid, ATHREADS, sizeof(AttnSmem), stream>>>(q, ldq, kv, ldkv, out, ldo, N, Nc, scale);
    fc_count_launch();
    FC_LAUNCH_OK();
    return FC_OK;
}

extern "C" int fc_cross_attention(const float* q, int ldq, const float* kv, int ldkv, float* out, int ldo,
                                  int B, int N, int Nc, int d, float scale, fc_stream_t stream) {
    return fc_launch_cross_attention(q, ldq, kv, ldkv, out, ldo, B, N, Nc, d, scale, (cudaStream_t)stream);
}
