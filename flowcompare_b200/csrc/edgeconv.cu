// EdgeConv neighbour aggregation: out_i = LeakyReLU_0.2( max_{j in idx_i} P_j + Q_i ).
//
// The reference (models/pytorch_gcn.py:23-47, :63-74, :86-100) gathers neighbours into a
// materialised [B, 2C, N, k] tensor (51 MB per cloud at C=128), runs a 1x1 conv + BN + LeakyReLU over
// all N*k edges and then takes the max over k.  Because the conv is linear in [x_j - x_i ; x_i],
// BN(eval) is a per-channel affine map and LeakyReLU is monotone, the same value is
// LeakyReLU(max_j P_j + Q_i) with P = x (a.W1)^T and Q = x (a.(W2-W1))^T + b computed once per POINT
// (k times fewer flops).  This kernel is the gather-max part: pure bandwidth, coalesced float4 row
// loads of P (rows are 256 B - 1 KB contiguous), one warp-slice per point.
#include "common.cuh"
#include "gemm.cuh"

namespace {

constexpr int EC_THREADS = 256;

template <int LANES>  // float4 lanes per point = Cout/4 (16, 32 or 64)
__global__ void __launch_bounds__(EC_THREADS) edgeconv_gather_max_kernel(const float* __restrict__ PQ, int ldpq,
                                                                         const int32_t* __restrict__ idx,
                                                                         long long total_points, int N, int k,
                                                                         int Cout, float* __restrict__ out, int ldo) {
    constexpr int PTS = EC_THREADS / LANES;
    const int lane4 = threadIdx.x % LANES;
    const int pl = threadIdx.x / LANES;
    const long long p = (long long)blockIdx.x * PTS + pl;
    if (p >= total_points) return;
    const long long cloud_base = (p / N) * N;  // first row of this point's cloud
    const int32_t* my_idx = idx + p * k;
    const float* Pbase = PQ + lane4 * 4;

    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    int j = 0;
    for (; j + 4 <= k; j += 4) {
        const int i0 = my_idx[j], i1 = my_idx[j + 1], i2 = my_idx[j + 2], i3 = my_idx[j + 3];
        const float4 v0 = *reinterpret_cast<const float4*>(Pbase + (size_t)(cloud_base + i0) * ldpq);
        const float4 v1 = *reinterpret_cast<const float4*>(Pbase + (size_t)(cloud_base + i1) * ldpq);
        const float4 v2 = *reinterpret_cast<const float4*>(Pbase + (size_t)(cloud_base + i2) * ldpq);
        const float4 v3 = *reinterpret_cast<const float4*>(Pbase + (size_t)(cloud_base + i3) * ldpq);
        m.x = fmaxf(fmaxf(fmaxf(m.x, v0.x), fmaxf(v1.x, v2.x)), v3.x);
        m.y = fmaxf(fmaxf(fmaxf(m.y, v0.y), fmaxf(v1.y, v2.y)), v3.y);
        m.z = fmaxf(fmaxf(fmaxf(m.z, v0.z), fmaxf(v1.z, v2.z)), v3.z);
        m.w = fmaxf(fmaxf(fmaxf(m.w, v0.w), fmaxf(v1.w, v2.w)), v3.w);
    }
    for (; j < k; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(Pbase + (size_t)(cloud_base + my_idx[j]) * ldpq);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
    const float4 qv = *reinterpret_cast<const float4*>(PQ + (size_t)p * ldpq + Cout + lane4 * 4);
    float4 r;
    r.x = fc_leaky_relu02(m.x + qv.x);
    r.y = fc_leaky_relu02(m.y + qv.y);
    r.z = fc_leaky_relu02(m.z + qv.z);
    r.w = fc_leaky_relu02(m.w + qv.w);
    *reinterpret_cast<float4*>(out + (size_t)p * ldo + lane4 * 4) = r;
}

}  // namespace

int fc_launch_edgeconv_gather_max(const float* PQ, int ldpq, const int32_t* idx, int B, int N, int k, int Cout,
                                  float* out, int ldo, cudaStream_t stream) {
    FC_REQUIRE(PQ && idx && out && B > 0 && N > 0 && k > 0);
    FC_REQUIRE((ldpq & 3) == 0 && (ldo & 3) == 0 && ldpq >= 2 * Cout && ldo >= Cout);
    FC_REQUIRE(((reinterpret_cast<uintptr_t>(PQ) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
    const long long total = (long long)B * N;
    auto blocks = [&](int lanes) { return (unsigned)((total + (EC_THREADS / lanes) - 1) / (EC_THREADS / lanes)); };
    FcProfScope prof(FC_CLS_EDGECONV, (double)total * k * Cout,
                     4.0 * ((double)total * Cout * 3 + (double)total * k), stream);
    switch (Cout) {
        case 64:  edgeconv_gather_max_kernel<16><<<blocks(16), EC_THREADS, 0, stream>>>(PQ, ldpq, idx, total, N, k, Cout, out, ldo); break;
        case 128: edgeconv_gather_max_kernel<32><<<blocks(32), EC_THREADS, 0, stream>>>(PQ, ldpq, idx, total, N, k, Cout, out, ldo); break;
        case 256: edgeconv_gather_max_kernel<64><<<blocks(64), EC_THREADS, 0, stream>>>(PQ, ldpq, idx, total, N, k, Cout, out, ldo); break;
        default: return FC_ERR_UNSUPPORTED;
    }
    fc_count_launch();
    FC_LAUNCH_OK();
    return FC_OK;
}

extern "C" int fc_edgeconv_gather_max(const float* PQ, int ldpq, const int32_t* idx, int B, int N, int k, int Cout,
                                      float* out, int ldo, fc_stream_t stream) {
    return fc_launch_edgeconv_gather_max(PQ, ldpq, idx, B, N, k, Cout, out, ldo, (cudaStream_t)stream);
}
