// fp32 FFMA GEMM with fused epilogues (see gemm.cuh).  128x128x16 (or 128x64x16) CTA tile,
// 256 threads, 8x8 (8x4) register tile, double-buffered shared memory, one barrier per k-tile.
// The tcgen05 3xTF32 variant lives in gemm_tc.cu; this kernel is the exact-fp32 path, the
// reference for the tensor-core one, and the path for small / odd shapes.
#include "gemm.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int NTHREADS = 256;
constexpr int AS_LD = BM + 4;

template <int BN, bool VEC_A>
__global__ void __launch_bounds__(NTHREADS, 2) gemm_ffma_kernel(const GemmArgs a) {
    constexpr int TN = BN / 16;          // columns per thread: 8 (BN=128) or 4 (BN=64)
    constexpr int NG = TN / 4;           // column groups of 4
    __shared__ __align__(16) float As[2][BK][AS_LD];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;

    const int Kp1 = (a.K1 + BK - 1) / BK * BK;
    const int T1 = Kp1 / BK;
    const int T2 = (a.K2 + BK - 1) / BK;
    const int T = T1 + T2;

    // A-load mapping: two passes of 64 rows, 4 k-quads per row
    const int a_row = tid >> 2;          // 0..63
    const int a_kq = (tid & 3) * 4;      // 0,4,8,12
    // B-load mapping
    const int b_k = tid / (BN / 4);      // rows per pass = 256/(BN/4): 8 (BN=128) or 16 (BN=64)
    const int b_n = (tid % (BN / 4)) * 4;
    constexpr int B_PASSES = (BK * BN / 4) / NTHREADS;  // 2 or 1
    constexpr int B_ROWS_PER_PASS = NTHREADS / (BN / 4);

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[2];
    float4 rb[B_PASSES];

    auto load_tile = [&](int t) {
        const float* A; int lda, K, kbase;
        if (t < T1) { A = a.A1; lda = a.lda1; K = a.K1; kbase = t * BK; }
        else        { A = a.A2; lda = a.lda2; K = a.K2; kbase = (t - T1) * BK; }
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int row = m0 + a_row + p * 64;
            const int k = kbase + a_kq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < a.M) {
                const float* src = A + (size_t)row * lda + k;
                if (VEC_A && k + 3 < K) {
                    v = *reinterpret_cast<const float4*>(src);
                } else {
                    if (k + 0 < K) v.x = src[0];
                    if (k + 1 < K) v.y = src[1];
                    if (k + 2 < K) v.z = src[2];
                    if (k + 3 < K) v.w = src[3];
                }
            }
            ra[p] = v;
        }
        const int wrow0 = t * BK;  // Wt rows are laid out tile after tile (Kp1 is a multiple of BK)
#pragma unroll
        for (int p = 0; p < B_PASSES; ++p) {
            const int k = b_k + p * B_ROWS_PER_PASS;
            rb[p] = *reinterpret_cast<const float4*>(a.Wt + (size_t)(wrow0 + k) * a.ldw + n0 + b_n);
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int r = a_row + p * 64;
            As[buf][a_kq + 0][r] = ra[p].x;
            As[buf][a_kq + 1][r] = ra[p].y;
            As[buf][a_kq + 2][r] = ra[p].z;
            As[buf][a_kq + 3][r] = ra[p].w;
        }
#pragma unroll
        for (int p = 0; p < B_PASSES; ++p) {
            const int k = b_k + p * B_ROWS_PER_PASS;
            *reinterpret_cast<float4*>(&Bs[buf][k][b_n]) = rb[p];
        }
    };

    load_tile(0);
    store_tile(0);
    __syncthreads();

    for (int t = 0; t < T; ++t) {
        const int buf = t & 1;
        if (t + 1 < T) load_tile(t + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float av[8], bv[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
            av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][g * 64 + tx * 4]);
                bv[g * 4 + 0] = b.x; bv[g * 4 + 1] = b.y; bv[g * 4 + 2] = b.z; bv[g * 4 + 3] = b.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (t + 1 < T) {
            store_tile(buf ^ 1);
            __syncthreads();
        }
    }

    // ------------------------------------------------------------------ epilogue
    const int epi = a.epi;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        const bool row_ok = row < a.M;
        float ldj = 0.f;
        float mu = 0.f, rstd = 0.f;
        if (epi == FC_EPI_LNQ && row_ok) { mu = a.row_mu[row]; rstd = a.row_rstd[row]; }
        const float* bias_row = a.bias;
        if (a.bias && a.bias_group > 0 && row_ok) bias_row = a.bias + (size_t)(row / a.bias_group) * a.bias_ld;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int col = n0 + g * 64 + tx * 4;
            if (!row_ok || col >= a.N) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = acc[i][g * 4 + j];
            const bool full = (col + 3 < a.N);
            if (epi == FC_EPI_LNQ) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (col + j < a.N) v[j] = rstd * (v[j] - mu * a.csum[col + j]) + a.bias[col + j];
            } else {
                if (bias_row) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (col + j < a.N) v[j] += bias_row[col + j];
                }
            }
            if (epi == FC_EPI_STORE || epi == FC_EPI_LNQ) {
                if (a.res) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (col + j < a.N) {
                            const float rv = a.res[(size_t)row * a.ldres + col + j];
                            v[j] = a.res_scale ? fmaf(a.res_scale[col + j], rv, v[j]) : v[j] + rv;
                        }
                }
                if (a.act == FC_ACT_GELU) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = fc_gelu_erf(v[j]);
                } else if (a.act == FC_ACT_LRELU) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = fc_leaky_relu02(v[j]);
                } else if (a.act == FC_ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                float* dst = a.C + (size_t)row * a.ldc + col;
                if (full && ((a.ldc & 3) == 0)) {
                    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (col + j < a.N) dst[j] = v[j];
                }
            } else if (epi == FC_EPI_COUPLING) {
                // reference models/affine_coupling.py:40-46, sigmoid scale with eps=1e-8:
                // s = (2*sigmoid(s_raw) - 1) * (1 - 1e-8) + 1 ; (1 - 1e-8) == 1.0f in fp32
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    if (col + 2 * p + 1 < a.N) {
                        const int j = (col >> 1) + p;
                        const float sraw = v[2 * p], tt = v[2 * p + 1];
                        const float sig = 1.0f / (1.0f + expf(-sraw));
                        const float s = (2.0f * sig - 1.0f) + 1.0f;
                        float* xp = a.x + (size_t)row * a.ldx + a.col0 + j;
                        *xp = fmaf(*xp, s, tt);
                        ldj += logf(s);
                    }
                }
            } else if (epi == FC_EPI_COUPLING_INV) {
                // reference models/affine_coupling.py:48-62: x2 = (y2 - t) / s with the same sigmoid scale
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    if (col + 2 * p + 1 < a.N) {
                        const int j = (col >> 1) + p;
                        const float sraw = v[2 * p], tt = v[2 * p + 1];
                        const float sig = 1.0f / (1.0f + expf(-sraw));
                        const float s = (2.0f * sig - 1.0f) + 1.0f;
                        float* xp = a.x + (size_t)row * a.ldx + a.col0 + j;
                        *xp = (*xp - tt) / s;
                    }
                }
            } else if (epi == FC_EPI_AUGMENT) {
                // reference models/distributions.py:128-153 + models/augmenter.py:49-63
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    if (col + 2 * p + 1 < a.N) {
                        const int j = (col >> 1) + p;
                        const float mean = v[2 * p], log_std = v[2 * p + 1];
                        const float e = a.eps[(size_t)row * a.ld_eps + j];
                        a.x[(size_t)row * a.ldx + a.col0 + j] = fmaf(expf(log_std), e, mean);
                        ldj += fmaf(0.5f * e, e, log_std) + 0.91893853320467274178f;
                    }
                }
            }
        }
        if (epi == FC_EPI_COUPLING || epi == FC_EPI_AUGMENT) {
            // reduce over the 16 threads (tx) that share this row: lanes (ty&1)*16 + tx
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) ldj += __shfl_xor_sync(0xffffffffu, ldj, o);
            if (tx == 0 && row_ok) {
                float* pp = a.part + (size_t)blockIdx.x * a.M + row;
                if (epi == FC_EPI_COUPLING) *pp += ldj; else *pp = ldj;
            }
        }
    }
}

template <int BN>
int launch(const GemmArgs& a, cudaStream_t stream) {
    dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM);
    FcProfScope prof(FC_CLS_GEMM_FFMA, 2.0 * a.M * a.N * (a.K1 + a.K2),
                     4.0 * ((double)a.M * (a.K1 + a.K2) + (double)a.N * (a.K1 + a.K2) + (double)a.M * a.N), stream);
    const bool vec = ((a.lda1 & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.A1) & 15) == 0) &&
                     (a.K2 == 0 || (((a.lda2 & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.A2) & 15) == 0)));
    if (vec) gemm_ffma_kernel<BN, true><<<grid, NTHREADS, 0, stream>>>(a);
    else     gemm_ffma_kernel<BN, false><<<grid, NTHREADS, 0, stream>>>(a);
    fc_count_launch();
    FC_LAUNCH_OK();
    return FC_OK;
}

}  // namespace

int fc_gemm_n_tiles(int N) { return N <= 64 ? 1 : (N + 127) / 128; }

int fc_launch_gemm_ffma(const GemmArgs& a, cudaStream_t stream) {
    FC_REQUIRE(a.M > 0 && a.N > 0 && a.K1 > 0 && a.A1 && a.Wt);
    if (a.epi == FC_EPI_KVSPLIT) return FC_ERR_UNSUPPORTED;   // tcgen05 path only
    FC_REQUIRE(a.K2 == 0 || a.A2);
    FC_REQUIRE(a.ldw >= fc_gemm_ldw(a.N) && (a.ldw & 3) == 0);
    FC_REQUIRE((reinterpret_cast<uintptr_t>(a.Wt) & 15) == 0);
    if (a.epi == FC_EPI_STORE || a.epi == FC_EPI_LNQ) FC_REQUIRE(a.C != nullptr);
    if (a.epi == FC_EPI_LNQ) FC_REQUIRE(a.row_mu && a.row_rstd && a.csum && a.bias && a.bias_group == 0);
    if (a.epi == FC_EPI_COUPLING || a.epi == FC_EPI_AUGMENT) FC_REQUIRE(a.x && a.part && (a.N % 2) == 0);
    if (a.epi == FC_EPI_COUPLING_INV) FC_REQUIRE(a.x && (a.N % 2) == 0);
    if (a.epi == FC_EPI_AUGMENT) FC_REQUIRE(a.eps != nullptr);
    if (a.N <= 64) return launch<64>(a, stream);
    return launch<128>(a, stream);
}
