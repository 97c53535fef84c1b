// Brute-force k-nearest-neighbour kernels: shared-memory tiled distances, warp-level top-k.
//
// Replaces reference models/pytorch_gcn.py:13-20 (`knn`: bmm + materialised [B,N,N] matrix + topk)
// and knn.py:40-52 (`KNN_torch_fun`).  Nothing of size N^2 is ever written: a CTA owns 32 queries,
// streams 128-candidate tiles through shared memory, and every warp keeps the running top-k of its
// 4 queries as a sorted list spread over its lanes (2 slots per lane, k <= 64).
//
// Arithmetic is the reference's algebraic form with a FIXED evaluation order (sequential fmaf over
// the feature index), which oracle/knn_ref.c restates bit-for-bit:
//   mode 0 (DGCNN):  key_ij = ((-xx_j) - (-2*dot_ij)) - xx_i          (k largest)
//   mode 1 (knn.py): key_ij = -((qq_i + tt_j) - 2*dot_ij)             (k largest == k smallest diss)
// Ties are broken by lower candidate index: candidates are visited in increasing index order and a
// candidate only displaces entries with a strictly smaller key.
#include "common.cuh"
#include "gemm.cuh"  // fc_count_launch

namespace {

constexpr int QT = 32;     // queries per CTA
constexpr int CT = 128;    // candidates per tile
constexpr int CK = 32;     // feature chunk
constexpr int KNN_THREADS = 256;

struct TopK {
    float key[2];
    int idx[2];
};

// insert (v, id) into the warp-distributed descending list; entry e lives in lane e%32, slot e/32
__device__ __forceinline__ void topk_insert(TopK& t, float v, int id, int lane) {
    const unsigned m0 = __ballot_sync(0xffffffffu, t.key[0] >= v);
    const unsigned m1 = __ballot_sync(0xffffffffu, t.key[1] >= v);
    const int p = __popc(m0) + __popc(m1);  // insertion position (entries are sorted, so these are prefixes)
    // shifted copies: entry e takes the value of entry e-1
    float k0 = __shfl_up_sync(0xffffffffu, t.key[0], 1);
    int i0 = __shfl_up_sync(0xffffffffu, t.idx[0], 1);
    float k1 = __shfl_up_sync(0xffffffffu, t.key[1], 1);
    int i1 = __shfl_up_sync(0xffffffffu, t.idx[1], 1);
    const float wrapk = __shfl_sync(0xffffffffu, t.key[0], 31);
    const int wrapi = __shfl_sync(0xffffffffu, t.idx[0], 31);
    if (lane == 0) { k1 = wrapk; i1 = wrapi; }
    const int e0 = lane, e1 = lane + 32;
    if (e0 == p) { t.key[0] = v; t.idx[0] = id; }
    else if (e0 > p) { t.key[0] = k0; t.idx[0] = i0; }
    if (e1 == p) { t.key[1] = v; t.idx[1] = id; }
    else if (e1 > p) { t.key[1] = k1; t.idx[1] = i1; }
}

__device__ __forceinline__ float topk_threshold(const TopK& t, int k) {
    const int e = k - 1;
    const float v = (e < 32) ? t.key[0] : t.key[1];
    return __shfl_sync(0xffffffffu, v, e & 31);
}

// q: [B][Nq][ldq], t: [B][Nt][ldt]; C features.  mode 0: self form; mode 1: query form.
__global__ void __launch_bounds__(KNN_THREADS) knn_kernel(const float* __restrict__ q, int ldq, long long q_bstride,
                                                           const float* __restrict__ t, int ldt, long long t_bstride,
                                                           int Nq, int Nt, int C, int k, int mode,
                                                           int32_t* __restrict__ idx32, int64_t* __restrict__ idx64) {
    __shared__ __align__(16) float Qs[CK][QT + 4];   // rows 16-byte aligned: a warp's 4 query values are one broadcast LDS.128
    __shared__ float Cs[CK][CT + 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int q0 = blockIdx.x * QT;
    const float* qb = q + (size_t)b * q_bstride;
    const float* tb = t + (size_t)b * t_bstride;

    TopK top[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        top[i].key[0] = -INFINITY; top[i].key[1] = -INFINITY;
        top[i].idx[0] = -1; top[i].idx[1] = -1;
    }

    for (int j0 = 0; j0 < Nt; j0 += CT) {
        float dot[4][4];
        float xxq[4], xxc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            xxq[i] = 0.f; xxc[i] = 0.f;
#pragma unroll
            for (int s = 0; s < 4; ++s) dot[i][s] = 0.f;
        }
        for (int c0 = 0; c0 < C; c0 += CK) {
            __syncthreads();
            // stage the query chunk [CK][QT] and candidate chunk [CK][CT] (zero fill out of range)
            for (int e = tid; e < QT * CK; e += KNN_THREADS) {
                const int p = e / CK, c = e % CK;
                float v = 0.f;
                if (q0 + p < Nq && c0 + c < C) v = qb[(size_t)(q0 + p) * ldq + c0 + c];
                Qs[c][p] = v;
            }
            for (int e = tid; e < CT * CK; e += KNN_THREADS) {
                const int p = e / CK, c = e % CK;
                float v = 0.f;
                if (j0 + p < Nt && c0 + c < C) v = tb[(size_t)(j0 + p) * ldt + c0 + c];
                Cs[c][p] = v;
            }
            __syncthreads();
            const int cmax = min(CK, C - c0);
            for (int c = 0; c < cmax; ++c) {
                float cv[4];
                const float4 q4 = *reinterpret_cast<const float4*>(&Qs[c][warp * 4]);
                const float qv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                for (int s = 0; s < 4; ++s) cv[s] = Cs[c][lane + 32 * s];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    xxq[i] = fmaf(qv[i], qv[i], xxq[i]);
#pragma unroll
                    for (int s = 0; s < 4; ++s) dot[i][s] = fmaf(qv[i], cv[s], dot[i][s]);
                }
#pragma unroll
                for (int s = 0; s < 4; ++s) xxc[s] = fmaf(cv[s], cv[s], xxc[s]);
            }
        }
        // selection: warp owns queries warp*4 .. +3; lane holds candidates j0 + lane + 32*s
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float thr = topk_threshold(top[i], k);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int j = j0 + lane + 32 * s;
                float key;
                if (mode == 0) {
                    const float inner = -2.0f * dot[i][s];
                    key = __fsub_rn(__fsub_rn(-xxc[s], inner), xxq[i]);
                } else {
                    key = -__fsub_rn(__fadd_rn(xxq[i], xxc[s]), 2.0f * dot[i][s]);
                }
                if (j >= Nt) key = -INFINITY;
                unsigned m = __ballot_sync(0xffffffffu, key > thr);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float v = __shfl_sync(0xffffffffu, key, src);
                    if (v > thr) {  // thr may have risen since the ballot
                        topk_insert(top[i], v, j0 + src + 32 * s, lane);
                        thr = topk_threshold(top[i], k);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int qi = q0 + warp * 4 + i;
        if (qi >= Nq) continue;
        const size_t base = ((size_t)b * Nq + qi) * k;
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            const int e = lane + 32 * sl;
            if (e < k) {
                if (idx32) idx32[base + e] = top[i].idx[sl];
                if (idx64) idx64[base + e] = (int64_t)top[i].idx[sl];
            }
        }
    }
}

}  // namespace

int fc_knn_launch(const float* q, int ldq, long long q_bstride, const float* t, int ldt, long long t_bstride,
                  int B, int Nq, int Nt, int C, int k, int mode, int32_t* idx32, int64_t* idx64,
                  cudaStream_t stream) {
    FC_REQUIRE(q && t && B > 0 && Nq > 0 && Nt > 0 && C > 0);
    FC_REQUIRE(k >= 1 && k <= 64 && k <= Nt);
    FC_REQUIRE(B <= 65535);
    FC_REQUIRE(idx32 || idx64);
    dim3 grid((Nq + QT - 1) / QT, B);
    FcProfScope prof(FC_CLS_KNN, 2.0 * B * (double)Nq * Nt * C,
                     4.0 * B * ((double)(q == t ? Nt : Nq + Nt) * C + (double)Nq * k), stream);
    knn_kernel<<<grid, KNN_THREADS, 0, stream>>>(q, ldq, q_bstride, t, ldt, t_bstride, Nq, Nt, C, k, mode, idx32, idx64);
    fc_count_launch();
    FC_LAUNCH_OK();
    return FC_OK;
}

extern "C" int fc_knn_self(const float* x, int ldx, int B, int N, int C, int k, int32_t* idx32, int64_t* idx64,
                           fc_stream_t stream) {
    FC_REQUIRE(ldx >= C);
    return fc_knn_launch(x, ldx, (long long)N * ldx, x, ldx, (long long)N * ldx, B, N, N, C, k, 0, idx32, idx64,
                         (cudaStream_t)stream);
}

extern "C" int fc_knn_query(const float* q, const float* t, int Nq, int Nt, int D, int k, int64_t* idx64,
                            fc_stream_t stream) {
    return fc_knn_launch(q, D, 0, t, D, 0, 1, Nq, Nt, D, k, 1, nullptr, idx64, (cudaStream_t)stream);
}
