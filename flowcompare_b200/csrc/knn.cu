// Brute-force k-nearest-neighbour kernels: register-tiled distances, one-thread-per-query selection.
//
// Replaces reference models/pytorch_gcn.py:13-20 (`knn`: bmm + materialised [B,N,N] matrix + topk)
// and knn.py:40-52 (`KNN_torch_fun`).  Nothing of size N^2 is ever written.
//
// Arithmetic is the reference's algebraic form with a FIXED evaluation order (sequential fmaf over
// the feature index), which oracle/knn_ref.c restates bit-for-bit:
//   mode 0 (DGCNN):  key_ij = ((-xx_j) - (-2*dot_ij)) - xx_i          (k largest)
//   mode 1 (knn.py): key_ij = -((qq_i + tt_j) - 2*dot_ij)             (k largest == k smallest diss)
// Ties are broken by lower candidate index.
//
// Shape of the kernel (round 2; the round-1 kernel -- 4x4 register tile, norms recomputed per tile, one warp-wide
// ballot/shuffle insertion per candidate -- ran at 8 TFLOP/s, shared-memory-issue bound, ncu: profiles/r01_small_kernels_*):
//   * the squared norms are computed ONCE per point by norms_kernel (same sequential fmaf order as before);
//   * a CTA owns 64 queries and walks the candidates in tiles of 128; the 64 x 128 tile of dot products is accumulated
//     in an 8 x 8 REGISTER tile per thread (128 threads): per feature step two broadcast and two contiguous LDS.128
//     feed 64 FMAs (was 5 LDS per 16 FMAs); the next feature chunk is prefetched from global memory into registers
//     while the current one is multiplied;
//   * the keys of the tile go to shared memory and the selection is done by ONE THREAD PER QUERY (64 of the 128): it
//     filters its row against the query's current k-th best (almost everything fails after the first tiles), compacts the
//     survivors in place and insertion-sorts them into the query's sorted list in shared memory.  A warp-wide insertion
//     spends 32 lanes on one candidate; this spends one, and 64 queries insert concurrently.
// Two CTAs fit per SM (88 KB of shared memory each), so one CTA's selection phase overlaps the other's FMA phase.
#include "common.cuh"
#include "gemm.cuh"  // fc_count_launch
#include <atomic>
#include <mutex>

namespace {

constexpr int TM = 64;      // queries per CTA
constexpr int TN = 128;     // candidates per tile
constexpr int CK = 16;      // feature chunk
constexpr int KNN_THREADS = 128;
constexpr int KS_LD = TN + 1;   // key tile row stride: the selecting thread r reads Ks[r][j] -> bank (r + j) % 32, conflict free
constexpr int KNN_KMAX = 64;

// xx[b][i] = sum_c x[b][i][c]^2 with the canonical order (sequential fmaf over c)
__global__ void norms_kernel(const float* __restrict__ x, int ld, long long bstride, int N, int C, float* __restrict__ xx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= N) return;
    const float* p = x + (size_t)b * bstride + (size_t)i * ld;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = fmaf(p[c], p[c], s);
    xx[(size_t)b * N + i] = s;
}

struct KnnSmem {
    float Qs[CK][TM];                 // query chunk, feature-major
    float Cs[CK][TN];                 // candidate chunk, feature-major
    float Ks[TM][KS_LD];              // keys of the current tile
    unsigned char Js[TM][TN];         // compacted survivor columns of the current tile
    float qn[TM];                     // query norms
    float cn[TN];                     // candidate norms of the current tile
    float thr_x[2][TM];               // each half list's current k-th best, published for the other half of the same query
    // followed by the final lists of the two column halves: Lk[TM][2][KCAP] keys, Li[TM][2][KCAP] indices
};

// q: [B][Nq][ldq], t: [B][Nt][ldt]; C features.  mode 0: self form; mode 1: query form.
//
// Selection: TWO threads per query (thread r and r + 64 take columns [0,64) and [64,128) of every tile), each with its own
// sorted top-KCAP list IN REGISTERS.  Per tile a thread filters its 64 keys against its current k-th best and compacts the
// survivors in place (shared memory), then the warp inserts in lock step: insertion into the register list is branch free
// (one compare-select pair per slot, no dependent chain, no dynamic indexing), so 32 queries insert at once at the cost of
// ~5 instructions per slot.  (A sorted list in shared memory walked by one thread per query was tried first: bit-exact but
// 2.7x SLOWER than the round-1 kernel -- every shift is a dependent shared-memory round trip and only 8 warps per SM were
// there to hide it.)  The two half lists are merged once at the end under the total order (key desc, index asc).
template <int KCAP>
__global__ void __launch_bounds__(KNN_THREADS, 2)
knn_kernel(const float* __restrict__ q, int ldq, long long q_bstride, const float* __restrict__ t, int ldt,
           long long t_bstride, const float* __restrict__ qnorm, const float* __restrict__ tnorm, int Nq, int Nt, int C, int k,
           int mode, int32_t* __restrict__ idx32, int64_t* __restrict__ idx64) {
    extern __shared__ __align__(16) unsigned char knn_raw[];
    KnnSmem& sm = *reinterpret_cast<KnnSmem*>(knn_raw);
    float* Lk = reinterpret_cast<float*>(knn_raw + sizeof(KnnSmem));
    int* Li = reinterpret_cast<int*>(Lk + TM * 2 * KCAP);

    const int tid = threadIdx.x, lane = tid & 31;
    const int tx = tid & 15, ty = tid >> 4;        // 16 column groups x 8 row groups; thread tile: rows ty*8.., cols tx*4.. and 64+tx*4..
    const int srow = tid & (TM - 1), shalf = tid >> 6;   // selection role: query row and column half
    const int b = blockIdx.y;
    const int q0 = blockIdx.x * TM;
    const float* qb = q + (size_t)b * q_bstride;
    const float* tb = t + (size_t)b * t_bstride;

    // sorted register lists: keys -inf, index 0 (a query whose keys are all NaN keeps index 0: valid addressing downstream)
    float lkey[KCAP];
    int lidx[KCAP];
#pragma unroll
    for (int p = 0; p < KCAP; ++p) { lkey[p] = -INFINITY; lidx[p] = 0; }
    if (tid < TM) sm.qn[tid] = (q0 + tid < Nq) ? qnorm[(size_t)b * Nq + q0 + tid] : 0.f;
    sm.thr_x[shalf][srow] = -INFINITY;
    float thr = -INFINITY;     // this thread's current k-th best key

    const int n_chunks = (C + CK - 1) / CK;
    // staging map: a chunk is TM x CK query values and TN x CK candidate values; thread e handles element (p, c) = (e / CK, e % CK)
    constexpr int Q_PER = TM * CK / KNN_THREADS;    // 8
    constexpr int C_PER = TN * CK / KNN_THREADS;    // 16
    float rq[Q_PER], rc[C_PER];
    auto load_chunk = [&](int j0, int c0) {
#pragma unroll
        for (int i = 0; i < Q_PER; ++i) {
            const int e = tid + i * KNN_THREADS, p = e / CK, c = e % CK;
            rq[i] = (q0 + p < Nq && c0 + c < C) ? qb[(size_t)(q0 + p) * ldq + c0 + c] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < C_PER; ++i) {
            const int e = tid + i * KNN_THREADS, p = e / CK, c = e % CK;
            rc[i] = (j0 + p < Nt && c0 + c < C) ? tb[(size_t)(j0 + p) * ldt + c0 + c] : 0.f;
        }
    };
    auto store_chunk = [&]() {
#pragma unroll
        for (int i = 0; i < Q_PER; ++i) { const int e = tid + i * KNN_THREADS; sm.Qs[e % CK][e / CK] = rq[i]; }
#pragma unroll
        for (int i = 0; i < C_PER; ++i) { const int e = tid + i * KNN_THREADS; sm.Cs[e % CK][e / CK] = rc[i]; }
    };

    for (int j0 = 0; j0 < Nt; j0 += TN) {
        {
            float dot[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dot[i][j] = 0.f;
            load_chunk(j0, 0);
            for (int ch = 0; ch < n_chunks; ++ch) {
                __syncthreads();                       // everyone is done with the previous chunk (and the previous tile's selection)
                store_chunk();
                if (ch == 0 && tid < TN) sm.cn[tid] = (j0 + tid < Nt) ? tnorm[(size_t)b * Nt + j0 + tid] : 0.f;
                __syncthreads();
                if (ch + 1 < n_chunks) load_chunk(j0, (ch + 1) * CK);      // prefetch while multiplying
                const int cmax = min(CK, C - ch * CK);
#pragma unroll 4
                for (int c = 0; c < cmax; ++c) {
                    const float4 qa = *reinterpret_cast<const float4*>(&sm.Qs[c][ty * 8]);
                    const float4 qb4 = *reinterpret_cast<const float4*>(&sm.Qs[c][ty * 8 + 4]);
                    const float4 ca = *reinterpret_cast<const float4*>(&sm.Cs[c][tx * 4]);
                    const float4 cb = *reinterpret_cast<const float4*>(&sm.Cs[c][64 + tx * 4]);
                    const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb4.x, qb4.y, qb4.z, qb4.w};
                    const float cv[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) dot[i][j] = fmaf(qv[i], cv[j], dot[i][j]);
                }
            }
            // keys -> shared memory (Ks is only read by the selection below, after the barrier)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = ty * 8 + i;
                const float xq = sm.qn[r];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int col = (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                    const float xc = sm.cn[col];
                    float key;
                    if (mode == 0) key = __fsub_rn(__fsub_rn(-xc, -2.0f * dot[i][j]), xq);
                    else           key = -__fsub_rn(__fadd_rn(xq, xc), 2.0f * dot[i][j]);
                    if (j0 + col >= Nt) key = -INFINITY;
                    sm.Ks[r][col] = key;
                }
            }
        }
        __syncthreads();
        {
            // filter + compact in place (cnt <= j, so the write never overtakes the read); candidates stay in index order
            float* row = &sm.Ks[srow][shalf * 64];
            unsigned char* jr = &sm.Js[srow][shalf * 64];
            // The query's true k-th best is at least the k-th best of EITHER half list, so a key below the other half's
            // threshold cannot be in the result; keys EQUAL to it stay (the final merge decides ties by index).  This removes
            // ~40 % of the insertions (each half alone sees k(1 + ln(N/2k)) record breakers, together k(1 + ln(N/k))).
            const float oth = sm.thr_x[shalf ^ 1][srow];    // published before this tile's barriers
            int cnt = 0;
#pragma unroll 8
            for (int j = 0; j < 64; ++j) {
                const float v = row[j];
                if (v > thr && v >= oth) { row[cnt] = v; jr[cnt] = (unsigned char)j; ++cnt; }
            }
            const int mx = __reduce_max_sync(0xffffffffu, cnt);
            for (int i = 0; i < mx; ++i) {
                float v = -INFINITY;
                int id = 0;
                if (i < cnt) { v = row[i]; id = j0 + shalf * 64 + jr[i]; }
                if (!(v > thr)) v = -INFINITY;                     // the k-th best may have risen since the filter
                if (__any_sync(0xffffffffu, v > thr)) {
                    // branch-free insertion: slot p takes its upper neighbour if that one is smaller than v (shift down), else v if
                    // the slot itself is smaller (insertion point), else stays.  Strict <: an equal key with a lower index stays ahead.
                    // v = -inf changes nothing.
#pragma unroll
                    for (int p = KCAP - 1; p >= 1; --p) {
                        const bool up = lkey[p - 1] < v, here = lkey[p] < v;
                        lidx[p] = up ? lidx[p - 1] : (here ? id : lidx[p]);
                        lkey[p] = up ? lkey[p - 1] : (here ? v : lkey[p]);
                    }
                    if (lkey[0] < v) { lkey[0] = v; lidx[0] = id; }
                    if (k == KCAP) thr = lkey[KCAP - 1];       // the usual case (DGCNN: k = 40): no dependent select chain
                    else {
#pragma unroll
                        for (int p = 0; p < KCAP; ++p) if (p == k - 1) thr = lkey[p];
                    }
                }
            }
        }
        sm.thr_x[shalf][srow] = thr;
        // the next tile's first barrier orders this selection before Ks / cn are overwritten
    }
    // merge the two half lists of every query under (key desc, index asc)
#pragma unroll
    for (int p = 0; p < KCAP; ++p) { Lk[(srow * 2 + shalf) * KCAP + p] = lkey[p]; Li[(srow * 2 + shalf) * KCAP + p] = lidx[p]; }
    __syncthreads();
    if (tid < TM && q0 + tid < Nq) {
        const float* ka = Lk + (tid * 2) * KCAP; const float* kb = ka + KCAP;
        const int* ia = Li + (tid * 2) * KCAP; const int* ib = ia + KCAP;
        const size_t o = ((size_t)b * Nq + q0 + tid) * k;
        int i = 0, j = 0;
        for (int p = 0; p < k; ++p) {
            const bool take_a = (i < k) && (j >= k || ka[i] > kb[j] || (ka[i] == kb[j] && ia[i] <= ib[j]));
            const int v = take_a ? ia[i] : ib[j];
            if (take_a) ++i; else ++j;
            if (idx32) idx32[o + p] = v;
            if (idx64) idx64[o + p] = (int64_t)v;
        }
    }
}

template <int KCAP, typename... Args>
cudaError_t launch_knn(int dev, dim3 grid, cudaStream_t stream, Args... args) {
    const size_t smem = sizeof(KnnSmem) + (size_t)TM * 2 * KCAP * 8;
    static std::atomic<uint64_t> configured{0};    // the dynamic shared-memory opt-in is per device (and per instantiation)
    const uint64_t bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        cudaError_t e = cudaFuncSetAttribute(knn_kernel<KCAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured.fetch_or(bit, std::memory_order_release);
    }
    knn_kernel<KCAP><<<grid, KNN_THREADS, smem, stream>>>(args...);
    return cudaGetLastError();
}

}  // namespace

int64_t fc_knn_scratch_floats(int B, int Nq, int Nt, bool self) {
    return (int64_t)B * (self ? Nt : Nq + Nt);
}

int fc_knn_launch(const float* q, int ldq, long long q_bstride, const float* t, int ldt, long long t_bstride,
                  int B, int Nq, int Nt, int C, int k, int mode, int32_t* idx32, int64_t* idx64, float* norms_scratch,
                  cudaStream_t stream) {
    FC_REQUIRE(q && t && B > 0 && Nq > 0 && Nt > 0 && C > 0 && norms_scratch);
    FC_REQUIRE(k >= 1 && k <= KNN_KMAX && k <= Nt);
    FC_REQUIRE(B <= 65535);
    FC_REQUIRE(idx32 || idx64);
    const bool self = (q == t && ldq == ldt && q_bstride == t_bstride && Nq == Nt);
    float* tn = norms_scratch;
    float* qn = self ? tn : norms_scratch + (size_t)B * Nt;
    FcProfScope prof(FC_CLS_KNN, 2.0 * B * (double)Nq * Nt * C,
                     4.0 * B * ((double)(self ? Nt : Nq + Nt) * C + (double)Nq * k), stream);
    norms_kernel<<<dim3((Nt + 255) / 256, B), 256, 0, stream>>>(t, ldt, t_bstride, Nt, C, tn);
    fc_count_launch();
    if (!self) { norms_kernel<<<dim3((Nq + 255) / 256, B), 256, 0, stream>>>(q, ldq, q_bstride, Nq, C, qn); fc_count_launch(); }
    int dev = 0;
    FC_CUDA_OK(cudaGetDevice(&dev));
    dim3 grid((Nq + TM - 1) / TM, B);
    cudaError_t le;
    if (k <= 16)      le = launch_knn<16>(dev, grid, stream, q, ldq, q_bstride, t, ldt, t_bstride, qn, tn, Nq, Nt, C, k, mode, idx32, idx64);
    else if (k <= 32) le = launch_knn<32>(dev, grid, stream, q, ldq, q_bstride, t, ldt, t_bstride, qn, tn, Nq, Nt, C, k, mode, idx32, idx64);
    else if (k <= 40) le = launch_knn<40>(dev, grid, stream, q, ldq, q_bstride, t, ldt, t_bstride, qn, tn, Nq, Nt, C, k, mode, idx32, idx64);
    else              le = launch_knn<64>(dev, grid, stream, q, ldq, q_bstride, t, ldt, t_bstride, qn, tn, Nq, Nt, C, k, mode, idx32, idx64);
    FC_CUDA_OK(le);
    fc_count_launch();
    FC_LAUNCH_OK();
    return FC_OK;
}

// The op-level entry points own no workspace argument in the C ABI of round 1 (include/flowcompare_b200.h): the norms (4 bytes
// per point) are kept in a small per-device scratch that grows on demand and is reused by later calls on the same device.
namespace {
struct NormScratch { float* p = nullptr; int64_t floats = 0; };
NormScratch g_norms[64];
std::mutex g_norms_mu;
float* norms_for(int64_t floats, cudaStream_t stream) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(g_norms_mu);
    NormScratch& s = g_norms[dev & 63];
    if (s.floats < floats) {
        if (s.p) { cudaStreamSynchronize(stream); cudaFree(s.p); s.p = nullptr; s.floats = 0; }
        if (cudaMalloc(&s.p, (size_t)floats * 4) != cudaSuccess) return nullptr;
        s.floats = floats;
    }
    return s.p;
}
}  // namespace

extern "C" int fc_knn_self(const float* x, int ldx, int B, int N, int C, int k, int32_t* idx32, int64_t* idx64,
                           fc_stream_t stream) {
    FC_REQUIRE(x && ldx >= C && B > 0 && N > 0 && C > 0 && k >= 1 && k <= KNN_KMAX && k <= N && (idx32 || idx64));   // before any GPU work
    float* scratch = norms_for(fc_knn_scratch_floats(B, N, N, true), (cudaStream_t)stream);
    if (!scratch) return FC_ERR_CUDA;
    return fc_knn_launch(x, ldx, (long long)N * ldx, x, ldx, (long long)N * ldx, B, N, N, C, k, 0, idx32, idx64, scratch,
                         (cudaStream_t)stream);
}

extern "C" int fc_knn_query(const float* q, const float* t, int Nq, int Nt, int D, int k, int64_t* idx64,
                            fc_stream_t stream) {
    FC_REQUIRE(q && t && idx64 && Nq > 0 && Nt > 0 && D > 0 && k >= 1 && k <= KNN_KMAX && k <= Nt);   // before any GPU work
    float* scratch = norms_for(fc_knn_scratch_floats(1, Nq, Nt, false), (cudaStream_t)stream);
    if (!scratch) return FC_ERR_CUDA;
    return fc_knn_launch(q, D, 0, t, D, 0, 1, Nq, Nt, D, k, 1, nullptr, idx64, scratch, (cudaStream_t)stream);
}
