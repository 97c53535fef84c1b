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
#include <cstdlib>
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

// The same sums for wide rows (C >= 32): a warp owns 32 points and moves their rows through shared memory 32 columns at a time, so
// the global loads are coalesced (one 128-byte row segment per instruction) instead of 32 rows per instruction; each lane still adds
// ITS point's squares in column order -- the result is bit-identical to norms_kernel.
__global__ void __launch_bounds__(256) norms_wide_kernel(const float* __restrict__ x, int ld, long long bstride, int N, int C,
                                                         float* __restrict__ xx) {
    __shared__ float tile[8][32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, b = blockIdx.y;
    const int p0 = (blockIdx.x * 8 + w) * 32;
    if (p0 >= N) return;
    const float* xb = x + (size_t)b * bstride;
    float s = 0.f;
    for (int c0 = 0; c0 < C; c0 += 32) {
#pragma unroll 8
        for (int r = 0; r < 32; ++r)
            tile[w][r][lane] = (p0 + r < N && c0 + lane < C) ? xb[(size_t)(p0 + r) * ld + c0 + lane] : 0.f;
        __syncwarp();
        const int cmax = min(32, C - c0);
        for (int c = 0; c < cmax; ++c) { const float v = tile[w][lane][c]; s = fmaf(v, v, s); }
        __syncwarp();
    }
    if (p0 + lane < N) xx[(size_t)b * N + p0 + lane] = s;
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

// ----------------------------------------------------------------------------- two-kernel form (the default)
// The fused kernel above keeps two sorted top-k lists per query IN REGISTERS next to the 8x8 distance tile: 251 registers, two
// CTAs of four warps per SM, and a data-dependent insertion loop in lock step over 32 queries -- 9 TFLOP/s, 12 % of the FP32
// bound, and nothing short of separating the two jobs changes that.  So:
//   knn_dist_kernel    a plain FFMA tile kernel (128 x 128 keys per CTA, 8 x 8 per thread, double-buffered shared memory, 2 CTAs
//                      per SM) that writes the KEYS of a chunk of clouds (up to 1 GB) to scratch -- the same canonical arithmetic
//                      (sequential fmaf over the features, then ((-xx_j) - (-2 dot)) - xx_i), so every key is bit-identical to the
//                      fused form.  Measured 28 TFLOP/s at C = 64 .. 128 (the k-loop is only 4 .. 8 tiles long).
//   knn_select_kernel  one WARP per query row, no data-dependent loop: every lane loads its 40 strided keys of a 1280-key segment
//                      into registers, the k-th largest of the 96 lane-local top-3 values is a LOWER BOUND L of the row's k-th
//                      best (a subset's k-th largest never exceeds the set's), every lane writes its keys >= L (~45-55 in all)
//                      to shared memory as 64-bit (key, ~index) words at its prefix-sum offset, and the warp sorts them with a
//                      bitonic network IN REGISTERS (2 words per lane, shuffles); the first k are the answer in (key desc,
//                      index asc) order.  Longer rows repeat this per segment, carrying the sorted list.  More than 256 survivors
//                      (masses of exact duplicates) take an exact 32-step bisection with index-ordered ties instead.
//                      ~1400 warp instructions per row, 64 registers, 0.45 ms per layer at 128 clouds of 1250 points.
constexpr int DT = 128, DK = 16, DLD = DT + 4;

// VEC: 16-byte loads (C % 4 == 0, rows 16-byte aligned); otherwise scalar loads (the 6-column input layer)
template <bool VEC>
__global__ void __launch_bounds__(256, 2)
knn_dist_kernel(const float* __restrict__ q, int ldq, long long q_bstride, const float* __restrict__ t, int ldt,
                long long t_bstride, const float* __restrict__ qnorm, const float* __restrict__ tnorm, int Nq, int Nt, int C,
                int mode, float* __restrict__ keys, int ldk, int b0) {
    __shared__ __align__(16) float Qs[2][DK][DLD];
    __shared__ __align__(16) float Ts[2][DK][DLD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int bl = blockIdx.z, b = b0 + bl;
    const int q0 = blockIdx.y * DT, t0 = blockIdx.x * DT;
    const int a_row = tid >> 2, a_kq = (tid & 3) * 4;
    // this thread's two query rows and two candidate rows (64 apart); rows past the end read row 0 and are zeroed
    const bool qok0 = q0 + a_row < Nq, qok1 = q0 + a_row + 64 < Nq, tok0 = t0 + a_row < Nt, tok1 = t0 + a_row + 64 < Nt;
    const float* qp0 = q + (size_t)b * q_bstride + (size_t)(qok0 ? q0 + a_row : 0) * ldq + a_kq;
    const float* tp0 = t + (size_t)b * t_bstride + (size_t)(tok0 ? t0 + a_row : 0) * ldt + a_kq;
    const size_t qs1 = (size_t)64 * ldq, ts1 = (size_t)64 * ldt;      // never dereferenced for rows past the end (ld4 checks)
    float4 rq0, rq1, rt0, rt1;
    auto ld4 = [&](const float* src, int k, bool ok) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (VEC) { if (ok && k < C) v = *reinterpret_cast<const float4*>(src); }
        else if (ok) { if (k < C) v.x = src[0]; if (k + 1 < C) v.y = src[1]; if (k + 2 < C) v.z = src[2]; if (k + 3 < C) v.w = src[3]; }
        return v;
    };
    auto load = [&](int c0) {
        const int k = c0 + a_kq;
        rq0 = ld4(qp0 + c0, k, qok0); rq1 = ld4(qp0 + qs1 + c0, k, qok1);
        rt0 = ld4(tp0 + c0, k, tok0); rt1 = ld4(tp0 + ts1 + c0, k, tok1);
    };
    auto store = [&](int buf) {
        Qs[buf][a_kq + 0][a_row] = rq0.x; Qs[buf][a_kq + 1][a_row] = rq0.y; Qs[buf][a_kq + 2][a_row] = rq0.z; Qs[buf][a_kq + 3][a_row] = rq0.w;
        Qs[buf][a_kq + 0][a_row + 64] = rq1.x; Qs[buf][a_kq + 1][a_row + 64] = rq1.y; Qs[buf][a_kq + 2][a_row + 64] = rq1.z; Qs[buf][a_kq + 3][a_row + 64] = rq1.w;
        Ts[buf][a_kq + 0][a_row] = rt0.x; Ts[buf][a_kq + 1][a_row] = rt0.y; Ts[buf][a_kq + 2][a_row] = rt0.z; Ts[buf][a_kq + 3][a_row] = rt0.w;
        Ts[buf][a_kq + 0][a_row + 64] = rt1.x; Ts[buf][a_kq + 1][a_row + 64] = rt1.y; Ts[buf][a_kq + 2][a_row + 64] = rt1.z; Ts[buf][a_kq + 3][a_row + 64] = rt1.w;
    };
    float dot[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dot[i][j] = 0.f;
    const int T = (C + DK - 1) / DK;
    load(0);
    store(0);
    __syncthreads();
    for (int tt = 0; tt < T; ++tt) {
        const int buf = tt & 1;
        if (tt + 1 < T) load((tt + 1) * DK);
        const int cmax = min(DK, C - tt * DK);      // canonical order: features 0 .. C-1, nothing beyond
#pragma unroll 4
        for (int kk = 0; kk < cmax; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&Qs[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&Qs[buf][kk][64 + ty * 4]);
            const float4 c0 = *reinterpret_cast<const float4*>(&Ts[buf][kk][tx * 4]);
            const float4 c1 = *reinterpret_cast<const float4*>(&Ts[buf][kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dot[i][j] = fmaf(av[i], cv[j], dot[i][j]);
        }
        if (tt + 1 < T) { store(buf ^ 1); __syncthreads(); }
    }
    // keys: one 4-column group at a time (the candidate norms are re-read per group: keeps the epilogue's registers low)
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const int col = t0 + g * 64 + tx * 4;
        float xc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) xc[j] = col + j < Nt ? tnorm[(size_t)b * Nt + col + j] : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = q0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
            if (row >= Nq) continue;
            const float xq = qnorm[(size_t)b * Nq + row];
            float kv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float d = dot[i][g * 4 + j];
                kv[j] = mode == 0 ? __fsub_rn(__fsub_rn(-xc[j], -2.0f * d), xq) : -__fsub_rn(__fadd_rn(xq, xc[j]), 2.0f * d);
            }
            float* dst = keys + ((size_t)bl * Nq + row) * ldk + col;
            if (col + 3 < Nt) *reinterpret_cast<float4*>(dst) = make_float4(kv[0], kv[1], kv[2], kv[3]);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (col + j < Nt) dst[j] = kv[j];
            }
        }
    }
}

#ifndef SEL_KO
#define SEL_KO 0                       // timing experiments: 1 no bisection, 2 no compaction, 4 no sort (bit mask; results are garbage)
#endif
constexpr int SEL_R = 40;              // keys per lane and segment
constexpr int SEL_SEG = 32 * SEL_R;    // 1280
constexpr int SEL_CAP = 256;           // survivor buffer (64-bit words) per warp
__device__ __forceinline__ unsigned knn_ord(float x) {      // monotone float -> unsigned (larger float <-> larger word); 0 is below every float
    const unsigned u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
    const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, m), hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m);
    return ((unsigned long long)hi << 32) | lo;
}
// Bitonic sort (descending) of 32*NPL words held NPL per lane, element index = lane + 32*j: partners closer than 32 are
// another lane's register j (shuffles), partners 32*m apart are this lane's own registers.
template <int NPL>
__device__ __forceinline__ void warp_bitonic_desc(unsigned long long (&v)[NPL], int lane) {
#pragma unroll
    for (int sz = 2; sz <= 32 * NPL; sz <<= 1) {
#pragma unroll
        for (int st = sz >> 1; st > 0; st >>= 1) {
            if (st >= 32) {
                const int js = st >> 5;
#pragma unroll
                for (int j = 0; j < NPL; ++j) {
                    if ((j & js) == 0) {
                        const bool desc = (((lane + 32 * j) & sz) == 0);
                        const unsigned long long a = v[j], c = v[j | js];
                        if ((a < c) == desc) { v[j] = c; v[j | js] = a; }
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < NPL; ++j) {
                    const unsigned long long other = shfl_xor_u64(v[j], st);
                    const bool desc = (((lane + 32 * j) & sz) == 0), lower = (lane & st) == 0;
                    // the lower index of a pair keeps the larger word when the run is descending
                    const bool keep_max = lower == desc;
                    v[j] = keep_max ? (v[j] > other ? v[j] : other) : (v[j] < other ? v[j] : other);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
knn_select_kernel(const float* __restrict__ keys, int ldk, int n_rows, int Nt, int k, int32_t* __restrict__ idx32,
                  int64_t* __restrict__ idx64, long long out_row0) {
    __shared__ unsigned long long sbuf[8][SEL_CAP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int row = blockIdx.x * 8 + wid;
    if (row >= n_rows) return;
    unsigned long long* sv = sbuf[wid];
    const float* kr = keys + (size_t)row * ldk;
    const unsigned lt = (1u << lane) - 1u;
    int have = 0;                          // entries of the running list (sorted, sv[0 .. have))
    for (int seg0 = 0; seg0 < Nt; seg0 += SEL_SEG) {
        unsigned r[SEL_R];
        unsigned m1 = 0u, m2 = 0u, m3 = 0u;   // this lane's three largest keys of the segment
#pragma unroll
        for (int i = 0; i < SEL_R; ++i) {
            const int e = seg0 + i * 32 + lane;
            const float x = kr[min(e, Nt - 1)];
            r[i] = e < Nt ? knn_ord(x) : 0u;
            const unsigned a = min(m1, r[i]);
            m1 = max(m1, r[i]);
            const unsigned c = min(m2, a);
            m2 = max(m2, a);
            m3 = max(m3, c);
        }
        // the running list's keys (two per lane; zero beyond `have`)
        const unsigned long long w0 = lane < have ? sv[lane] : 0ull, w1 = 32 + lane < have ? sv[32 + lane] : 0ull;
        const unsigned l0 = (unsigned)(w0 >> 32), l1 = (unsigned)(w1 >> 32);
        // L = k-th largest of the 96 lane-local top-3 keys (a lower bound of the row's k-th best): bisection on the ordered words
        unsigned L = 0u;
        if (SEL_KO & 1) L = __reduce_max_sync(0xffffffffu, m3);
        else
        for (int bit = 31; bit >= 0; --bit) {
            const unsigned cand = L | (1u << bit);
            const unsigned cnt = __reduce_add_sync(0xffffffffu, (unsigned)(m1 >= cand) + (unsigned)(m2 >= cand) + (unsigned)(m3 >= cand));
            if (cnt >= (unsigned)k) L = cand;
        }
        if (have >= k) { const unsigned lk = (unsigned)(sv[k - 1] >> 32); L = max(L, lk); }   // the list's k-th best is a lower bound too
        unsigned thr_gt = L ? L - 1u : 0u;  // keys > thr_gt always survive
        unsigned cnt_l = (unsigned)(l0 > thr_gt) + (unsigned)(l1 > thr_gt);
#pragma unroll
        for (int i = 0; i < SEL_R; ++i) cnt_l += (unsigned)(r[i] > thr_gt);
        unsigned tot = __reduce_add_sync(0xffffffffu, cnt_l);
        __syncwarp();
        int base;
        if (tot <= (unsigned)SEL_CAP) {
            // every lane writes its own survivors at its own offset (order in the buffer is irrelevant: the sort below orders
            // by (key, index)); exclusive prefix sum of the per-lane counts
            unsigned off = cnt_l;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned n = __shfl_up_sync(0xffffffffu, off, d); if (lane >= d) off += n; }
            off -= cnt_l;
            if (l0 > thr_gt) sv[off++] = w0;
            if (l1 > thr_gt) sv[off++] = w1;
#pragma unroll
            for (int i = 0; i < ((SEL_KO & 2) ? 2 : SEL_R); ++i) {
                if (r[i] > thr_gt) sv[off++] = ((unsigned long long)r[i] << 32) | (unsigned)(~(unsigned)(seg0 + i * 32 + lane));
            }
            base = (int)tot;
        } else {
            // masses of equal keys: exact k-th largest T of (list + segment) by bisection, then everything above T plus the first
            // equals in index order (the list entries precede the segment)
            unsigned Tk = 0u;
            for (int bit = 31; bit >= 0; --bit) {
                const unsigned cand = Tk | (1u << bit);
                unsigned cc = (unsigned)(l0 >= cand) + (unsigned)(l1 >= cand);
#pragma unroll
                for (int i = 0; i < SEL_R; ++i) cc += (unsigned)(r[i] >= cand);
                if (__reduce_add_sync(0xffffffffu, cc) >= (unsigned)k) Tk = cand;
            }
            unsigned cg = (unsigned)(l0 > Tk) + (unsigned)(l1 > Tk);
#pragma unroll
            for (int i = 0; i < SEL_R; ++i) cg += (unsigned)(r[i] > Tk);
            int need = k - (int)__reduce_add_sync(0xffffffffu, cg);
            base = 0;
            auto push = [&](bool gt, bool eq, unsigned long long word) {
                const unsigned meq = __ballot_sync(0xffffffffu, eq);
                const bool take = gt || (eq && __popc(meq & lt) < need);
                need = max(0, need - __popc(meq));
                const unsigned m = __ballot_sync(0xffffffffu, take);
                if (take) sv[base + __popc(m & lt)] = word;
                base += __popc(m);
            };
            push(l0 > Tk, l0 == Tk && lane < have, w0);
            push(l1 > Tk, l1 == Tk && 32 + lane < have, w1);
#pragma unroll 1
            for (int i = 0; i < SEL_R; ++i) {
                unsigned ri = 0u;
#pragma unroll
                for (int jj = 0; jj < SEL_R; ++jj) if (jj == i) ri = r[jj];
                const int e = seg0 + i * 32 + lane;
                push(ri > Tk, ri == Tk && e < Nt, ((unsigned long long)ri << 32) | (unsigned)(~(unsigned)e));
            }
        }
        __syncwarp();
        // sort the survivors (descending), padded with zeros: in registers, 2 / 4 / 8 words per lane
        if (!(SEL_KO & 4)) {
            if (base <= 64) {
                unsigned long long v[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) v[j] = lane + 32 * j < base ? sv[lane + 32 * j] : 0ull;
                warp_bitonic_desc<2>(v, lane);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 2; ++j) sv[lane + 32 * j] = v[j];
            } else if (base <= 128) {
                unsigned long long v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = lane + 32 * j < base ? sv[lane + 32 * j] : 0ull;
                warp_bitonic_desc<4>(v, lane);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j) sv[lane + 32 * j] = v[j];
            } else {
                unsigned long long v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = lane + 32 * j < base ? sv[lane + 32 * j] : 0ull;
                warp_bitonic_desc<8>(v, lane);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j) sv[lane + 32 * j] = v[j];
            }
        }
        __syncwarp();
        have = min(base, k);
    }
    const size_t o = (size_t)(out_row0 + row) * k;
    for (int p = lane; p < k; p += 32) {
        const int v = p < have ? (int)(~(unsigned)sv[p]) : 0;
        if (idx32) idx32[o + p] = v;
        if (idx64) idx64[o + p] = (int64_t)v;
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

namespace {
// clouds per chunk of the two-kernel form: up to 1 GB of keys at a time.  (L2-sized chunks of 100 MB were measured SLOWER, 5.58
// vs 5.02 ms for the four layers at 128 clouds of 1250 points: eight short launches with their tails cost more than the HBM
// round trip of 0.8 GB, which takes 0.25 ms.)
int knn_chunk_clouds(int B, int Nq, int Nt) {
    const long long per = (long long)Nq * fc_round_up(Nt, 4);
    static long long budget = 0;       // floats of keys per chunk (FC_KNN_CHUNK_MB: A/B knob)
    if (!budget) { const char* e = getenv("FC_KNN_CHUNK_MB"); budget = (e ? atoll(e) : 1024ll) << 18; if (budget < 1) budget = 256ll << 20; }
    long long cb = budget / (per > 0 ? per : 1);
    if (cb < 1) cb = 1;
    if (cb > B) cb = B;
    return (int)cb;
}
int64_t knn_norm_floats(int B, int Nq, int Nt, bool self) { return fc_round_up_ll((int64_t)B * (self ? Nt : Nq + Nt), 64); }
}  // namespace

// scratch of a kNN call: the squared norms of the points, then the keys of one chunk of clouds
int64_t fc_knn_scratch_floats(int B, int Nq, int Nt, bool self) {
    return knn_norm_floats(B, Nq, Nt, self) + (int64_t)knn_chunk_clouds(B, Nq, Nt) * Nq * fc_round_up(Nt, 4);
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
    auto norms = [&](const float* p, int ld, long long bs, int n, float* out) {
        if (C >= 32) norms_wide_kernel<<<dim3((n + 255) / 256, B), 256, 0, stream>>>(p, ld, bs, n, C, out);
        else         norms_kernel<<<dim3((n + 255) / 256, B), 256, 0, stream>>>(p, ld, bs, n, C, out);
        fc_count_launch();
    };
    norms(t, ldt, t_bstride, Nt, tn);
    if (!self) norms(q, ldq, q_bstride, Nq, qn);
    static int fused_env = -1;     // FC_KNN=fused: the one-kernel form (A/B runs)
    if (fused_env < 0) { const char* e = getenv("FC_KNN"); fused_env = (e && e[0] == 'f') ? 1 : 0; }
    if (!fused_env) {
        float* keys = norms_scratch + knn_norm_floats(B, Nq, Nt, self);
        const int ldk = fc_round_up(Nt, 4);
        const int cb = knn_chunk_clouds(B, Nq, Nt);
        static int skip = -1;          // FC_KNN_SKIP=d|s: timing experiments (results are garbage)
        if (skip < 0) { const char* e = getenv("FC_KNN_SKIP"); skip = !e ? 0 : (e[0] == 'd' ? 1 : 2); }
        for (int b0 = 0; b0 < B; b0 += cb) {
            const int nb = B - b0 < cb ? B - b0 : cb;
            if (skip != 1)
            {
                const dim3 grid((Nt + DT - 1) / DT, (Nq + DT - 1) / DT, nb);
                const bool vec = (C & 3) == 0 && (ldq & 3) == 0 && (ldt & 3) == 0 && (q_bstride & 3) == 0 && (t_bstride & 3) == 0 &&
                                 ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(t)) & 15) == 0;
                if (vec) knn_dist_kernel<true><<<grid, 256, 0, stream>>>(q, ldq, q_bstride, t, ldt, t_bstride, qn, tn, Nq, Nt, C, mode, keys, ldk, b0);
                else     knn_dist_kernel<false><<<grid, 256, 0, stream>>>(q, ldq, q_bstride, t, ldt, t_bstride, qn, tn, Nq, Nt, C, mode, keys, ldk, b0);
            }
            const long long rows = (long long)nb * Nq;
            if (skip != 2)
            knn_select_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(keys, ldk, (int)rows, Nt, k, idx32, idx64, (long long)b0 * Nq);
            fc_count_launch(2);
        }
        FC_LAUNCH_OK();
        return FC_OK;
    }
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

// Workspace forms of the two entry points above: nothing is allocated and no per-device state is touched, so calls on different
// streams (or host threads) of one device are independent.  The workspace holds the norms and the keys of one chunk of clouds.
extern "C" int64_t fc_knn_workspace_bytes(int B, int Nq, int Nt, int self) {
    if (B < 1 || Nq < 1 || Nt < 1) return 0;
    return fc_knn_scratch_floats(B, Nq, Nt, self != 0) * 4 + 256;
}

namespace {
float* knn_ws_base(void* workspace, int64_t workspace_bytes, int64_t need_floats) {
    if (!workspace) return nullptr;
    const uintptr_t p = reinterpret_cast<uintptr_t>(workspace);
    const uintptr_t a = (p + 255) & ~(uintptr_t)255;
    if ((int64_t)(a - p) + need_floats * 4 > workspace_bytes) return nullptr;
    return reinterpret_cast<float*>(a);
}
}  // namespace

extern "C" int fc_knn_self_ws(const float* x, int ldx, int B, int N, int C, int k, int32_t* idx32, int64_t* idx64,
                              void* workspace, int64_t workspace_bytes, fc_stream_t stream) {
    FC_REQUIRE(x && ldx >= C && B > 0 && N > 0 && C > 0 && k >= 1 && k <= KNN_KMAX && k <= N && (idx32 || idx64));   // before any GPU work
    float* scratch = knn_ws_base(workspace, workspace_bytes, fc_knn_scratch_floats(B, N, N, true));
    if (!scratch) return FC_ERR_WORKSPACE;   // missing, or smaller than fc_knn_workspace_bytes says
    return fc_knn_launch(x, ldx, (long long)N * ldx, x, ldx, (long long)N * ldx, B, N, N, C, k, 0, idx32, idx64, scratch,
                         (cudaStream_t)stream);
}

extern "C" int fc_knn_query_ws(const float* q, const float* t, int Nq, int Nt, int D, int k, int64_t* idx64,
                               void* workspace, int64_t workspace_bytes, fc_stream_t stream) {
    FC_REQUIRE(q && t && idx64 && Nq > 0 && Nt > 0 && D > 0 && k >= 1 && k <= KNN_KMAX && k <= Nt);   // before any GPU work
    float* scratch = knn_ws_base(workspace, workspace_bytes, fc_knn_scratch_floats(1, Nq, Nt, false));
    if (!scratch) return FC_ERR_WORKSPACE;   // missing, or smaller than fc_knn_workspace_bytes says
    return fc_knn_launch(q, D, 0, t, D, 0, 1, Nq, Nt, D, k, 1, nullptr, idx64, scratch, (cudaStream_t)stream);
}

extern "C" int fc_knn_query(const float* q, const float* t, int Nq, int Nt, int D, int k, int64_t* idx64,
                            fc_stream_t stream) {
    FC_REQUIRE(q && t && idx64 && Nq > 0 && Nt > 0 && D > 0 && k >= 1 && k <= KNN_KMAX && k <= Nt);   // before any GPU work
    float* scratch = norms_for(fc_knn_scratch_floats(1, Nq, Nt, false), (cudaStream_t)stream);
    if (!scratch) return FC_ERR_CUDA;
    return fc_knn_launch(q, D, 0, t, D, 0, 1, Nq, Nt, D, k, 1, nullptr, idx64, scratch, (cudaStream_t)stream);
}
