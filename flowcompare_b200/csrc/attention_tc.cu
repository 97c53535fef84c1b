// Single-head cross attention on the 5th-gen tensor cores (tcgen05 + TMEM), fp32-faithful (3xTF32).
// Replaces reference models/perceiver.py:108-115 (AttentionMine.forward: softmax(q k^T * scale) v) for
// precision = tf32x3; attention_mma.cu (mma.sync) and attention.cu (exact fp32 SIMT) stay as the smaller twins.
//
// CTA = 128 queries of one cloud; the context cloud's keys are streamed in blocks of 64 (flash style, nothing of
// size N x Nc is ever written).  Per key block i:
//     S_i = Q K_i^T          24 tcgen05.mma (128x64x8, TF32):  Qlo*Khi + Qhi*Klo + Qhi*Khi     A = Q  in TMEM
//     P_i = exp2(S_i - m_i)  one softmax thread per query row straight out of TMEM (no shuffles), split to hi/lo
//     O_i = P_i V_i          24 tcgen05.mma:  Plo*Vhi + Phi*Vlo + Phi*Vhi                      A = P  in TMEM
//     O   = O * alpha + O_i  in registers (fp32), so the tensor core never chains more than 24 accumulations
// K and V arrive pre-split into TF32 hi/lo (kv_split_kernel below; V transposed to [dim][key] so that both MMAs
// take K-major B operands through the same 128B-swizzle descriptors as the GEMM) by TMA into two 2-stage rings.
//   warp 0   TMA producer      warp 1   TMEM allocator + issuer of the S MMAs      warp 2   issuer of the PV MMAs
//   warps 3..   Q converter, softmax, O accumulation, output: A_SPL threads per query row
// TMEM (512 columns): Qhi [0,64) | Qlo [64,128) | S x2 [128,256) | Phi [256,320) | Plo [320,384) | O_i x2 [384,512).
// (F16: Qhi [0,32) | Qlo [32,64) | S x2 [128,256) | P buffer b: hi [256+64b, +32), lo [288+64b, +32) | O_i x2 [384,512))
//
// F16 = true (precision fp16x3, round 2): the same pipeline on 3xFP16 operands -- kind::f16 MMAs (K = 16 per instruction: 12
// per product instead of 24, each at the same 32 cycles), K / V tiles of half the bytes (one 128-byte swizzle atom per hi / lo
// part, four stages instead of two), Q / P as packed fp16 pairs in half the TMEM columns.  hi = fp16(x), lo = fp16(x - hi)
// UNSCALED (the three products share one accumulator, as in the TF32 form); to keep lo clear of fp16's subnormal range Q is
// carried times 16 and P times 1024 -- exact powers of two: Q's is folded into the exponent's constant, P's is an exponent shift
// that cancels in O / l.  The kernel was tensor bound at TF32 rate (62 % pipe active, profiles/r02_attention_tc_*).
#include <cuda_fp16.h>
#include "common.cuh"
#include "gemm.cuh"
#include "tcgen05.cuh"
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace {

constexpr int AQ = 128;                 // queries per CTA
constexpr int AK = 64;                  // keys per block
constexpr int AD = 64;                  // head dim
#ifndef A_SPL
#define A_SPL 4                        // softmax threads per query row (1, 2 or 4): 4 measured +3 % (1 k keys) .. +6.5 % (32 k) in the fp16 form
#endif
constexpr int A_THREADS = 96 + 128 * A_SPL;
constexpr int A_ATOM = 64 * 128;        // 64 rows x 128 B  (one swizzle atom of a 64-row tile)
constexpr int A_QBYTES = 2 * AQ * 128;  // raw fp32 Q tile: two atoms of 128 rows
constexpr uint32_t C_S = 128, C_O = 384;
template <bool F16> struct ACfg {
    static constexpr int TILE = F16 ? 2 * A_ATOM : 4 * A_ATOM;     // hi + lo of a K (or V^T) block
    static constexpr int STAGES = F16 ? 4 : 2;
    static constexpr int SMEM = A_QBYTES + 2 * STAGES * TILE + 1024;
    static constexpr uint32_t C_QHI = 0, C_QLO = F16 ? 32 : 64, C_PHI = 256, C_PLO = F16 ? 288 : 320;
};
constexpr float A_QS = 16.0f;   // F16: power-of-two scale of Q (P's 2^10 is an exponent shift, see the softmax)

struct AttnTcParams {
    float* out; int ldo;
    int N, Nc, nblk;
    float scale;
    int debug;
    int dbg_loads;   // experiment: 1 = skip the lo tiles' loads (timing only, wrong results)
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t tf32_hi(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }

// D[128 x 64] (+)= A[tmem, 64 columns of K] * B[64-row K-major tile: two 32-wide swizzle atoms], 3xTF32
__device__ __forceinline__ void issue_3xtf32(uint32_t d, uint32_t a_hi, uint32_t a_lo, const unsigned char* b_tile, uint32_t idesc) {
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        const uint32_t atom = (uint32_t)(kk >> 2) * A_ATOM;
        const uint64_t adv = (uint64_t)((kk & 3) * 32 >> 4);
        const uint64_t dbh = make_kmajor_sw128_desc(smem_u32(b_tile + atom)) + adv;
        const uint64_t dbl = make_kmajor_sw128_desc(smem_u32(b_tile + 2 * A_ATOM + atom)) + adv;
        umma_tf32_ts(d, a_lo + 8 * kk, dbh, idesc, kk != 0);
        umma_tf32_ts(d, a_hi + 8 * kk, dbl, idesc, 1);
        umma_tf32_ts(d, a_hi + 8 * kk, dbh, idesc, 1);
    }
}

// D[128 x 64] (+)= A[tmem, 32 columns = 64 fp16 of K] * B[64-row K-major tile, ONE 128-byte swizzle atom per part], 3xFP16
__device__ __forceinline__ void issue_3xf16(uint32_t d, uint32_t a_hi, uint32_t a_lo, const unsigned char* b_tile, uint32_t idesc) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const uint64_t adv = (uint64_t)(kk * 32 >> 4);
        const uint64_t dbh = make_kmajor_sw128_desc(smem_u32(b_tile)) + adv;
        const uint64_t dbl = make_kmajor_sw128_desc(smem_u32(b_tile + A_ATOM)) + adv;
        umma_f16_ts(d, a_lo + 8 * kk, dbh, idesc, kk != 0);
        umma_f16_ts(d, a_hi + 8 * kk, dbl, idesc, 1);
        umma_f16_ts(d, a_hi + 8 * kk, dbh, idesc, 1);
    }
}

// debug (FC_ATTN_DEBUG=1): cycles [0] CTAs [1] total [2..6] MMA warp waits kfull,s_free,vfull,o_free,p_ready
//   [7..9] softmax thread waits s_ready,p_free,o_ready [10] softmax total [11] Q phase
__device__ unsigned long long fc_attn_dbg[16];
#ifndef A_PHASE_TIMERS
#define A_PHASE_TIMERS 0
#endif
#if A_PHASE_TIMERS
#define DBG_T(var) const long long var = p.debug ? clock64() : 0
#else
#define DBG_T(var) const long long var = 0
#endif

template <bool F16>
__global__ void __launch_bounds__(A_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapKhi,
                    const __grid_constant__ CUtensorMap mapKlo, const __grid_constant__ CUtensorMap mapVhi,
                    const __grid_constant__ CUtensorMap mapVlo, const AttnTcParams p) {
    constexpr int A_STAGES = ACfg<F16>::STAGES, A_TILE = ACfg<F16>::TILE;
    constexpr uint32_t C_QHI = ACfg<F16>::C_QHI, C_QLO = ACfg<F16>::C_QLO, C_PHI = ACfg<F16>::C_PHI, C_PLO = ACfg<F16>::C_PLO;
    static_assert(!F16 || A_SPL == 2 || A_SPL == 4, "the fp16 form packs 32 (16) dims / keys per thread into one 16 (8) column store");
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 + 4 * A_STAGES + 12];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float mx_sm[2][A_SPL][AQ];     // row-maximum exchange between the threads of a row (by block parity)

    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    unsigned char* q_raw = smem;
    auto k_tile = [&](int s) { return smem + A_QBYTES + s * A_TILE; };
    auto v_tile = [&](int s) { return smem + A_QBYTES + (A_STAGES + s) * A_TILE; };
    uint64_t* qfull = bars;                   // Q tile landed
    uint64_t* qready = bars + 1;              // Q hi/lo in TMEM                      (128 arrivals)
    uint64_t* kfull = bars + 2;               // [A_STAGES]
    uint64_t* kfree = kfull + A_STAGES;       // [A_STAGES] S MMAs retired
    uint64_t* vfull = kfree + A_STAGES;       // [A_STAGES]
    uint64_t* vfree = vfull + A_STAGES;       // [A_STAGES] PV MMAs retired
    uint64_t* s_ready = vfree + A_STAGES;     // [2] S_i complete in TMEM
    uint64_t* s_free = s_ready + 2;           // [2] softmax threads have S_i in registers (128 arrivals)
    // P is double buffered in the fp16 form (its packed pairs take half the columns): the PV MMAs of block i run while the
    // softmax of block i+1 already writes its P -- with one buffer every block paid softmax + PV back to back
    constexpr int NPB = F16 ? 2 : 1;
    uint64_t* p_ready = s_free + 2;           // [2] P_i hi/lo in TMEM                (128 arrivals)
    uint64_t* p_free = p_ready + 2;           // [2] PV_i retired: P may be overwritten
    uint64_t* o_ready = p_free + 2;           // [2] O_i complete in TMEM
    uint64_t* o_free = o_ready + 2;           // [2] O_i read back                    (128 arrivals)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int q0 = blockIdx.x * AQ;
    const int nblk = p.nblk;

    if (threadIdx.x == 0) {
        mbar_init(qfull, 1); mbar_init(qready, 128 * A_SPL);
        for (int s = 0; s < A_STAGES; ++s) { mbar_init(&kfull[s], 1); mbar_init(&kfree[s], 1); mbar_init(&vfull[s], 1); mbar_init(&vfree[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&s_ready[s], 1); mbar_init(&s_free[s], 128 * A_SPL); mbar_init(&o_ready[s], 1); mbar_init(&o_free[s], 128 * A_SPL); }
        for (int s = 0; s < 2; ++s) { mbar_init(&p_ready[s], 128 * A_SPL); mbar_init(&p_free[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_slot;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            mbar_expect_tx(qfull, A_QBYTES);
            tma_load_2d(&mapQ, q_raw, qfull, 0, b * p.N + q0);
            tma_load_2d(&mapQ, q_raw + AQ * 128, qfull, 32, b * p.N + q0);
            for (int i = 0; i < nblk; ++i) {
                const int s = i % A_STAGES;
                const uint32_t ph = (i / A_STAGES) & 1;
                const int krow = b * p.Nc + i * AK;      // rows past this cloud are masked in the softmax
                mbar_wait(&kfree[s], ph ^ 1, 100 + i);
                if constexpr (F16) {
                    // one 64 x 128-byte box per part: 64 fp16 dims of 64 keys (K) / 64 fp16 keys of 64 dims (V^T)
                    mbar_expect_tx(&kfull[s], A_TILE);
                    tma_load_2d(&mapKhi, k_tile(s), &kfull[s], 0, krow);
                    tma_load_2d(&mapKlo, k_tile(s) + A_ATOM, &kfull[s], 0, krow);
                    mbar_wait(&vfree[s], ph ^ 1, 150 + i);
                    mbar_expect_tx(&vfull[s], A_TILE);
                    tma_load_3d(&mapVhi, v_tile(s), &vfull[s], i * AK, 0, b);        // keys past Nc: zero fill
                    tma_load_3d(&mapVlo, v_tile(s) + A_ATOM, &vfull[s], i * AK, 0, b);
                } else {
                mbar_expect_tx(&kfull[s], p.dbg_loads ? A_TILE / 2 : A_TILE);
                tma_load_2d(&mapKhi, k_tile(s), &kfull[s], 0, krow);
                tma_load_2d(&mapKhi, k_tile(s) + A_ATOM, &kfull[s], 32, krow);
                if (!p.dbg_loads) {
                tma_load_2d(&mapKlo, k_tile(s) + 2 * A_ATOM, &kfull[s], 0, krow);
                tma_load_2d(&mapKlo, k_tile(s) + 3 * A_ATOM, &kfull[s], 32, krow);
                }
                mbar_wait(&vfree[s], ph ^ 1, 150 + i);
                mbar_expect_tx(&vfull[s], p.dbg_loads ? A_TILE / 2 : A_TILE);
                tma_load_3d(&mapVhi, v_tile(s), &vfull[s], i * AK, 0, b);            // keys past Nc: zero fill
                tma_load_3d(&mapVhi, v_tile(s) + A_ATOM, &vfull[s], i * AK + 32, 0, b);
                if (!p.dbg_loads) {
                tma_load_3d(&mapVlo, v_tile(s) + 2 * A_ATOM, &vfull[s], i * AK, 0, b);
                tma_load_3d(&mapVlo, v_tile(s) + 3 * A_ATOM, &vfull[s], i * AK + 32, 0, b);
                }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer 1: S_j = Q K_j^T  (whole warp loops, one elected lane issues)
        // instruction descriptor: D=f32, A=B=tf32, K-major, N=64 (>>3 at bit 17), M=128 (>>4 at bit 24)
        const uint32_t idesc = (1u << 4) | ((F16 ? 0u : 2u) << 7) | ((F16 ? 0u : 2u) << 10) | ((uint32_t)(AK >> 3) << 17) | ((uint32_t)(AQ >> 4) << 24);
        long long w_kfull = 0, w_sfree = 0;
        DBG_T(t_start);
        mbar_wait(qready, 0, 200);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        DBG_T(t_q);
        for (int j = 0; j < nblk; ++j) {
            const int ks = j % A_STAGES, sb = j & 1;
            DBG_T(c0);
            mbar_wait(&kfull[ks], (j / A_STAGES) & 1, 210 + j);
            DBG_T(c1);
            mbar_wait(&s_free[sb], ((j >> 1) & 1) ^ 1, 220 + j);     // softmax has S_{j-2} in registers
            DBG_T(c2);
            w_kfull += c1 - c0; w_sfree += c2 - c1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                if constexpr (F16) issue_3xf16(tmem + C_S + 64 * sb, tmem + C_QHI, tmem + C_QLO, k_tile(ks), idesc);
                else issue_3xtf32(tmem + C_S + 64 * sb, tmem + C_QHI, tmem + C_QLO, k_tile(ks), idesc);
                umma_commit(&s_ready[sb]);
                umma_commit(&kfree[ks]);
            }
            __syncwarp();
        }
        if (A_PHASE_TIMERS && p.debug && lane == 0) {
            atomicAdd(&fc_attn_dbg[0], 1ull);
            atomicAdd(&fc_attn_dbg[1], (unsigned long long)(clock64() - t_start));
            atomicAdd(&fc_attn_dbg[2], (unsigned long long)w_kfull); atomicAdd(&fc_attn_dbg[3], (unsigned long long)w_sfree);
            atomicAdd(&fc_attn_dbg[11], (unsigned long long)(t_q - t_start));
        }
    } else if (warp == 2) {
        // ===================================================== MMA issuer 2: O_i = P_i V_i
        // (its own warp: a wait for P_i must not hold back S_{i+1}, and the other way round; each warp's
        // tcgen05.commit tracks the MMAs that warp issued)
        const uint32_t idesc = (1u << 4) | ((F16 ? 0u : 2u) << 7) | ((F16 ? 0u : 2u) << 10) | ((uint32_t)(AK >> 3) << 17) | ((uint32_t)(AQ >> 4) << 24);
        long long w_vfull = 0, w_ofree = 0, w_pready = 0;
        for (int i = 0; i < nblk; ++i) {
            const int vs = i % A_STAGES, ob = i & 1;
            DBG_T(c3);
            mbar_wait(&vfull[vs], (i / A_STAGES) & 1, 230 + i);
            DBG_T(c4);
            mbar_wait(&o_free[ob], ((i >> 1) & 1) ^ 1, 240 + i);       // O_{i-2} has been read back
            DBG_T(c5);
            const int pb = NPB == 2 ? (i & 1) : 0;
            mbar_wait(&p_ready[pb], NPB == 2 ? (i >> 1) & 1 : i & 1, 250 + i);
            DBG_T(c6);
            w_vfull += c4 - c3; w_ofree += c5 - c4; w_pready += c6 - c5;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                if constexpr (F16) issue_3xf16(tmem + C_O + 64 * ob, tmem + C_PHI + 64 * pb, tmem + C_PLO + 64 * pb, v_tile(vs), idesc);
                else issue_3xtf32(tmem + C_O + 64 * ob, tmem + C_PHI, tmem + C_PLO, v_tile(vs), idesc);
                umma_commit(&o_ready[ob]);
                umma_commit(&p_free[pb]);
                umma_commit(&vfree[vs]);
            }
            __syncwarp();
        }
        if (A_PHASE_TIMERS && p.debug && lane == 0) {
            atomicAdd(&fc_attn_dbg[4], (unsigned long long)w_vfull); atomicAdd(&fc_attn_dbg[5], (unsigned long long)w_ofree);
            atomicAdd(&fc_attn_dbg[6], (unsigned long long)w_pready);
        }
    } else {
        // ===================================================== Q converter + softmax + O accumulation
        // A_SPL threads per query row (in the A_SPL warps that may touch the row's TMEM lane quadrant): each owns
        // 64/A_SPL keys of every block and 64/A_SPL output dims.  Only the block's row maximum is exchanged (shared
        // memory + one named barrier per quadrant); the partial row sums are combined once at the end.
        constexpr int KS = AK / A_SPL;            // keys (and output dims) per thread
        const int quad = warp & 3;
        const int sub = (warp - 3) >> 2;
        const int r = quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const int off = sub * KS;                 // first key of the block / first output dim of this thread
        const float LOG2E = F16 ? 1.4426950408889634f / A_QS : 1.4426950408889634f;   // the scores carry Q's power-of-two scale
        const int bar_id = 1 + quad;

        // ---- Q: fp32 smem row -> scale -> (hi, lo) in TMEM   (dims [off, off + KS))
        mbar_wait(qfull, 0, 300);
        if constexpr (F16) {
            // this thread's KS dims (32: a whole 128-byte row of one atom; 16: half of it) as packed fp16 pairs: KS/2 columns hi, KS/2 lo
            uint32_t hi[KS / 2], lo[KS / 2];
            const unsigned char* atom = q_raw + (off >> 5) * (AQ * 128) + r * 128;
            const int j0 = (off & 31) >> 2;                            // first 16-byte chunk inside the row
            const float qs = p.scale * A_QS;
#pragma unroll
            for (int j = 0; j < KS / 4; ++j) {
                const float4 x = *reinterpret_cast<const float4*>(atom + (((j0 + j) ^ (r & 7)) << 4));
                const float xv[4] = {x.x * qs, x.y * qs, x.z * qs, x.w * qs};
#pragma unroll
                for (int e = 0; e < 4; e += 2) {
                    const uint32_t hp = pack_f16x2(xv[e], xv[e + 1]);
                    float h0, h1;
                    unpack_f16x2(hp, h0, h1);
                    hi[2 * j + (e >> 1)] = hp;
                    lo[2 * j + (e >> 1)] = pack_f16x2(xv[e] - h0, xv[e + 1] - h1);
                }
            }
            if constexpr (KS == 32) { tmem_st16(tmem + lane_addr + C_QHI + (off >> 1), hi); tmem_st16(tmem + lane_addr + C_QLO + (off >> 1), lo); }
            else                    { tmem_st8(tmem + lane_addr + C_QHI + (off >> 1), hi);  tmem_st8(tmem + lane_addr + C_QLO + (off >> 1), lo); }
        } else
#pragma unroll
        for (int c = 0; c < KS / 16; ++c) {
            uint32_t hi[16], lo[16];
            const int d0 = off + 16 * c;
            const unsigned char* atom = q_raw + (d0 >> 5) * (AQ * 128) + r * 128;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int j = ((d0 & 31) >> 2) + jj;                  // 16-byte chunk inside the 128-byte row
                const float4 x = *reinterpret_cast<const float4*>(atom + ((j ^ (r & 7)) << 4));
                const float xv[4] = {x.x * p.scale, x.y * p.scale, x.z * p.scale, x.w * p.scale};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    hi[4 * jj + e] = tf32_hi(xv[e]);
                    lo[4 * jj + e] = __float_as_uint(xv[e] - __uint_as_float(hi[4 * jj + e]));
                }
            }
            tmem_st16(tmem + lane_addr + C_QHI + d0, hi);
            tmem_st16(tmem + lane_addr + C_QLO + d0, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(qready);

        float o[KS];
#pragma unroll
        for (int j = 0; j < KS; ++j) o[j] = 0.f;
        float m = -INFINITY, l = 0.f, alpha_prev = 0.f;

        auto load_cols = [&](uint32_t col, float (&dst)[KS]) {
#pragma unroll
            for (int c = 0; c < KS / 16; ++c) {
                uint32_t t16[16];
                tmem_ld16(tmem + lane_addr + col + 16 * c, t16);
#pragma unroll
                for (int j = 0; j < 16; ++j) dst[16 * c + j] = __uint_as_float(t16[j]);
            }
        };

        long long w_sready = 0, w_pfree = 0, w_oready = 0;
        DBG_T(t_sm0);
        for (int i = 0; i < nblk; ++i) {
            const int sb = i & 1;
            float sv[KS];
            DBG_T(d0);
            mbar_wait(&s_ready[sb], (i >> 1) & 1, 310 + i);
            DBG_T(d1);
            w_sready += d1 - d0;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            load_cols(C_S + 64 * sb + off, sv);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&s_free[sb]);
            const int valid = p.Nc - i * AK - off;    // of this thread's keys, how many exist
            if (valid < KS) {
#pragma unroll
                for (int j = 0; j < KS; ++j) if (j >= valid) sv[j] = -INFINITY;
            }
            float mx4[4] = {sv[0], sv[1], sv[2], sv[3]};
#pragma unroll
            for (int j = 4; j < KS; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], sv[j]);
            float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            if (A_SPL > 1) {
                mx_sm[sb][sub][r] = mx;
                asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * A_SPL) : "memory");
#pragma unroll
                for (int u = 0; u < A_SPL; ++u) mx = fmaxf(mx, mx_sm[sb][u][r]);
            }
            const float m_new = fmaxf(m, mx);
            const float alpha = ex2_approx((m - m_new) * LOG2E);   // first block: exp2(-inf) = 0
            // F16: P is carried times 2^10 (clear of fp16's subnormal range) by shifting the exponent; the row sum l is accumulated
            // from the same scaled values, so the scale cancels in O / l
            const float mc = F16 ? fmaf(m_new, LOG2E, -10.0f) : m_new * LOG2E;
            float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < KS; j += 4) {
                // packed fp32 (FFMA2 / FADD2): two scores per instruction, per component the same operations as the scalar form
                const float2 t0 = __ffma2_rn(make_float2(sv[j], sv[j + 1]), make_float2(LOG2E, LOG2E), make_float2(-mc, -mc));
                const float2 t1 = __ffma2_rn(make_float2(sv[j + 2], sv[j + 3]), make_float2(LOG2E, LOG2E), make_float2(-mc, -mc));
                sv[j] = ex2_approx(t0.x); sv[j + 1] = ex2_approx(t0.y); sv[j + 2] = ex2_approx(t1.x); sv[j + 3] = ex2_approx(t1.y);
                const float2 s01 = __fadd2_rn(make_float2(sum4[0], sum4[1]), make_float2(sv[j], sv[j + 1]));
                const float2 s23 = __fadd2_rn(make_float2(sum4[2], sum4[3]), make_float2(sv[j + 2], sv[j + 3]));
                sum4[0] = s01.x; sum4[1] = s01.y; sum4[2] = s23.x; sum4[3] = s23.y;
            }
            l = fmaf(l, alpha, (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
            m = m_new;

            // P -> TMEM (hi, lo); the previous block's PV MMAs must have retired
            DBG_T(d2);
            const int pb = NPB == 2 ? (i & 1) : 0;
            mbar_wait(&p_free[pb], (NPB == 2 ? (i >> 1) & 1 : i & 1) ^ 1, 330 + i);
            DBG_T(d3);
            w_pfree += d3 - d2;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if constexpr (F16) {
                uint32_t hi[KS / 2], lo[KS / 2];
#pragma unroll
                for (int j = 0; j < KS / 2; ++j) {
                    const float p0 = sv[2 * j], p1 = sv[2 * j + 1];
                    const uint32_t hp = pack_f16x2(p0, p1);
                    float h0, h1;
                    unpack_f16x2(hp, h0, h1);
                    hi[j] = hp;
                    const float2 d2 = __fadd2_rn(make_float2(p0, p1), make_float2(-h0, -h1));
                    lo[j] = pack_f16x2(d2.x, d2.y);
                }
                if constexpr (KS == 32) { tmem_st16(tmem + lane_addr + C_PHI + 64 * pb + (off >> 1), hi); tmem_st16(tmem + lane_addr + C_PLO + 64 * pb + (off >> 1), lo); }
                else                    { tmem_st8(tmem + lane_addr + C_PHI + 64 * pb + (off >> 1), hi);  tmem_st8(tmem + lane_addr + C_PLO + 64 * pb + (off >> 1), lo); }
            } else
#pragma unroll
            for (int c = 0; c < KS / 16; ++c) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    hi[j] = tf32_hi(sv[16 * c + j]);
                    lo[j] = __float_as_uint(sv[16 * c + j] - __uint_as_float(hi[j]));
                }
                tmem_st16(tmem + lane_addr + C_PHI + off + 16 * c, hi);
                tmem_st16(tmem + lane_addr + C_PLO + off + 16 * c, lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&p_ready[pb]);

            // fold in the PREVIOUS block's O (its MMAs ran while this block's softmax was computed)
            if (i > 0) {
                const int ob = (i - 1) & 1;
                float ov[KS];
                DBG_T(d4);
                mbar_wait(&o_ready[ob], ((i - 1) >> 1) & 1, 350 + i);
                DBG_T(d5);
                w_oready += d5 - d4;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                load_cols(C_O + 64 * ob + off, ov);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(&o_free[ob]);
#pragma unroll
                for (int j = 0; j < KS; j += 2) {
                    const float2 o2 = __ffma2_rn(make_float2(o[j], o[j + 1]), make_float2(alpha_prev, alpha_prev), make_float2(ov[j], ov[j + 1]));
                    o[j] = o2.x; o[j + 1] = o2.y;
                }
            }
            alpha_prev = alpha;
        }
        if (A_PHASE_TIMERS && p.debug && threadIdx.x == 96) {
            atomicAdd(&fc_attn_dbg[7], (unsigned long long)w_sready); atomicAdd(&fc_attn_dbg[8], (unsigned long long)w_pfree);
            atomicAdd(&fc_attn_dbg[9], (unsigned long long)w_oready); atomicAdd(&fc_attn_dbg[10], (unsigned long long)(clock64() - t_sm0));
        }
        {
            const int ob = (nblk - 1) & 1;
            float ov[KS];
            mbar_wait(&o_ready[ob], ((nblk - 1) >> 1) & 1, 390);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            load_cols(C_O + 64 * ob + off, ov);
#pragma unroll
            for (int j = 0; j < KS; ++j) o[j] = fmaf(o[j], alpha_prev, ov[j]);
        }
        if (A_SPL > 1) {
            // all partial sums share the same running maximum: the row sum is their plain sum
            mx_sm[nblk & 1][sub][r] = l;      // (the buffer the last block did not use)
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * A_SPL) : "memory");
            l = 0.f;
#pragma unroll
            for (int u = 0; u < A_SPL; ++u) l += mx_sm[nblk & 1][u][r];
        }
        if (q0 + r < p.N) {
            const float inv = 1.0f / l;
            float4* dst = reinterpret_cast<float4*>(p.out + ((size_t)b * p.N + q0 + r) * p.ldo + off);
#pragma unroll
            for (int j = 0; j < KS / 4; ++j)
                dst[j] = make_float4(o[4 * j] * inv, o[4 * j + 1] * inv, o[4 * j + 2] * inv, o[4 * j + 3] * inv);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
    }
}

// kv [B*Nc][ldkv] (K = columns 0..63, V = 64..127)  ->  Khi, Klo [B*Nc][64]  and  Vt hi, lo [B][64][Ncp]: TF32 splits stored as
// fp32, or (F16) fp16 hi / fp16 (x - hi)
template <bool F16>
__global__ void __launch_bounds__(256)
kv_split_kernel(const float* __restrict__ kv, int ldkv, int Nc, int Ncp, void* __restrict__ khi_, void* __restrict__ klo_,
                void* __restrict__ vthi_, void* __restrict__ vtlo_) {
    __shared__ float tile[32][AD + 1];
    const int b = blockIdx.y, key0 = blockIdx.x * 32, tid = threadIdx.x;
    auto put = [](void* hi_, void* lo_, size_t i, float x) {
        if (F16) {
            const __half h = __float2half_rn(x);
            static_cast<__half*>(hi_)[i] = h;
            static_cast<__half*>(lo_)[i] = __float2half_rn(x - __half2float(h));
        } else {
            const float h = __uint_as_float(tf32_hi(x));
            static_cast<float*>(hi_)[i] = h;
            static_cast<float*>(lo_)[i] = x - h;
        }
    };
    for (int idx = tid; idx < 32 * AD; idx += 256) {
        const int kk = idx >> 6, d = idx & 63;
        const int key = key0 + kk;
        if (key < Nc) {
            const float* src = kv + ((size_t)b * Nc + key) * ldkv;
            put(khi_, klo_, ((size_t)b * Nc + key) * AD + d, src[d]);
            tile[kk][d] = src[AD + d];
        } else {
            tile[kk][d] = 0.f;
        }
    }
    __syncthreads();
    for (int idx = tid; idx < 32 * AD; idx += 256) {
        const int d = idx >> 5, kk = idx & 31;
        const int key = key0 + kk;
        if (key < Ncp) put(vthi_, vtlo_, ((size_t)b * AD + d) * Ncp + key, tile[kk][d]);
    }
}

// ----------------------------------------------------------------------------- host: tensor maps (cached)
struct AMapKey {
    const void* base; uint64_t d0, d1, d2, s1, s2; uint32_t b0, b1; int f16;
    bool operator==(const AMapKey& o) const {
        return base == o.base && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && s1 == o.s1 && s2 == o.s2 && b0 == o.b0 && b1 == o.b1 && f16 == o.f16;
    }
};
std::vector<std::pair<AMapKey, CUtensorMap>> g_amaps;
std::mutex g_amaps_mu;

// fp32 (or fp16) tensor, dims (d0 fastest, d1, d2), strides in ELEMENTS of d1 / d2, box = b0 x b1 x 1, 128B swizzle, OOB -> 0
bool get_amap(const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2, uint32_t b0, uint32_t b1,
              CUtensorMap* out, int f16 = 0) {
    AMapKey key{base, d0, d1, d2, s1, s2, b0, b1, f16};
    std::lock_guard<std::mutex> lk(g_amaps_mu);
    for (auto& kv : g_amaps) if (kv.first == key) { *out = kv.second; return true; }
    FcEncodeTiledFn enc = fc_get_encode_fn();
    if (!enc) return false;
    const cuuint32_t rank = d2 > 0 ? 3 : 2;
    cuuint64_t dims[3] = {d0, d1, d2 > 0 ? d2 : 1};
    const cuuint64_t esz = f16 ? 2 : 4;
    cuuint64_t strides[2] = {s1 * esz, s2 * esz};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap m;
    if (enc(&m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (g_amaps.size() > 256) g_amaps.clear();
    g_amaps.emplace_back(key, m);
    *out = m;
    return true;
}

}  // namespace

int64_t fc_attention_tc_scratch_floats(int B, int Nc) {
    const int64_t Ncp = fc_round_up(Nc, 8);
    return 2 * fc_round_up_ll((int64_t)B * Nc * AD, 32) + 2 * fc_round_up_ll((int64_t)B * AD * Ncp, 32) + 64;
}

// fmt 0: TF32 hi/lo as fp32, Ncp = Nc rounded to 4 floats; fmt 1: fp16 hi/lo in the first half of each region, Ncp = Nc rounded
// to 8 halfs (TMA row strides are multiples of 16 bytes).  All pointers are 128-byte aligned when `scratch` is.
void fc_attention_tc_scratch_layout(int B, int Nc, float* scratch, float** khi, float** klo, float** vthi, float** vtlo,
                                    int* ncp, int fmt) {
    const int Ncp = fc_round_up(Nc, fmt ? 8 : 4);
    *khi = scratch;
    *klo = *khi + fc_round_up_ll((int64_t)B * Nc * AD, 32);
    *vthi = *klo + fc_round_up_ll((int64_t)B * Nc * AD, 32);
    *vtlo = *vthi + fc_round_up_ll((int64_t)B * AD * fc_round_up(Nc, 8), 32);
    *ncp = Ncp;
}

namespace {
template <bool F16>
int launch_attention_tc(const float* q, int ldq, const float* kv, int ldkv, float* out, int ldo, int B, int N, int Nc, int d,
                        float scale, float* scratch, int presplit, cudaStream_t stream) {
    float *khi, *klo, *vthi, *vtlo;
    int Ncp;
    fc_attention_tc_scratch_layout(B, Nc, scratch, &khi, &klo, &vthi, &vtlo, &Ncp, F16 ? 1 : 0);
    CUtensorMap mQ, mKh, mKl, mVh, mVl;
    if (!get_amap(q, AD, (uint64_t)B * N, 0, (uint64_t)ldq, 0, 32, AQ, &mQ)) return FC_ERR_CUDA;
    const uint32_t bx = F16 ? 64 : 32;      // box width in elements = one 128-byte swizzle row
    if (!get_amap(khi, AD, (uint64_t)B * Nc, 0, AD, 0, bx, AK, &mKh, F16)) return FC_ERR_CUDA;
    if (!get_amap(klo, AD, (uint64_t)B * Nc, 0, AD, 0, bx, AK, &mKl, F16)) return FC_ERR_CUDA;
    if (!get_amap(vthi, (uint64_t)Nc, AD, (uint64_t)B, (uint64_t)Ncp, (uint64_t)AD * Ncp, bx, AD, &mVh, F16)) return FC_ERR_CUDA;
    if (!get_amap(vtlo, (uint64_t)Nc, AD, (uint64_t)B, (uint64_t)Ncp, (uint64_t)AD * Ncp, bx, AD, &mVl, F16)) return FC_ERR_CUDA;
    int dev = 0;
    FC_CUDA_OK(cudaGetDevice(&dev));
    static std::atomic<uint64_t> configured{0};     // the dynamic shared-memory opt-in is per device (and per instantiation)
    const uint64_t bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        FC_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ACfg<F16>::SMEM));
        configured.fetch_or(bit, std::memory_order_release);
    }
    FcProfScope prof(FC_CLS_ATTENTION, 4.0 * B * (double)N * Nc * d, 4.0 * B * ((double)N * d * 2 + (double)Nc * d * 2), stream);
    if (!presplit) {
        kv_split_kernel<F16><<<dim3((Ncp + 31) / 32, B), 256, 0, stream>>>(kv, ldkv, Nc, Ncp, khi, klo, vthi, vtlo);
        fc_count_launch();
    }
    static int dbg_loads = -1;
    if (dbg_loads < 0) { const char* e = getenv("FC_ATTN_DBG_LOADS"); dbg_loads = (e && e[0] == '1') ? 1 : 0; }
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("FC_ATTN_DEBUG"); dbg = (e && e[0] == '1') ? 1 : 0; }
    AttnTcParams p{out, ldo, N, Nc, (Nc + AK - 1) / AK, scale, dbg, dbg_loads};
    attention_tc_kernel<F16><<<dim3((N + AQ - 1) / AQ, B), A_THREADS, ACfg<F16>::SMEM, stream>>>(mQ, mKh, mKl, mVh, mVl, p);
    fc_count_launch();
    FC_LAUNCH_OK();
    return FC_OK;
}
}  // namespace

// presplit != 0: the scratch already holds k / v^T hi/lo (written by the to_kv GEMM's FC_EPI_KVSPLIT epilogue), kv unused.
// fmt: 0 = 3xTF32, 1 = 3xFP16 operands.
int fc_launch_cross_attention_tc(const float* q, int ldq, const float* kv, int ldkv, float* out, int ldo, int B, int N,
                                 int Nc, int d, float scale, float* scratch, int presplit, int fmt, cudaStream_t stream) {
    FC_REQUIRE(q && (kv || presplit) && out && scratch && B > 0 && N > 0 && Nc > 0);
    FC_REQUIRE((int64_t)B * N < (1ll << 31) - AQ && (int64_t)B * Nc < (1ll << 31) - AK);   // TMA row coordinates are int32
    if (d != AD) return FC_ERR_UNSUPPORTED;
    FC_REQUIRE((ldq & 3) == 0 && (ldo & 3) == 0 && (presplit || ldkv >= 2 * AD) && ldq >= AD && ldo >= AD && B <= 65535);
    FC_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(scratch) & 127) == 0);
    return fmt ? launch_attention_tc<true>(q, ldq, kv, ldkv, out, ldo, B, N, Nc, d, scale, scratch, presplit, stream)
               : launch_attention_tc<false>(q, ldq, kv, ldkv, out, ldo, B, N, Nc, d, scale, scratch, presplit, stream);
}

extern "C" __attribute__((visibility("default"))) int fc_debug_attn_phases(unsigned long long* out16) {
    if (cudaMemcpyFromSymbol(out16, fc_attn_dbg, 16 * sizeof(unsigned long long)) != cudaSuccess) return FC_ERR_CUDA;
    unsigned long long z[16] = {};
    if (cudaMemcpyToSymbol(fc_attn_dbg, z, sizeof(z)) != cudaSuccess) return FC_ERR_CUDA;
    return FC_OK;
}

extern "C" __attribute__((visibility("default"))) int64_t fc_cross_attention_tc_scratch_bytes(int B, int Nc) {
    if (B <= 0 || Nc <= 0) return FC_ERR_INVALID_ARG;
    return fc_attention_tc_scratch_floats(B, Nc) * 4;
}

extern "C" __attribute__((visibility("default"))) int fc_cross_attention_tc(const float* q, int ldq, const float* kv, int ldkv,
                                                                             float* out, int ldo, int B, int N, int Nc, int d,
                                                                             float scale, void* scratch, int64_t scratch_bytes,
                                                                             fc_stream_t stream) {
    FC_REQUIRE(scratch && scratch_bytes >= fc_cross_attention_tc_scratch_bytes(B, Nc));
    return fc_launch_cross_attention_tc(q, ldq, kv, ldkv, out, ldo, B, N, Nc, d, scale, static_cast<float*>(scratch), 0, 0,
                                        (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int fc_cross_attention_tc_f16(const float* q, int ldq, const float* kv, int ldkv,
                                                                                 float* out, int ldo, int B, int N, int Nc, int d,
                                                                                 float scale, void* scratch, int64_t scratch_bytes,
                                                                                 fc_stream_t stream) {
    FC_REQUIRE(scratch && scratch_bytes >= fc_cross_attention_tc_scratch_bytes(B, Nc));
    return fc_launch_cross_attention_tc(q, ldq, kv, ldkv, out, ldo, B, N, Nc, d, scale, static_cast<float*>(scratch), 0, 1,
                                        (cudaStream_t)stream);
}
