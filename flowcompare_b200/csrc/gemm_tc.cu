// fp32-faithful GEMM on the 5th-gen tensor cores: 3xTF32 error-compensated tcgen05.mma with TMEM accumulators,
// operands staged by TMA (128B swizzle), the same fused epilogues as the FFMA kernel (gemm.cu).
//
//   C = A W^T  with  A = Ahi + Alo,  W = Whi + Wlo  (TF32 splits)   =>   D = Ahi Whi + Alo Whi + Ahi Wlo
// (the dropped Alo Wlo term is ~2^-22 relative).  W is split once at pack time; A (an activation that only
// exists in fp32) is split ON CHIP: TMA lands the raw fp32 tile in shared memory, four converter warps rewrite
// it in place as the TF32 hi part and write the lo part beside it (same swizzled offsets), then the single MMA
// thread issues three tcgen05.mma per 8-wide k-step against the TMEM accumulator.  Nothing but the fp32
// activations themselves ever travels through HBM.
//
// CTA = one 128 x BN output tile (BN <= 256, runtime), 6 warps:
//   warp 0  TMA producer (one lane)          warp 1  TMEM allocator + MMA issuer (one lane)
//   warps 2-5  fp32 -> (hi, lo) converters during the main loop, then the epilogue (one thread per row,
//              tcgen05.ld 32x32b, fused bias / GELU / residual / LayerNorm-q / coupling / augment epilogues)
// Two shared-memory stages of {A hi, A lo, W hi, W lo} x 32 k, mbarrier pipeline:
//   full[s]  (TMA bytes landed)  ->  conv[s] (128 converter arrivals)  ->  tcgen05.commit -> empty[s]
#include "gemm.cuh"
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <cstdio>
#include <cstdlib>

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                     // 32 fp32 = one 128-byte swizzle row
constexpr int TC_STAGES = 3;
constexpr int TC_THREADS = 192;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;  // 16 KB
constexpr int TC_TMEM_COLS = 512;
// The tensor core TRUNCATES when it adds a product block to the TMEM accumulator (measured: ~0.25 ulp of
// one-sided error per tcgen05.mma, i.e. a systematic shrink of every output that grows with K and shows up
// as a constant offset in the log-density).  Two counter-measures keep the result at fp32-FFMA quality:
//   * the two small compensation products (Alo Whi, Ahi Wlo) go to their OWN accumulator, where the
//     truncation happens at 2^-11 of the magnitude;
//   * the main product Ahi Whi is spread round-robin by k-block over TC_MAIN_ACC accumulators, so each sees
//     1/3 of the sequential truncations; the epilogue adds the four partial sums in round-to-nearest fp32.
// 4 accumulators x BN (<= 128) columns = the whole 512-column TMEM.
constexpr int TC_MAIN_ACC = 3;
constexpr int TC_ACC_STRIDE = 128;             // TMEM columns between accumulators

struct TcParams {
    GemmArgs g;
    int BN;        // N tile (multiple of 16, <= 256)
    int T1, T2;    // k-blocks of segment 1 / 2
    int passes;    // 3 = 3xTF32, 1 = plain TF32 (debug)
};

// ----------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a pipeline bug must surface as a trapped kernel with a message, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return;
    }
    printf("flowcompare_b200 gemm_tc: mbarrier wait timed out (tag %d, block %d,%d thread %d, parity %u)\n", tag,
           blockIdx.x, blockIdx.y, threadIdx.x, parity);
    __trap();
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, void* smem_dst, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row groups are 1024 B apart (SBO), version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address  [0,14)
    d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset [32,46)
    d |= (uint64_t)1 << 46;                              // descriptor version
    d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"   // same asm statement: no use of r[] can be scheduled before the wait
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

// ----------------------------------------------------------------------------- kernel
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapA2,
               const __grid_constant__ CUtensorMap mapWhi, const __grid_constant__ CUtensorMap mapWlo, const TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[3 * TC_STAGES + 1];
    __shared__ uint32_t tmem_base_slot;

    const uint32_t raw = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);   // SWIZZLE_128B tiles need 1024 B alignment
    const int BN = p.BN;
    const int w_bytes = BN * TC_BK * 4;
    const int stage_bytes = 2 * TC_A_BYTES + 2 * w_bytes;
    auto a_hi = [&](int s) { return smem + s * stage_bytes; };
    auto a_lo = [&](int s) { return smem + s * stage_bytes + TC_A_BYTES; };
    auto w_hi = [&](int s) { return smem + s * stage_bytes + 2 * TC_A_BYTES; };
    auto w_lo = [&](int s) { return smem + s * stage_bytes + 2 * TC_A_BYTES + w_bytes; };
    uint64_t* full = bars;
    uint64_t* conv = bars + TC_STAGES;
    uint64_t* empty = bars + 2 * TC_STAGES;
    uint64_t* accum = bars + 3 * TC_STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TC_BM;
    const int n_tile = blockIdx.x;
    const int T = p.T1 + p.T2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&conv[s], 128); mbar_init(&empty[s], 1); }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "n"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            for (int t = 0; t < T; ++t) {
                const int s = t % TC_STAGES;
                const uint32_t ph = (t / TC_STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1, 100 + t);
                mbar_expect_tx(&full[s], (uint32_t)(TC_A_BYTES + (p.passes == 3 ? 2 : 1) * w_bytes));
                if (t < p.T1) tma_load_2d(&mapA1, a_hi(s), &full[s], t * TC_BK, m0);
                else          tma_load_2d(&mapA2, a_hi(s), &full[s], (t - p.T1) * TC_BK, m0);
                tma_load_2d(&mapWhi, w_hi(s), &full[s], t * TC_BK, n_tile * BN);
                if (p.passes == 3) tma_load_2d(&mapWlo, w_lo(s), &full[s], t * TC_BK, n_tile * BN);
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            // instruction descriptor: D=f32 (bits 4-5 = 1), A=B=tf32 (2), both K-major, N>>3 at bit 17, M>>4 at bit 24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            for (int t = 0; t < T; ++t) {
                const int s = t % TC_STAGES;
                const uint32_t ph = (t / TC_STAGES) & 1;
                mbar_wait(&full[s], ph, 200 + t);
                mbar_wait(&conv[s], ph, 300 + t);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t dah = make_kmajor_sw128_desc(smem_u32(a_hi(s)));
                const uint64_t dal = make_kmajor_sw128_desc(smem_u32(a_lo(s)));
                const uint64_t dbh = make_kmajor_sw128_desc(smem_u32(w_hi(s)));
                const uint64_t dbl = make_kmajor_sw128_desc(smem_u32(w_lo(s)));
                const uint32_t d_main = tmem_d + (uint32_t)((t % TC_MAIN_ACC) * TC_ACC_STRIDE);
                const uint32_t d_corr = tmem_d + (uint32_t)(TC_MAIN_ACC * TC_ACC_STRIDE);
                const bool main_first = t < TC_MAIN_ACC;   // first k-block landing in this main accumulator
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k) {
                    const uint64_t adv = (uint64_t)(k * 32 >> 4);   // 8 tf32 = 32 bytes per k-step
                    if (p.passes == 3) {
                        umma_tf32(d_corr, dal + adv, dbh + adv, idesc, (t | k) != 0);
                        umma_tf32(d_corr, dah + adv, dbl + adv, idesc, 1);
                    }
                    umma_tf32(d_main, dah + adv, dbh + adv, idesc, !(main_first && k == 0));
                }
                umma_commit(&empty[s]);      // stage s may be refilled once these MMAs retire
            }
            umma_commit(accum);              // accumulator complete
        }
    } else {
        // ===================================================== converters, then epilogue
        const int cid = threadIdx.x - 64;   // 0..127
        for (int t = 0; t < T; ++t) {
            const int s = t % TC_STAGES;
            const uint32_t ph = (t / TC_STAGES) & 1;
            mbar_wait(&full[s], ph, 400 + t);
            if (p.passes == 3) {
                float4* hi = reinterpret_cast<float4*>(a_hi(s));
                float4* lo = reinterpret_cast<float4*>(a_lo(s));
#pragma unroll
                for (int i = 0; i < TC_A_BYTES / 16 / 128; ++i) {
                    const int c = cid + i * 128;
                    const float4 x = hi[c];
                    float4 h, l;
                    uint32_t u;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.x)); h.x = __uint_as_float(u); l.x = x.x - h.x;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.y)); h.y = __uint_as_float(u); l.y = x.y - h.y;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.z)); h.z = __uint_as_float(u); l.z = x.z - h.z;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.w)); h.w = __uint_as_float(u); l.w = x.w - h.w;
                    hi[c] = h;
                    lo[c] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
            }
            mbar_arrive(&conv[s]);
        }
        // ---- epilogue: TMEM lane = row within the tile; warp w may touch lanes 32*(w%4) .. +31
        mbar_wait(accum, 0, 500);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const GemmArgs& a = p.g;
        const int lane_base = (warp & 3) * 32;
        const int row = m0 + lane_base + lane;
        const bool row_ok = row < a.M;
        const int n0 = n_tile * BN;
        float ldj = 0.f;
        float mu = 0.f, rstd = 0.f;
        if (a.epi == FC_EPI_LNQ && row_ok) { mu = a.row_mu[row]; rstd = a.row_rstd[row]; }
        const float* bias_row = a.bias;
        if (a.bias && a.bias_group > 0 && row_ok) bias_row = a.bias + (size_t)(row / a.bias_group) * a.bias_ld;
        const int n_main = T < TC_MAIN_ACC ? T : TC_MAIN_ACC;   // main accumulators that were written
        for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t r[16];
            float v[16];
            __syncwarp();
            const uint32_t tbase = tmem_d + ((uint32_t)lane_base << 16) + (uint32_t)c0;
            tmem_ld16(tbase, r);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
            for (int m = 1; m < n_main; ++m) {
                tmem_ld16(tbase + (uint32_t)(m * TC_ACC_STRIDE), r);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(r[j]);
            }
            if (p.passes == 3) {
                tmem_ld16(tbase + (uint32_t)(TC_MAIN_ACC * TC_ACC_STRIDE), r);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(r[j]);
            }
            const int col = n0 + c0;
            if (!row_ok || col >= a.N) continue;
            if (a.epi == FC_EPI_LNQ) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (col + j < a.N) v[j] = rstd * (v[j] - mu * a.csum[col + j]) + a.bias[col + j];
            } else if (bias_row) {
#pragma unroll
                for (int j = 0; j < 16; ++j) if (col + j < a.N) v[j] += bias_row[col + j];
            }
            if (a.epi == FC_EPI_STORE || a.epi == FC_EPI_LNQ) {
                if (a.res) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (col + j < a.N) {
                            const float rv = a.res[(size_t)row * a.ldres + col + j];
                            v[j] = a.res_scale ? fmaf(a.res_scale[col + j], rv, v[j]) : v[j] + rv;
                        }
                }
                if (a.act == FC_ACT_GELU) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fc_gelu_erf(v[j]);
                } else if (a.act == FC_ACT_LRELU) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fc_leaky_relu02(v[j]);
                }
                float* dst = a.C + (size_t)row * a.ldc + col;
                if (col + 15 < a.N && (a.ldc & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (col + j < a.N) dst[j] = v[j];
                }
            } else if (a.epi == FC_EPI_COUPLING) {
                // reference models/affine_coupling.py:40-46 (see gemm.cu for the arithmetic notes)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (col + 2 * q + 1 < a.N) {
                        const int j = (col >> 1) + q;
                        const float sig = 1.0f / (1.0f + expf(-v[2 * q]));
                        const float sc = (2.0f * sig - 1.0f) + 1.0f;
                        float* xp = a.x + (size_t)row * a.ldx + a.col0 + j;
                        *xp = fmaf(*xp, sc, v[2 * q + 1]);
                        ldj += logf(sc);
                    }
                }
            } else if (a.epi == FC_EPI_AUGMENT) {
                // reference models/distributions.py:128-153 + models/augmenter.py:49-63
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (col + 2 * q + 1 < a.N) {
                        const int j = (col >> 1) + q;
                        const float e = a.eps[(size_t)row * a.ld_eps + j];
                        a.x[(size_t)row * a.ldx + a.col0 + j] = fmaf(expf(v[2 * q + 1]), e, v[2 * q]);
                        ldj += fmaf(0.5f * e, e, v[2 * q + 1]) + 0.91893853320467274178f;
                    }
                }
            }
        }
        if ((a.epi == FC_EPI_COUPLING || a.epi == FC_EPI_AUGMENT) && row_ok) {
            float* pp = a.part + (size_t)n_tile * a.M + row;
            if (a.epi == FC_EPI_COUPLING) *pp += ldj; else *pp = ldj;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TC_TMEM_COLS));
    }
}

// ----------------------------------------------------------------------------- host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

struct MapKey {
    const void* base; uint64_t inner, outer, stride; uint32_t box_outer;
    bool operator==(const MapKey& o) const {
        return base == o.base && inner == o.inner && outer == o.outer && stride == o.stride && box_outer == o.box_outer;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = std::hash<const void*>()(k.base);
        h ^= std::hash<uint64_t>()(k.inner * 1315423911u + k.outer) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
        h ^= std::hash<uint64_t>()(k.stride * 31 + k.box_outer) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
        return h;
    }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
std::mutex g_maps_mu;

// fp32 [outer][inner] row-major with row stride `stride_floats`; box = 32 x box_outer, 128B swizzle, OOB -> 0
bool get_map(const float* base, uint64_t inner, uint64_t outer, uint64_t stride_floats, uint32_t box_outer, CUtensorMap* out) {
    MapKey key{base, inner, outer, stride_floats, box_outer};
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return true; }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {stride_floats * 4};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "flowcompare_b200: cuTensorMapEncodeTiled failed (%d): base=%p inner=%llu outer=%llu stride=%llu box=%ux%u\n",
                (int)r, (const void*)base, (unsigned long long)inner, (unsigned long long)outer,
                (unsigned long long)stride_floats, (unsigned)TC_BK, box_outer);
        return false;
    }
    if (g_maps.size() > 65536) g_maps.clear();
    g_maps.emplace(key, m);
    *out = m;
    return true;
}

}  // namespace

bool fc_gemm_tc_supported(const GemmArgs& a) {
    if (!a.Whi || !a.Wlo || a.ldk <= 0) return false;
    if ((a.lda1 & 3) || (reinterpret_cast<uintptr_t>(a.A1) & 15)) return false;
    if (a.K2 && ((a.lda2 & 3) || (reinterpret_cast<uintptr_t>(a.A2) & 15))) return false;
    if ((reinterpret_cast<uintptr_t>(a.Whi) & 15) || (reinterpret_cast<uintptr_t>(a.Wlo) & 15)) return false;
    if (a.M < 1 || a.N < 16) return false;
    return true;
}

int fc_launch_gemm_tc(const GemmArgs& a, cudaStream_t stream) {
    FC_REQUIRE(fc_gemm_tc_supported(a));
    if (a.epi == FC_EPI_STORE || a.epi == FC_EPI_LNQ) FC_REQUIRE(a.C != nullptr);
    if (a.epi == FC_EPI_LNQ) FC_REQUIRE(a.row_mu && a.row_rstd && a.csum && a.bias && a.bias_group == 0);
    if (a.epi == FC_EPI_COUPLING || a.epi == FC_EPI_AUGMENT) FC_REQUIRE(a.x && a.part && (a.N % 4) == 0);
    TcParams p;
    p.g = a;
    p.BN = fc_tc_bn(a.N);
    p.T1 = fc_tc_kpad(a.K1) / TC_BK;
    p.T2 = a.K2 ? fc_tc_kpad(a.K2) / TC_BK : 0;
    static int passes_env = -1;
    if (passes_env < 0) { const char* e = getenv("FC_TC_PASSES"); passes_env = (e && e[0] == '1') ? 1 : 3; }
    p.passes = passes_env;
    FC_REQUIRE(a.ldk == (p.T1 + p.T2) * TC_BK);
    const int n_tiles = fc_tc_n_tiles(a.N);
    CUtensorMap mA1, mA2, mWh, mWl;
    if (!get_map(a.A1, (uint64_t)a.K1, (uint64_t)a.M, (uint64_t)a.lda1, TC_BM, &mA1)) return FC_ERR_CUDA;
    if (a.K2) { if (!get_map(a.A2, (uint64_t)a.K2, (uint64_t)a.M, (uint64_t)a.lda2, TC_BM, &mA2)) return FC_ERR_CUDA; }
    else mA2 = mA1;
    if (!get_map(a.Whi, (uint64_t)a.ldk, (uint64_t)n_tiles * p.BN, (uint64_t)a.ldk, (uint32_t)p.BN, &mWh)) return FC_ERR_CUDA;
    if (!get_map(a.Wlo, (uint64_t)a.ldk, (uint64_t)n_tiles * p.BN, (uint64_t)a.ldk, (uint32_t)p.BN, &mWl)) return FC_ERR_CUDA;
    const int smem = TC_STAGES * (2 * TC_A_BYTES + 2 * p.BN * TC_BK * 4) + 1024;
    static bool configured = false;
    if (!configured) {
        // 2 stages x (2 x 16 KB A + 2 x 32 KB W) + 1 KB alignment slack; static smem (barriers) comes on top
        FC_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    dim3 grid(n_tiles, (a.M + TC_BM - 1) / TC_BM);
    FcProfScope prof(FC_CLS_GEMM_TC, 2.0 * a.M * a.N * (a.K1 + a.K2),
                     4.0 * ((double)a.M * (a.K1 + a.K2) + (double)a.N * (a.K1 + a.K2) + (double)a.M * a.N), stream);
    gemm_tc_kernel<<<grid, TC_THREADS, smem, stream>>>(mA1, mA2, mWh, mWl, p);
    fc_count_launch();
    FC_LAUNCH_OK();
    return FC_OK;
}
