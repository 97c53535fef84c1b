// The conditional normalizing flow: `Flow.log_prob` of reference models/transform.py:70-76 over the
// transform list of model_initialization.py:134-152, as a stream of fused GEMM / attention launches.
//
// Per coupling layer (attention configs), with M = B*N rows:
//   pre-attention MLP (reference models/cif_block.py:14-20, models/nets.py:19-30)   4 GEMMs (GELU/residual fused)
//   LayerNorm row statistics (models/perceiver.py:26-35)                            1 small kernel
//   to_q with LayerNorm folded into the epilogue (perceiver.py:109)                 1 GEMM
//   to_kv on the context (perceiver.py:110)                                         1 GEMM
//   softmax(q k^T / 8) v, flash style (perceiver.py:111-113)                        1 kernel
//   coupling MLP; attention out-projection (perceiver.py:95-96) and the concat with x1 / extra
//   context (transform.py:49-50, affine_coupling.py:34-35) folded into its first layer           4 GEMMs
//     last GEMM's epilogue = sigmoid scale, y2 = x2*s + t in place, sum log s -> partial log-det
//   ActNorm + LinearLU folded into one 300x300 GEMM (act_norm.py:37-43, permuters.py:164-169)     1 GEMM
// The running log-det never leaves fp32 registers/partials until the final reduction.
#include "model.cuh"
#include <cstdlib>
#include <new>

namespace {

__global__ void copy_cols_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd,
                                 long long M, int cols) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * cols) return;
    const long long r = i / cols; const int c = (int)(i % cols);
    dst[r * ldd + c] = src[r * lds + c];
}

// one warp per row: mean and 1/sqrt(var + eps) (biased variance, two passes in registers).  VEC: 16-byte loads
// (width % 4 == 0, 16-byte aligned rows) -- the scalar form reached only ~1.8 TB/s.
template <bool VEC>
__global__ void ln_stats_kernel(const float* __restrict__ h, int ldh, int M, int width, float eps,
                                float* __restrict__ mu, float* __restrict__ rstd) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* p = h + (size_t)row * ldh;
    float v[16];
    float s = 0.f;
    if (VEC) {
        const float4* p4 = reinterpret_cast<const float4*>(p);
        const int n4 = width >> 2;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = i * 32 + lane;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < n4) x = __ldg(p4 + c);
            v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
            s += (x.x + x.y) + (x.z + x.w);
        }
    } else {
        const int per = (width + 31) / 32;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            v[i] = 0.f;
            if (i < per) { const int c = i * 32 + lane; if (c < width) { v[i] = p[c]; s += v[i]; } }
        }
    }
    s = fc_warp_sum(s);
    const float m = s / (float)width;
    float q = 0.f;
    if (VEC) {
        const int n4 = width >> 2;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i * 32 + lane < n4) {
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float d = v[4 * i + e] - m; q = fmaf(d, d, q); }
            }
    } else {
        const int per = (width + 31) / 32;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < per) { const int c = i * 32 + lane; if (c < width) { const float d = v[i] - m; q = fmaf(d, d, q); } }
    }
    q = fc_warp_sum(q);
    if (lane == 0) { mu[row] = m; rstd[row] = rsqrtf(q / (float)width + eps); }
}

// cbA[b] = [extra_b (if extra) | ctx_b (if global)], zero padded to lda
__global__ void build_cb_input_kernel(const float* __restrict__ extra, const float* __restrict__ ctx, int E,
                                      int has_extra, int is_global, int B, int lda, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * lda) return;
    const int b = i / lda, c = i % lda;
    float v = 0.f;
    if (has_extra && c == 0) v = extra[b];
    else if (is_global) { const int e = c - has_extra; if (e >= 0 && e < E) v = ctx[(size_t)b * E + e]; }
    out[i] = v;
}

// log_prob[row] = sum(augment partials) + sum(coupling partials) + const + sum_d(-0.5 z_d^2 - 0.5 log 2pi)
// (reference models/distributions.py:192-195 for the base density).  One warp per row.
__global__ void finalize_kernel(const float* __restrict__ z, int ldz, int D, const float* __restrict__ apart,
                                int n_apart, const float* __restrict__ cpart, int n_cpart, int M, float ldj_const,
                                float* __restrict__ log_prob) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* p = z + (size_t)row * ldz;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) { const float v = p[c]; s += -0.91893853320467274178f - 0.5f * v * v; }
    s = fc_warp_sum(s);
    if (lane == 0) {
        float acc = 0.f;
        for (int i = 0; i < n_apart; ++i) acc += apart[(size_t)i * M + row];
        for (int i = 0; i < n_cpart; ++i) acc += cpart[(size_t)i * M + row];
        log_prob[row] = (acc + ldj_const) + s;
    }
}

}  // namespace

int fc_launch_ln_stats(const float* h, int ldh, int M, int width, float eps, float* mu, float* rstd, cudaStream_t s) {
    FC_REQUIRE(width <= 512);
    const int wpb = 8;
    if ((width & 3) == 0 && (ldh & 3) == 0 && (reinterpret_cast<uintptr_t>(h) & 15) == 0)
        ln_stats_kernel<true><<<(M + wpb - 1) / wpb, wpb * 32, 0, s>>>(h, ldh, M, width, eps, mu, rstd);
    else
        ln_stats_kernel<false><<<(M + wpb - 1) / wpb, wpb * 32, 0, s>>>(h, ldh, M, width, eps, mu, rstd);
    fc_count_launch();
    FC_LAUNCH_OK();
    return FC_OK;
}

static int gemm_plain(const FcLinear& l, const float* A1, int lda1, const float* A2, int lda2, const float* bias,
                      int bias_ld, int bias_group, const float* res, int ldres, int act, float* C, int ldc, int M,
                      int precision, cudaStream_t stream) {
    GemmArgs g = fc_gemm_args_zero();
    g.A1 = A1; g.lda1 = lda1; g.K1 = l.K1; g.A2 = A2; g.lda2 = lda2; g.K2 = l.K2;
    g.Wt = l.w; g.ldw = l.ldw; g.Whi = l.whi; g.Wlo = l.wlo; g.ldk = l.ldk; g.tc_fmt = l.tc_fmt;
    if (bias) { g.bias = bias; g.bias_ld = bias_ld; g.bias_group = bias_group; }
    else { g.bias = l.b; }
    g.res = res; g.ldres = ldres; g.act = act; g.C = C; g.ldc = ldc; g.M = M; g.N = l.N;
    g.precision = precision;
    return fc_launch_gemm(g, stream);
}

int fc_run_mlp_hidden(const FcMlp& m, const FcMlpIn& in, int M, float* bufA, float* bufB, int ldh, int precision,
                      cudaStream_t stream, float** last) {
    // reference models/nets.py:19-30: x = act(in(x)); even hidden i: res = x, x = act(layer(x));
    // odd i: x = act(res + layer(x))
    int rc = gemm_plain(m.in, in.A1, in.lda1, in.A2, in.lda2, in.bias, in.bias_ld, in.bias_group, nullptr, 0,
                        m.act, bufA, ldh, M, precision, stream);
    if (rc) return rc;
    float* cur = bufA; float* other = bufB;
    for (int i = 0; i < m.n_hidden; ++i) {
        if ((i & 1) == 0) {
            rc = gemm_plain(m.hidden[i], cur, ldh, nullptr, 0, nullptr, 0, 0, nullptr, 0, m.act, other, ldh, M,
                            precision, stream);
        } else {
            // residual source is `other` (the activation before the previous layer); write in place over it
            rc = gemm_plain(m.hidden[i], cur, ldh, nullptr, 0, nullptr, 0, 0, other, ldh, m.act, other, ldh, M,
                            precision, stream);
        }
        if (rc) return rc;
        float* t = cur; cur = other; other = t;
    }
    *last = cur;
    return FC_OK;
}

// ------------------------------------------------------------------------------------------ create
static FcAttn read_attn(FcCursor& c, const fc_flow& f) {
    FcAttn a;
    a.csum = c.ptr(c.next(), f.inner);
    a.qbias = c.ptr(c.next(), f.inner);
    a.q = c.linear(f.attn_in, 0, f.inner, /*has_bias=*/false);
    a.kv = c.linear(f.E, 0, 2 * f.inner, /*has_bias=*/false);
    if (!a.csum || !a.qbias) c.ok = false;
    return a;
}

extern "C" int fc_flow_create(const int32_t* header, int n_header, const int64_t* table, int n_table,
                              const float* arena, int64_t arena_floats, fc_flow** out) {
    FC_REQUIRE(header && table && arena && out && n_header >= 19);
    if (header[0] != FC_FLOW_MAGIC || (header[1] != FC_ARENA_VERSION && header[1] != FC_ARENA_VERSION_F16)) return FC_ERR_MODEL;
    if (reinterpret_cast<uintptr_t>(arena) & 15) return FC_ERR_MODEL;
    fc_flow* f = new (std::nothrow) fc_flow();
    if (!f) return FC_ERR_MODEL;
    f->L = header[2]; f->D = header[3]; f->d_in = header[4]; f->half = header[5]; f->extra = header[6];
    f->is_global = header[7]; f->E = header[8]; f->inner = header[9]; f->attn_in = header[10];
    f->hid = header[11]; f->n_hid = header[12]; f->pre_hid = header[13]; f->n_pre_hid = header[14];
    f->aug_hid = header[15]; f->n_aug_hid = header[16]; f->augpre_hid = header[17]; f->n_augpre_hid = header[18];
    f->arena = arena; f->arena_floats = arena_floats;
    if (n_header >= 29) {
        f->cpl_kind = header[19]; f->num_bins = header[20]; f->act = header[21]; f->has_aug = header[22];
        f->cif_dim = header[23]; f->cif_hid = header[24]; f->n_cif_hid = header[25]; f->affcif_hid = header[26];
        f->n_affcif_hid = header[27];
        memcpy(&f->cif_clamp, &header[28], sizeof(float));
    }
    const int n2 = f->D - f->half, S = f->cif_dim ? f->cif_dim - f->D : 0;
    bool dims_ok = f->L >= 1 && f->D >= 2 && f->D <= 512 && f->d_in >= 1 && f->d_in <= f->D && f->half == f->D / 2 &&
                   f->inner == 64 && f->attn_in <= 512 && f->hid <= 512 && f->pre_hid <= 512 && f->aug_hid <= 512 &&
                   f->augpre_hid <= 512 && (f->D % 2) == 0 && ((f->D - f->d_in) % 2) == 0 &&
                   (f->has_aug ? (f->d_in < f->D && f->hid == f->aug_hid) : f->d_in == f->D) &&
                   f->cpl_kind >= FC_CPL_AFFINE && f->cpl_kind <= FC_CPL_EXPO && (f->act == FC_ACT_GELU || f->act == FC_ACT_RELU) &&
                   (f->cpl_kind != FC_CPL_SPLINE || (f->num_bins >= 1 && f->num_bins <= 16)) &&
                   (f->cpl_kind != FC_CPL_EXPO || n2 <= fc_expm_max_n()) &&
                   (f->cif_dim == 0 || (S > 0 && f->cif_dim <= 512 && f->cif_hid <= 512 && f->affcif_hid <= 512 &&
                                        f->cif_hid > 0 && f->affcif_hid > 0 && !f->extra && !f->is_global));
    if (!dims_ok) { delete f; return FC_ERR_UNSUPPORTED; }
    f->layers = new (std::nothrow) FcFlowLayer[f->L];
    if (!f->layers) { delete f; return FC_ERR_MODEL; }

    FcCursor c{table, n_table, 0, arena, arena_floats, true};
    c.tc_fmt = header[1] == FC_ARENA_VERSION_F16 ? 1 : 0;
    const int64_t cbits = c.next();
    memcpy(&f->ldj_const, &cbits, sizeof(double));
    const int attn_k2 = f->is_global ? 0 : f->inner;
    if (f->has_aug) {
        if (!f->is_global) {
            f->augpre = c.mlp(f->d_in, 0, f->augpre_hid, f->n_augpre_hid, f->attn_in);
            f->augpre.act = f->act;
            f->augattn = read_attn(c, *f);
        }
        f->aug = c.mlp(f->d_in, attn_k2, f->aug_hid, f->n_aug_hid, 2 * (f->D - f->d_in));
        f->aug.act = f->act;
    }
    const int cpl_out = f->cpl_kind == FC_CPL_AFFINE ? 2 * n2 : f->cpl_kind == FC_CPL_SPLINE ? (3 * f->num_bins + 1) * f->half
                                                                                              : n2 * n2 + n2;
    f->has_cb = f->extra || f->is_global;
    if (f->has_cb) f->cb = c.linear(f->extra + (f->is_global ? f->E : 0), 0, (f->L + 1) * f->hid);
    for (int l = 0; l < f->L; ++l) {
        FcFlowLayer& y = f->layers[l];
        if (f->cif_dim) {
            y.cifnet = c.mlp(f->D, 0, f->cif_hid, f->n_cif_hid, 2 * S);                 // GELU (cif_block.py:54)
            y.affcif = c.mlp(S, 0, f->affcif_hid, f->n_affcif_hid, 2 * f->D);           // GELU (cif_block.py:64)
            y.cif_sc = c.ptr(c.next(), f->cif_dim);
            y.cif_bi = c.ptr(c.next(), f->cif_dim);
            if (!y.cif_sc || !y.cif_bi) c.ok = false;
        }
        if (!f->is_global) {
            y.pre = c.mlp(f->half, 0, f->pre_hid, f->n_pre_hid, f->attn_in);
            y.pre.act = f->cif_dim ? FC_ACT_GELU : f->act;                              // the CIF block's own MLP is GELU (:62)
            y.attn = read_attn(c, *f);
        }
        if (f->cpl_kind == FC_CPL_EXPO) { y.expo_sq = c.ptr(c.next(), 4); if (!y.expo_sq) c.ok = false; }
        y.cpl = c.mlp(f->half, attn_k2, f->hid, f->n_hid, cpl_out);
        y.cpl.act = f->act;
        y.has_lu = (l != f->L - 1);
        if (y.has_lu) {
            y.lu = c.linear(f->D, 0, f->D);
            y.lu_diag = c.ptr(c.next(), f->D);
            if (!y.lu_diag) c.ok = false;
        }
    }
    if (!c.ok || c.pos != n_table) { fc_flow_destroy(f); return FC_ERR_MODEL; }
    *out = f;
    return FC_OK;
}

extern "C" void fc_flow_destroy(fc_flow* f) {
    if (!f) return;
    delete[] f->layers;
    delete f;
}

// ------------------------------------------------------------------------------------------ workspace
namespace {
struct FlowWs {
    float *lat0, *lat1, *hA, *hB, *q, *o, *mu, *rstd, *cpart, *apart, *kv, *kvs, *cb, *cbA, *pbuf;
    int ldx, ldh, n_cpart, n_apart, cb_ld, cbA_ld, ldp, p_rows, ldp_cif;
    int64_t total_bytes;
};

FlowWs carve_flow_ws(const fc_flow* f, int B, int N, int Nc, void* base) {
    FlowWs w{};
    const int64_t M = (int64_t)B * N;
    w.ldx = fc_round_up(f->cif_dim > f->D ? f->cif_dim : f->D, 4);
    w.ldh = 512;
    // one partial log-det slab per N-tile of whichever GEMM kernel runs (FFMA: 128-wide tiles, tcgen05: <= 96-wide)
    auto slabs = [](int n) { const int a = fc_gemm_n_tiles(n), b = fc_tc_n_tiles(n); return a > b ? a : b; };
    w.n_cpart = slabs(2 * (f->D - f->half));
    if (f->cif_dim) { const int a = slabs(2 * f->D); if (a > w.n_cpart) w.n_cpart = a; }
    w.n_apart = f->has_aug ? slabs(2 * (f->D - f->d_in)) : 1;
    // raw conditioner outputs of the spline / exponential couplings and of the CIF block's ConditionalNormal net; the
    // exponential coupling's (n^2 + n floats per point) are produced and consumed in chunks of whole clouds
    w.ldp = 0; w.p_rows = 0; w.ldp_cif = 0;
    {
        const int n2 = f->D - f->half;
        if (f->cpl_kind == FC_CPL_SPLINE) { w.ldp = fc_round_up((3 * f->num_bins + 1) * f->half, 4); w.p_rows = (int)M; }
        if (f->cpl_kind == FC_CPL_EXPO) {
            w.ldp = fc_round_up(n2 * n2 + n2, 4);
            const long long clouds = (1ll << 28) / ((long long)w.ldp * N);     // ~1 GB of parameters at a time
            w.p_rows = (int)((clouds < 1 ? 1 : (clouds > B ? B : clouds)) * N);
        }
        if (f->cif_dim) w.ldp_cif = fc_round_up(2 * (f->cif_dim - f->D), 4);
    }
    w.cb_ld = (f->L + 1) * f->hid;
    w.cbA_ld = fc_round_up(f->extra + (f->is_global ? f->E : 0), 4);
    if (w.cbA_ld == 0) w.cbA_ld = 4;
    int64_t off = 0;
    auto take = [&](int64_t floats) { float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
                                      off += fc_round_up_ll(floats * 4, 256); return p; };
    w.lat0 = take(M * w.ldx); w.lat1 = take(M * w.ldx);
    w.hA = take(M * w.ldh); w.hB = take(M * w.ldh);
    w.q = take(M * 64); w.o = take(M * 64);
    w.mu = take(M); w.rstd = take(M);
    w.cpart = take(M * w.n_cpart); w.apart = take(M * w.n_apart);
    // attention configs only (the global embedder's flow never runs an attention kernel)
    w.kv = take(f->is_global ? 0 : (int64_t)B * Nc * 128);
    w.kvs = take(f->is_global ? 0 : fc_attention_tc_scratch_floats(B, Nc));   // TF32 hi/lo copies of k, v^T for the tcgen05 attention
    w.cb = take((int64_t)B * w.cb_ld);
    w.cbA = take((int64_t)B * w.cbA_ld);
    {
        const int64_t a = (int64_t)w.ldp * w.p_rows, b = (int64_t)w.ldp_cif * M;
        w.pbuf = take(a > b ? a : b);
    }
    w.total_bytes = off;
    return w;
}
}  // namespace

extern "C" int64_t fc_flow_workspace_bytes(const fc_flow* f, int B, int N, int Nc) {
    if (!f || B <= 0 || N <= 0 || Nc <= 0) return FC_ERR_INVALID_ARG;
    return carve_flow_ws(f, B, N, Nc, nullptr).total_bytes;
}

// ------------------------------------------------------------------------------------------ forward
static int run_attention_block(const fc_flow* f, const FcMlp& pre, const FcAttn& at, const float* lat, int ldx, int K_in,
                               const float* context, int B, int N, int Nc, FlowWs& w, int precision, cudaStream_t s) {
    const int M = B * N;
    FcMlpIn in{lat, ldx, nullptr, 0, nullptr, 0, 0};
    (void)K_in;
    float* last = nullptr;
    int rc = fc_run_mlp_hidden(pre, in, M, w.hA, w.hB, w.ldh, precision, s, &last);
    if (rc) return rc;
    float* h4 = (last == w.hA) ? w.hB : w.hA;
    rc = gemm_plain(pre.out, last, w.ldh, nullptr, 0, nullptr, 0, 0, nullptr, 0, FC_ACT_NONE, h4, w.ldh, M, precision, s);
    if (rc) return rc;
    rc = fc_launch_ln_stats(h4, w.ldh, M, f->attn_in, 1e-5f, w.mu, w.rstd, s);
    if (rc) return rc;
    {
        GemmArgs g = fc_gemm_args_zero();
        g.A1 = h4; g.lda1 = w.ldh; g.K1 = f->attn_in; g.Wt = at.q.w; g.ldw = at.q.ldw; g.bias = at.qbias; g.Whi = at.q.whi; g.Wlo = at.q.wlo; g.ldk = at.q.ldk; g.tc_fmt = at.q.tc_fmt;
        g.C = w.q; g.ldc = 64; g.M = M; g.N = f->inner; g.epi = FC_EPI_LNQ; g.row_mu = w.mu; g.row_rstd = w.rstd;
        g.csum = at.csum; g.precision = precision;
        rc = fc_launch_gemm(g, s);
        if (rc) return rc;
    }
    // reference models/perceiver.py:104: scale = inner_dim ** -0.5
    const float scale = 1.0f / sqrtf((float)f->inner);
    static int attn_kind = -1;   // FC_ATTN=mma: warp-level MMA kernel; FC_ATTN=split: tcgen05 with a separate split pass (A/B runs)
    if (attn_kind < 0) { const char* e = getenv("FC_ATTN"); attn_kind = !e ? 0 : (e[0] == 'm' ? 1 : (e[0] == 's' ? 2 : 0)); }
    if (precision == 1 && attn_kind == 0 && at.kv.N == 128 && at.kv.whi) {
        // to_kv writes the TF32 hi/lo copies of k and v^T the tcgen05 attention consumes, straight from its epilogue
        GemmArgs g = fc_gemm_args_zero();
        g.A1 = context; g.lda1 = f->E; g.K1 = at.kv.K1; g.Wt = at.kv.w; g.ldw = at.kv.ldw; g.bias = at.kv.b;
        g.Whi = at.kv.whi; g.Wlo = at.kv.wlo; g.ldk = at.kv.ldk; g.tc_fmt = at.kv.tc_fmt; g.M = B * Nc; g.N = 128; g.precision = 1;
        g.epi = FC_EPI_KVSPLIT; g.ldc = 64; g.kv_nc = Nc;
        g.kv_f16 = at.kv.tc_fmt ? 1 : 0;      // 3xFP16 models run the attention on fp16 operands as well
        fc_attention_tc_scratch_layout(B, Nc, w.kvs, &g.C, &g.kv_klo, &g.kv_vthi, &g.kv_vtlo, &g.kv_ncp, g.kv_f16);
        rc = fc_launch_gemm(g, s);
        if (rc) return rc;
        return fc_launch_cross_attention_tc(w.q, 64, nullptr, 0, w.o, 64, B, N, Nc, f->inner, scale, w.kvs, 1, g.kv_f16, s);
    }
    rc = gemm_plain(at.kv, context, f->E, nullptr, 0, nullptr, 0, 0, nullptr, 0, FC_ACT_NONE, w.kv, 128, B * Nc,
                    precision, s);
    if (rc) return rc;
    if (precision == 1) {
        if (attn_kind == 1) return fc_launch_cross_attention_mma(w.q, 64, w.kv, 128, w.o, 64, B, N, Nc, f->inner, scale, s);
        return fc_launch_cross_attention_tc(w.q, 64, w.kv, 128, w.o, 64, B, N, Nc, f->inner, scale, w.kvs, 0, at.kv.tc_fmt ? 1 : 0, s);
    }
    return fc_launch_cross_attention(w.q, 64, w.kv, 128, w.o, 64, B, N, Nc, f->inner, scale, s);
}

// CIF block up to (not including) its attention-conditioned coupling, reference models/cif_block.py:71-93.  In the
// reference's order: Augment -> Reverse -> AffineCoupling(split = cif_dim - D) -> ActNorm -> Reverse -> Slice.  The two
// Reverse permutations cancel once the affine coupling's weights and the ActNorm vectors are re-indexed at pack time
// (packing.py: _pack_cif), so on the device the block is: z2 ~ N(mean(x), sigma(x)) into columns [D, cif_dim);
// x = x * s(z2) + t(z2); every column c: a*sc[c] + bi[c]; log N(z2; mean(x), sigma(x)) with the SAME net (cif_block.py:58).
static int run_cif_block(const fc_flow* f, const FcFlowLayer& y, float* lat, FlowWs& w, int M, const float* eps_l,
                         int precision, cudaStream_t s) {
    const int D = f->D, S = f->cif_dim - f->D;
    int rc;
    for (int pass = 0; pass < 2; ++pass) {
        FcMlpIn in{lat, w.ldx, nullptr, 0, nullptr, 0, 0};
        float* last = nullptr;
        rc = fc_run_mlp_hidden(y.cifnet, in, M, w.hA, w.hB, w.ldh, precision, s, &last);
        if (rc) return rc;
        rc = gemm_plain(y.cifnet.out, last, w.ldh, nullptr, 0, nullptr, 0, 0, nullptr, 0, FC_ACT_NONE, w.pbuf, w.ldp_cif, M,
                        precision, s);
        if (rc) return rc;
        rc = fc_launch_cond_normal(w.pbuf, w.ldp_cif, lat, w.ldx, D, S, eps_l, M, f->cif_clamp, w.cpart, pass, s);
        if (rc || pass == 1) return rc;
        FcMlpIn in2{lat + D, w.ldx, nullptr, 0, nullptr, 0, 0};
        rc = fc_run_mlp_hidden(y.affcif, in2, M, w.hA, w.hB, w.ldh, precision, s, &last);
        if (rc) return rc;
        GemmArgs g = fc_gemm_args_zero();
        const FcLinear& o = y.affcif.out;
        g.A1 = last; g.lda1 = w.ldh; g.K1 = o.K1; g.Wt = o.w; g.ldw = o.ldw; g.bias = o.b; g.Whi = o.whi; g.Wlo = o.wlo; g.ldk = o.ldk;
        g.tc_fmt = o.tc_fmt; g.M = M; g.N = o.N; g.epi = FC_EPI_COUPLING; g.x = lat; g.ldx = w.ldx; g.col0 = 0; g.part = w.cpart;
        g.precision = precision;
        rc = fc_launch_gemm(g, s);
        if (rc) return rc;
        rc = fc_launch_col_affine(lat, w.ldx, f->cif_dim, M, y.cif_sc, y.cif_bi, s);
        if (rc) return rc;
    }
    return FC_OK;
}

static int flow_forward(const fc_flow* f, const float* x, const float* context, const float* extra, const float* eps,
                        const float* eps_cif, float* log_prob_out, float* z_out, int B, int N, int Nc, void* workspace,
                        int64_t workspace_bytes, int precision, fc_stream_t stream_) {
    FC_REQUIRE(f && x && context && log_prob_out && B > 0 && N > 0 && Nc > 0);
    FC_REQUIRE(eps || !f->has_aug);
    FC_REQUIRE(eps_cif || !f->cif_dim);
    FC_REQUIRE((f->extra != 0) == (extra != nullptr));
    FC_REQUIRE((int64_t)B * N < (1ll << 31) && (int64_t)B * Nc < (1ll << 31));
    cudaStream_t s = (cudaStream_t)stream_;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255)) return FC_ERR_WORKSPACE;
    FlowWs w = carve_flow_ws(f, B, N, Nc, workspace);
    if (w.total_bytes > workspace_bytes) return FC_ERR_WORKSPACE;
    const int M = B * N;
    int rc;

    // x -> latent columns [0, d_in)
    {
        const long long tot = (long long)M * f->d_in;
        copy_cols_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(x, f->d_in, w.lat0, w.ldx, M, f->d_in);
        fc_count_launch();
        FC_LAUNCH_OK();
    }
    // partial log-det slabs: the FFMA and tcgen05 GEMMs tile N differently (128 vs <=256 wide), so slabs a
    // path does not write must read as zero in finalize
    FC_CUDA_OK(cudaMemsetAsync(w.cpart, 0, (size_t)M * w.n_cpart * sizeof(float), s));
    FC_CUDA_OK(cudaMemsetAsync(w.apart, 0, (size_t)M * w.n_apart * sizeof(float), s));
    // per-cloud bias of every conditioner's first layer (extra context and/or global embedding)
    if (f->has_cb) {
        const int tot = B * w.cbA_ld;
        build_cb_input_kernel<<<(tot + 255) / 256, 256, 0, s>>>(extra, context, f->E, f->extra, f->is_global, B, w.cbA_ld, w.cbA);
        fc_count_launch();
        FC_LAUNCH_OK();
        rc = gemm_plain(f->cb, w.cbA, w.cbA_ld, nullptr, 0, nullptr, 0, 0, nullptr, 0, FC_ACT_NONE, w.cb, w.cb_ld, B,
                        0 /* always exact fp32: tiny */, s);
        if (rc) return rc;
    }
    auto cb_ptr = [&](int slot) -> const float* { return f->has_cb ? w.cb + (size_t)slot * f->hid : nullptr; };

    // ---- transforms.0: augment (reference models/augmenter.py:15-19, :49-63); latent_dim == input_dim: IdentityTransform
    if (f->has_aug && !f->is_global) {
        rc = run_attention_block(f, f->augpre, f->augattn, w.lat0, w.ldx, f->d_in, context, B, N, Nc, w, precision, s);
        if (rc) return rc;
    }
    if (f->has_aug) {
        FcMlpIn in{w.lat0, w.ldx, f->is_global ? nullptr : w.o, 64, cb_ptr(0), w.cb_ld, N};
        if (!f->has_cb) { in.bias = nullptr; in.bias_ld = 0; in.bias_group = 0; }
        float* last = nullptr;
        rc = fc_run_mlp_hidden(f->aug, in, M, w.hA, w.hB, w.ldh, precision, s, &last);
        if (rc) return rc;
        GemmArgs g = fc_gemm_args_zero();
        g.A1 = last; g.lda1 = w.ldh; g.K1 = f->aug_hid; g.Wt = f->aug.out.w; g.ldw = f->aug.out.ldw; g.bias = f->aug.out.b; g.Whi = f->aug.out.whi; g.Wlo = f->aug.out.wlo; g.ldk = f->aug.out.ldk; g.tc_fmt = f->aug.out.tc_fmt;
        g.M = M; g.N = f->aug.out.N; g.epi = FC_EPI_AUGMENT; g.x = w.lat0; g.ldx = w.ldx; g.col0 = f->d_in;
        g.part = w.apart; g.eps = eps; g.ld_eps = f->D - f->d_in; g.precision = precision;
        rc = fc_launch_gemm(g, s);
        if (rc) return rc;
    }

    // ---- coupling layers
    float* lat = w.lat0; float* lat_next = w.lat1;
    for (int l = 0; l < f->L; ++l) {
        const FcFlowLayer& y = f->layers[l];
        if (f->cif_dim) {
            rc = run_cif_block(f, y, lat, w, M, eps_cif + (size_t)l * M * (f->cif_dim - f->D), precision, s);
            if (rc) return rc;
        }
        if (!f->is_global) {
            rc = run_attention_block(f, y.pre, y.attn, lat, w.ldx, f->half, context, B, N, Nc, w, precision, s);
            if (rc) return rc;
        }
        FcMlpIn in{lat, w.ldx, f->is_global ? nullptr : w.o, 64, cb_ptr(l + 1), w.cb_ld, N};
        if (!f->has_cb) { in.bias = nullptr; in.bias_ld = 0; in.bias_group = 0; }
        float* last = nullptr;
        rc = fc_run_mlp_hidden(y.cpl, in, M, w.hA, w.hB, w.ldh, precision, s, &last);
        if (rc) return rc;
        if (f->cpl_kind == FC_CPL_SPLINE) {
            // reference models/spline_coupling.py:187-210: the conditioner's raw output, then the spline on x2 in place
            rc = gemm_plain(y.cpl.out, last, w.ldh, nullptr, 0, nullptr, 0, 0, nullptr, 0, FC_ACT_NONE, w.pbuf, w.ldp, M, precision, s);
            if (rc) return rc;
            rc = fc_launch_rq_spline(w.pbuf, w.ldp, lat, w.ldx, f->half, f->D - f->half, f->num_bins, M, w.cpart, 0, s);
            if (rc) return rc;
        } else if (f->cpl_kind == FC_CPL_EXPO) {
            // reference models/exponential_coupling.py:44-58, n^2 + n conditioner outputs per point: whole clouds at a time
            for (int r0 = 0; r0 < M; r0 += w.p_rows) {
                const int rows = M - r0 < w.p_rows ? M - r0 : w.p_rows;
                rc = gemm_plain(y.cpl.out, last + (size_t)r0 * w.ldh, w.ldh, nullptr, 0, nullptr, 0, 0, nullptr, 0, FC_ACT_NONE,
                                w.pbuf, w.ldp, rows, precision, s);
                if (rc) return rc;
                rc = fc_launch_expm_action(w.pbuf, w.ldp, lat, w.ldx, f->half, f->D - f->half, y.expo_sq, w.cpart, r0, rows, 0, s);
                if (rc) return rc;
            }
        } else {
            GemmArgs g = fc_gemm_args_zero();
            g.A1 = last; g.lda1 = w.ldh; g.K1 = f->hid; g.Wt = y.cpl.out.w; g.ldw = y.cpl.out.ldw; g.bias = y.cpl.out.b; g.Whi = y.cpl.out.whi; g.Wlo = y.cpl.out.wlo; g.ldk = y.cpl.out.ldk; g.tc_fmt = y.cpl.out.tc_fmt;
            g.M = M; g.N = y.cpl.out.N; g.epi = FC_EPI_COUPLING; g.x = lat; g.ldx = w.ldx; g.col0 = f->half;
            g.part = w.cpart; g.precision = precision;
            rc = fc_launch_gemm(g, s);
            if (rc) return rc;
        }
        if (y.has_lu) {
            // z' = diag.z + (W' - diag) z + b': the diagonal (~1) is applied to the fp32 latent in the epilogue,
            // the GEMM only carries the off-diagonal mixing, so its rounding (and the tensor core's accumulate
            // truncation) acts on the small part instead of shrinking z itself 114 times in a row.
            GemmArgs g = fc_gemm_args_zero();
            g.A1 = lat; g.lda1 = w.ldx; g.K1 = y.lu.K1; g.Wt = y.lu.w; g.ldw = y.lu.ldw; g.Whi = y.lu.whi; g.Wlo = y.lu.wlo;
            g.ldk = y.lu.ldk; g.tc_fmt = y.lu.tc_fmt; g.bias = y.lu.b; g.res = lat; g.ldres = w.ldx; g.res_scale = y.lu_diag;
            g.C = lat_next; g.ldc = w.ldx; g.M = M; g.N = y.lu.N; g.precision = precision;
            rc = fc_launch_gemm(g, s);
            if (rc) return rc;
            float* t = lat; lat = lat_next; lat_next = t;
        }
    }
    {
        const int wpb = 8;
        finalize_kernel<<<(M + wpb - 1) / wpb, wpb * 32, 0, s>>>(lat, w.ldx, f->D, w.apart, w.n_apart, w.cpart, w.n_cpart,
                                                                 M, (float)f->ldj_const, log_prob_out);
        fc_count_launch();
        FC_LAUNCH_OK();
    }
    if (z_out) {   // the final latent, [B*N, D] contiguous
        const long long tot = (long long)M * f->D;
        copy_cols_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(lat, w.ldx, z_out, f->D, M, f->D);
        fc_count_launch();
        FC_LAUNCH_OK();
    }
    return FC_OK;
}

extern "C" int fc_flow_log_prob(const fc_flow* f, const float* x, const float* context, const float* extra,
                                const float* eps, float* log_prob_out, int B, int N, int Nc, void* workspace,
                                int64_t workspace_bytes, int precision, fc_stream_t stream_) {
    return flow_forward(f, x, context, extra, eps, nullptr, log_prob_out, nullptr, B, N, Nc, workspace, workspace_bytes, precision, stream_);
}

// Same pass for flows built with CIF blocks (latent_dim < cif_latent_dim): `eps_cif` [L, B, N, cif_latent_dim - latent_dim] holds
// the draw of every block's augmenter (reference models/cif_block.py:74), in transform-list order.
extern "C" int fc_flow_log_prob_cif(const fc_flow* f, const float* x, const float* context, const float* extra,
                                    const float* eps, const float* eps_cif, float* log_prob_out, int B, int N, int Nc,
                                    void* workspace, int64_t workspace_bytes, int precision, fc_stream_t stream_) {
    return flow_forward(f, x, context, extra, eps, eps_cif, log_prob_out, nullptr, B, N, Nc, workspace, workspace_bytes, precision, stream_);
}

extern "C" int fc_flow_cif_noise_dim(const fc_flow* f) { return f ? (f->cif_dim ? f->cif_dim - f->D : 0) : FC_ERR_INVALID_ARG; }

extern "C" int fc_flow_forward(const fc_flow* f, const float* x, const float* context, const float* extra, const float* eps,
                               float* log_prob_out, float* z_out, int B, int N, int Nc, void* workspace, int64_t workspace_bytes,
                               int precision, fc_stream_t stream_) {
    FC_REQUIRE(z_out != nullptr);
    return flow_forward(f, x, context, extra, eps, nullptr, log_prob_out, z_out, B, N, Nc, workspace, workspace_bytes, precision, stream_);
}

// ------------------------------------------------------------------------------------------ inverse / sampling pass
extern "C" int fc_flow_set_inverse(fc_flow* f, const int32_t* header, int n_header, const int64_t* table, int n_table,
                                   const float* arena, int64_t arena_floats) {
    FC_REQUIRE(f && header && table && arena && n_header >= 4);
    if (header[0] != FC_INV_MAGIC || (header[1] != FC_ARENA_VERSION && header[1] != FC_ARENA_VERSION_F16)) return FC_ERR_MODEL;
    if (header[2] != f->L || header[3] != f->D || (reinterpret_cast<uintptr_t>(arena) & 15)) return FC_ERR_MODEL;
    if ((n_header >= 5 ? header[4] : 0) != f->cif_dim) return FC_ERR_MODEL;
    FcCursor c{table, n_table, 0, arena, arena_floats, true};
    c.tc_fmt = header[1] == FC_ARENA_VERSION_F16 ? 1 : 0;
    for (int l = 0; l < f->L; ++l) {
        FcFlowLayer& y = f->layers[l];
        if (f->cif_dim) {
            y.cif_isc = c.ptr(c.next(), f->cif_dim);
            y.cif_ibi = c.ptr(c.next(), f->cif_dim);
            if (!y.cif_isc || !y.cif_ibi) c.ok = false;
        }
        if (!y.has_lu) continue;
        y.lu_inv = c.linear(f->D, 0, f->D);
        y.lu_inv_diag = c.ptr(c.next(), f->D);
        if (!y.lu_inv_diag) c.ok = false;
    }
    if (!c.ok || c.pos != n_table) return FC_ERR_MODEL;
    f->has_inverse = true;
    return FC_OK;
}

// `Flow.sample` after the base draw (reference models/transform.py:79-84): the transform list walked backwards with each
// `.inverse` -- LinearLU (models/permuters.py:171-177) and ActNorm (models/act_norm.py:45-46) as ONE folded GEMM
// z = Wp^-1 z' + shift (packing.fold_inverse_actnorm_lu), the affine coupling's inverse x2 = (y2 - t)/s
// (models/affine_coupling.py:48-62) in the epilogue of its conditioner's last GEMM, the conditioner itself exactly as in the
// forward pass (it reads the untouched half y1); the augmenter's inverse keeps the first input_dim columns
// (models/augmenter.py:20-21,65-67).  z: [B, P, D] base draw, x_out: [B, P, d_in].
static int flow_sample(const fc_flow* f, const float* z, const float* context, const float* extra, const float* eps_cif,
                       float* x_out, int B, int P, int Nc, void* workspace, int64_t workspace_bytes, int precision,
                       fc_stream_t stream_) {
    FC_REQUIRE(f && z && context && x_out && B > 0 && P > 0 && Nc > 0);
    FC_REQUIRE((f->extra != 0) == (extra != nullptr));
    FC_REQUIRE(eps_cif || !f->cif_dim);
    FC_REQUIRE((int64_t)B * P < (1ll << 31) && (int64_t)B * Nc < (1ll << 31));
    if (!f->has_inverse) return FC_ERR_MODEL;
    cudaStream_t s = (cudaStream_t)stream_;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255)) return FC_ERR_WORKSPACE;
    FlowWs w = carve_flow_ws(f, B, P, Nc, workspace);
    if (w.total_bytes > workspace_bytes) return FC_ERR_WORKSPACE;
    const int M = B * P;
    const int n2 = f->D - f->half, S = f->cif_dim ? f->cif_dim - f->D : 0;
    int rc;
    {
        const long long tot = (long long)M * f->D;
        copy_cols_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(z, f->D, w.lat0, w.ldx, M, f->D);
        fc_count_launch();
        FC_LAUNCH_OK();
    }
    if (f->has_cb) {
        const int tot = B * w.cbA_ld;
        build_cb_input_kernel<<<(tot + 255) / 256, 256, 0, s>>>(extra, context, f->E, f->extra, f->is_global, B, w.cbA_ld, w.cbA);
        fc_count_launch();
        FC_LAUNCH_OK();
        rc = gemm_plain(f->cb, w.cbA, w.cbA_ld, nullptr, 0, nullptr, 0, 0, nullptr, 0, FC_ACT_NONE, w.cb, w.cb_ld, B, 0, s);
        if (rc) return rc;
    }
    auto cb_ptr = [&](int slot) -> const float* { return f->has_cb ? w.cb + (size_t)slot * f->hid : nullptr; };
    auto coupling_inv_gemm = [&](const FcLinear& o, const float* last, float* lat) {
        GemmArgs g = fc_gemm_args_zero();
        g.A1 = last; g.lda1 = w.ldh; g.K1 = o.K1; g.Wt = o.w; g.ldw = o.ldw; g.bias = o.b; g.Whi = o.whi; g.Wlo = o.wlo;
        g.ldk = o.ldk; g.tc_fmt = o.tc_fmt; g.M = M; g.N = o.N; g.epi = FC_EPI_COUPLING_INV; g.x = lat; g.ldx = w.ldx;
        g.precision = precision;
        return g;
    };
    float* lat = w.lat0; float* lat_next = w.lat1;
    for (int l = f->L - 1; l >= 0; --l) {
        const FcFlowLayer& y = f->layers[l];
        if (y.has_lu) {
            GemmArgs g = fc_gemm_args_zero();
            g.A1 = lat; g.lda1 = w.ldx; g.K1 = y.lu_inv.K1; g.Wt = y.lu_inv.w; g.ldw = y.lu_inv.ldw; g.Whi = y.lu_inv.whi; g.Wlo = y.lu_inv.wlo;
            g.ldk = y.lu_inv.ldk; g.tc_fmt = y.lu_inv.tc_fmt; g.bias = y.lu_inv.b; g.res = lat; g.ldres = w.ldx; g.res_scale = y.lu_inv_diag;
            g.C = lat_next; g.ldc = w.ldx; g.M = M; g.N = y.lu_inv.N; g.precision = precision;
            rc = fc_launch_gemm(g, s);
            if (rc) return rc;
            float* t = lat; lat = lat_next; lat_next = t;
        }
        // ---- the coupling's inverse: the conditioner reads the untouched half y1 exactly as in the forward pass
        if (!f->is_global) {
            rc = run_attention_block(f, y.pre, y.attn, lat, w.ldx, f->half, context, B, P, Nc, w, precision, s);
            if (rc) return rc;
        }
        FcMlpIn in{lat, w.ldx, f->is_global ? nullptr : w.o, 64, cb_ptr(l + 1), w.cb_ld, P};
        if (!f->has_cb) { in.bias = nullptr; in.bias_ld = 0; in.bias_group = 0; }
        float* last = nullptr;
        rc = fc_run_mlp_hidden(y.cpl, in, M, w.hA, w.hB, w.ldh, precision, s, &last);
        if (rc) return rc;
        if (f->cpl_kind == FC_CPL_SPLINE) {            // reference models/spline_coupling.py:212-227
            rc = gemm_plain(y.cpl.out, last, w.ldh, nullptr, 0, nullptr, 0, 0, nullptr, 0, FC_ACT_NONE, w.pbuf, w.ldp, M, precision, s);
            if (rc) return rc;
            rc = fc_launch_rq_spline(w.pbuf, w.ldp, lat, w.ldx, f->half, n2, f->num_bins, M, nullptr, 1, s);
            if (rc) return rc;
        } else if (f->cpl_kind == FC_CPL_EXPO) {       // reference models/exponential_coupling.py:60-77
            for (int r0 = 0; r0 < M; r0 += w.p_rows) {
                const int rows = M - r0 < w.p_rows ? M - r0 : w.p_rows;
                rc = gemm_plain(y.cpl.out, last + (size_t)r0 * w.ldh, w.ldh, nullptr, 0, nullptr, 0, 0, nullptr, 0, FC_ACT_NONE,
                                w.pbuf, w.ldp, rows, precision, s);
                if (rc) return rc;
                rc = fc_launch_expm_action(w.pbuf, w.ldp, lat, w.ldx, f->half, n2, y.expo_sq, nullptr, r0, rows, 1, s);
                if (rc) return rc;
            }
        } else {                                       // reference models/affine_coupling.py:48-62
            GemmArgs g = coupling_inv_gemm(y.cpl.out, last, lat);
            g.col0 = f->half;
            rc = fc_launch_gemm(g, s);
            if (rc) return rc;
        }
        if (f->cif_dim) {
            // CIFblock.inverse after its flow (reference models/cif_block.py:102-111), Reverse folded as in the forward pass:
            // the sliced-off columns are SAMPLED from the ConditionalNormal of what was kept (models/slice.py:46-58), then the
            // inverse ActNorm on all cif_dim columns and x = (x - t(z2)) / s(z2); Augment.inverse drops the extra columns.
            FcMlpIn inc{lat, w.ldx, nullptr, 0, nullptr, 0, 0};
            rc = fc_run_mlp_hidden(y.cifnet, inc, M, w.hA, w.hB, w.ldh, precision, s, &last);
            if (rc) return rc;
            rc = gemm_plain(y.cifnet.out, last, w.ldh, nullptr, 0, nullptr, 0, 0, nullptr, 0, FC_ACT_NONE, w.pbuf, w.ldp_cif, M,
                            precision, s);
            if (rc) return rc;
            rc = fc_launch_cond_normal(w.pbuf, w.ldp_cif, lat, w.ldx, f->D, S, eps_cif + (size_t)l * M * S, M, f->cif_clamp,
                                       w.cpart /* log-density not needed: scratch */, 0, s);
            if (rc) return rc;
            rc = fc_launch_col_affine(lat, w.ldx, f->cif_dim, M, y.cif_isc, y.cif_ibi, s);
            if (rc) return rc;
            FcMlpIn in2{lat + f->D, w.ldx, nullptr, 0, nullptr, 0, 0};
            rc = fc_run_mlp_hidden(y.affcif, in2, M, w.hA, w.hB, w.ldh, precision, s, &last);
            if (rc) return rc;
            GemmArgs g = coupling_inv_gemm(y.affcif.out, last, lat);
            g.col0 = 0;
            rc = fc_launch_gemm(g, s);
            if (rc) return rc;
        }
    }
    const long long tot = (long long)M * f->d_in;
    copy_cols_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(lat, w.ldx, x_out, f->d_in, M, f->d_in);
    fc_count_launch();
    FC_LAUNCH_OK();
    return FC_OK;
}

extern "C" int fc_flow_sample(const fc_flow* f, const float* z, const float* context, const float* extra, float* x_out, int B,
                              int P, int Nc, void* workspace, int64_t workspace_bytes, int precision, fc_stream_t stream_) {
    return flow_sample(f, z, context, extra, nullptr, x_out, B, P, Nc, workspace, workspace_bytes, precision, stream_);
}

// Flows with CIF blocks: eps_cif [L, B, P, fc_flow_cif_noise_dim(f)] = the N(0,1) draws of every block's `Slice.inverse`
// (reference models/slice.py:46-58), indexed by the block's position in the FORWARD list.
extern "C" int fc_flow_sample_cif(const fc_flow* f, const float* z, const float* context, const float* extra,
                                  const float* eps_cif, float* x_out, int B, int P, int Nc, void* workspace,
                                  int64_t workspace_bytes, int precision, fc_stream_t stream_) {
    return flow_sample(f, z, context, extra, eps_cif, x_out, B, P, Nc, workspace, workspace_bytes, precision, stream_);
}
