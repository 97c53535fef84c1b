// C-ABI glue: error reporting, the whole-path entry points (`inner_loop`), consumers of log_prob.
#include "model.cuh"
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

int fc_launch_gemm_tc(const GemmArgs& a, cudaStream_t stream);  // gemm_tc.cu
bool fc_gemm_tc_supported(const GemmArgs& a);

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace

void fc_set_last_cuda_error(int code, const char* file, int line) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d", code, cudaGetErrorString((cudaError_t)code), file, line);
}
void fc_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

namespace {
struct ProfRec { int cls; double flops, bytes; cudaEvent_t a, b; long long tag; };
std::atomic<bool> g_prof_on{false};
std::vector<ProfRec> g_prof;      // guarded by g_prof_mu: launchers on several host threads may record concurrently
std::mutex g_prof_mu;
}  // namespace
bool fc_prof_enabled() { return g_prof_on.load(std::memory_order_relaxed); }
int fc_prof_open(int cls, double flops, double bytes, cudaStream_t s, long long tag) {
    ProfRec r{cls, flops, bytes, nullptr, nullptr, tag};
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, s);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
    return (int)g_prof.size() - 1;
}
void fc_prof_close(int id, cudaStream_t s) {     // each scope closes ITS record (not whatever was opened last)
    cudaEvent_t b = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        if (id >= 0 && id < (int)g_prof.size()) b = g_prof[(size_t)id].b;
    }
    if (b) cudaEventRecord(b, s);
}

extern "C" int fc_profile_begin(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on.store(true);
    return FC_OK;
}
extern "C" int fc_profile_end(double* ms, double* flops, double* bytes, int64_t* launches, int n_classes) {
    g_prof_on.store(false);
    FC_REQUIRE(ms && flops && bytes && launches && n_classes >= FC_N_CLASSES);
    for (int i = 0; i < n_classes; ++i) { ms[i] = 0; flops[i] = 0; bytes[i] = 0; launches[i] = 0; }
    FC_CUDA_OK(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof_mu);
    // FC_PROFILE_DUMP=<path>: one line per instrumented launch (class, shape tag, flops, ms) for per-shape breakdowns
    const char* dump = getenv("FC_PROFILE_DUMP");
    FILE* df = dump && dump[0] ? fopen(dump, "w") : nullptr;
    for (auto& r : g_prof) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
            if (df) fprintf(df, "%d %lld %.0f %.6f\n", r.cls, r.tag, r.flops, (double)t);
            ms[r.cls] += t; flops[r.cls] += r.flops; bytes[r.cls] += r.bytes; launches[r.cls] += 1;
        }
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    g_prof.clear();
    if (df) fclose(df);
    return FC_OK;
}

extern "C" int fc_version(void) { return 100; }
extern "C" const char* fc_last_error(void) { return g_err; }
extern "C" int64_t fc_launch_count(void) { return g_launches.load(); }

int fc_launch_gemm(const GemmArgs& a, cudaStream_t stream) {
    if (a.precision == 1 && fc_gemm_tc_supported(a)) return fc_launch_gemm_tc(a, stream);
    return fc_launch_gemm_ffma(a, stream);
}

extern "C" int fc_gemm(const float* A, int lda, const float* Wt, int ldw, const float* bias, float* C, int ldc, int M,
                       int N, int K, int act, int precision, fc_stream_t stream) {
    GemmArgs g = fc_gemm_args_zero();
    g.A1 = A; g.lda1 = lda; g.K1 = K; g.Wt = Wt; g.ldw = ldw; g.bias = bias; g.act = act; g.C = C; g.ldc = ldc;
    g.M = M; g.N = N; g.precision = precision;
    FC_REQUIRE(act >= 0 && act <= 3 && (precision == 0 || precision == 1));
    if (precision == 1 && !fc_gemm_tc_supported(g)) return FC_ERR_UNSUPPORTED;
    return fc_launch_gemm(g, (cudaStream_t)stream);
}

extern "C" int fc_gemm_tf32x3(const float* A, int lda, const float* Whi, const float* Wlo, int ldk, const float* bias,
                              float* C, int ldc, int M, int N, int K, int act, fc_stream_t stream) {
    GemmArgs g = fc_gemm_args_zero();
    g.A1 = A; g.lda1 = lda; g.K1 = K; g.Whi = Whi; g.Wlo = Wlo; g.ldk = ldk; g.bias = bias; g.act = act; g.C = C;
    g.ldc = ldc; g.M = M; g.N = N; g.precision = 1;
    FC_REQUIRE(act >= 0 && act <= 3 && A && Whi && Wlo && C);
    if (!fc_gemm_tc_supported(g)) return FC_ERR_UNSUPPORTED;
    return fc_launch_gemm_tc(g, (cudaStream_t)stream);
}

extern "C" int fc_gemm_f16x3(const float* A, int lda, const void* Whi16, const void* Wlo16, int ldk, const float* bias,
                             float* C, int ldc, int M, int N, int K, int act, fc_stream_t stream) {
    GemmArgs g = fc_gemm_args_zero();
    g.A1 = A; g.lda1 = lda; g.K1 = K; g.Whi = static_cast<const float*>(Whi16); g.Wlo = static_cast<const float*>(Wlo16);
    g.ldk = ldk; g.tc_fmt = 1; g.bias = bias; g.act = act; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.precision = 1;
    FC_REQUIRE(act >= 0 && act <= 3 && A && Whi16 && Wlo16 && C);
    if (!fc_gemm_tc_supported(g)) return FC_ERR_UNSUPPORTED;
    return fc_launch_gemm_tc(g, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------ stats
namespace {
// loss = -mean(log_prob), bpd = loss*log2(e)/input_dim  (reference model_initialization.py:225-227)
__global__ void stats_kernel(const float* __restrict__ lp, long long n, int input_dim, float* __restrict__ stats) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += (double)lp[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
        const double loss = -t / (double)n;
        stats[0] = (float)loss;
        stats[1] = (float)(loss * 1.4426950408889634 / (double)input_dim);
    }
}

// One CTA per cloud.  reference test_flow.py:241-275.
__global__ void change_score_kernel(const float* __restrict__ lp10, const float* __restrict__ lp00,
                                    float* __restrict__ out, int N, long long total, float multiple, int use_cut,
                                    float cut) {
    const int b = blockIdx.x;
    const float* a = lp10 + (size_t)b * N;
    const float* c = lp00 + (size_t)b * N;
    __shared__ float red[4][32];
    __shared__ double dred[2][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    // pass 1: min finite value of each WHOLE tensor (reference clamp_infs, test_flow.py:241-247, takes the
    // minimum over everything it is given, not per cloud); every CTA rescans -- B*N is tiny.
    float mn10 = INFINITY, mn00 = INFINITY;
    for (long long i = threadIdx.x; i < total; i += blockDim.x) {
        const float v = lp10[i]; if (!isinf(v)) mn10 = fminf(mn10, v);
        const float u = lp00[i]; if (!isinf(u)) mn00 = fminf(mn00, u);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn10 = fminf(mn10, __shfl_xor_sync(0xffffffffu, mn10, o));
        mn00 = fminf(mn00, __shfl_xor_sync(0xffffffffu, mn00, o));
    }
    if (lane == 0) { red[0][wid] = mn10; red[1][wid] = mn00; }
    __syncthreads();
    mn10 = INFINITY; mn00 = INFINITY;
    for (int i = 0; i < nw; ++i) { mn10 = fminf(mn10, red[0][i]); mn00 = fminf(mn00, red[1][i]); }
    __syncthreads();
    // pass 2: mean / unbiased std of clamped lp00; min / max of clamped lp10
    double s = 0.0;
    float mx = -INFINITY, mn = INFINITY;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float u = c[i]; if (isinf(u)) u = mn00;
        s += (double)u;
        float v = a[i]; if (isinf(v)) v = mn10;
        mx = fmaxf(mx, v); mn = fminf(mn, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if (lane == 0) { dred[0][wid] = s; red[2][wid] = mx; red[3][wid] = mn; }
    __syncthreads();
    s = 0.0; mx = -INFINITY; mn = INFINITY;
    for (int i = 0; i < nw; ++i) { s += dred[0][i]; mx = fmaxf(mx, red[2][i]); mn = fminf(mn, red[3][i]); }
    const double mean = s / (double)N;
    __syncthreads();
    double q = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float u = c[i]; if (isinf(u)) u = mn00;
        const double d = (double)u - mean; q += d * d;
    }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) dred[1][wid] = q;
    __syncthreads();
    q = 0.0;
    for (int i = 0; i < nw; ++i) q += dred[1][i];
    const float stdv = (float)sqrt(q / (double)(N - 1));
    const float thr = use_cut ? cut : ((float)mean - multiple * stdv);
    const float range = mx - mn;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float v = a[i]; if (isinf(v)) v = mn10;
        const float ch = 1.0f - (v - mn) / range;
        out[(size_t)b * N + i] = (v < thr) ? ch : 0.0f;
    }
}

// Philox-4x32-10 counter RNG + Box-Muller
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__global__ void fill_normal_kernel(float* __restrict__ out, long long n, uint64_t seed, uint64_t offset) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long base = t * 4;
    if (base >= n) return;
    const uint64_t ctr = offset + (uint64_t)t;
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    float u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = ((float)c[i] + 0.5f) * (1.0f / 4294967296.0f);
    float z[4];
    const float r0 = sqrtf(-2.0f * logf(u[0])), r1 = sqrtf(-2.0f * logf(u[2]));
    sincosf(6.283185307179586f * u[1], &z[1], &z[0]);
    sincosf(6.283185307179586f * u[3], &z[3], &z[2]);
    z[0] *= r0; z[1] *= r0; z[2] *= r1; z[3] *= r1;
#pragma unroll
    for (int i = 0; i < 4; ++i) if (base + i < n) out[base + i] = z[i];
}
}  // namespace

extern "C" int fc_change_score(const float* lp10, const float* lp00, float* out, int B, int N, float multiple,
                               int use_hard_cutoff, float hard_cutoff, fc_stream_t stream) {
    FC_REQUIRE(lp10 && lp00 && out && B > 0 && N > 1);
    change_score_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(lp10, lp00, out, N, (long long)B * N, multiple,
                                                                 use_hard_cutoff, hard_cutoff);
    fc_count_launch();
    FC_LAUNCH_OK();
    return FC_OK;
}

extern "C" int fc_fill_normal(float* out, int64_t n, uint64_t seed, uint64_t offset, fc_stream_t stream) {
    FC_REQUIRE(out && n > 0);
    const long long threads = (n + 3) / 4;
    fill_normal_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out, n, seed, offset);
    fc_count_launch();
    FC_LAUNCH_OK();
    return FC_OK;
}

// ------------------------------------------------------------------------------------------ inner_loop
namespace {
struct LoopWs { float* ctx; void* emb_ws; void* flow_ws; int64_t emb_bytes, flow_bytes, total; };
LoopWs carve_loop(const fc_embedder* e, const fc_flow* f, int B, int N, int Nc, void* base) {
    LoopWs w{};
    const int64_t ctx_floats = (e->kind == 1) ? (int64_t)B * e->E : (int64_t)B * Nc * e->E;
    w.emb_bytes = fc_embedder_workspace_bytes(e, B, Nc);
    w.flow_bytes = fc_flow_workspace_bytes(f, B, N, Nc);
    int64_t off = 0;
    char* b = reinterpret_cast<char*>(base);
    w.ctx = reinterpret_cast<float*>(b + off); off += fc_round_up_ll(ctx_floats * 4, 256);
    // the embedder runs to completion before the flow starts (same stream): the two scratch areas overlap
    w.emb_ws = b + off; w.flow_ws = b + off;
    off += (w.emb_bytes > w.flow_bytes ? w.emb_bytes : w.flow_bytes);
    w.total = off;
    return w;
}
}  // namespace

extern "C" int64_t fc_inner_loop_workspace_bytes(const fc_embedder* e, const fc_flow* f, int B, int N, int Nc) {
    if (!e || !f || B <= 0 || N <= 0 || Nc <= 0) return FC_ERR_INVALID_ARG;
    if (fc_embedder_workspace_bytes(e, B, Nc) < 0) return FC_ERR_INVALID_ARG;
    return carve_loop(e, f, B, N, Nc, nullptr).total;
}

extern "C" int fc_inner_loop(const fc_embedder* e, const fc_flow* f, const float* extract_0, const float* extract_1,
                             const float* extra, const float* eps, float* log_prob_out, float* stats_out, int B, int N,
                             int Nc, void* workspace, int64_t workspace_bytes, int precision, fc_stream_t stream) {
    FC_REQUIRE(e && f && extract_0 && extract_1 && eps && log_prob_out);
    FC_REQUIRE(e->E == f->E && (e->kind == 1) == (f->is_global != 0) && e->d_in == f->d_in);
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255)) return FC_ERR_WORKSPACE;
    LoopWs w = carve_loop(e, f, B, N, Nc, workspace);
    if (w.total > workspace_bytes) return FC_ERR_WORKSPACE;
    int rc = fc_embed(e, extract_0, w.ctx, B, Nc, nullptr, w.emb_ws, w.emb_bytes, precision, stream);
    if (rc) return rc;
    rc = fc_flow_log_prob(f, extract_1, w.ctx, extra, eps, log_prob_out, B, N, Nc, w.flow_ws, w.flow_bytes, precision,
                          stream);
    if (rc) return rc;
    if (stats_out) {
        stats_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(log_prob_out, (long long)B * N, f->d_in, stats_out);
        fc_count_launch();
        FC_LAUNCH_OK();
    }
    return FC_OK;
}

namespace {
struct HostWs { float *e0, *e1, *extra, *eps, *lp, *stats; void* inner; int64_t inner_bytes, total; };
HostWs carve_host(const fc_embedder* e, const fc_flow* f, int B, int N, int Nc, void* base) {
    HostWs w{};
    int64_t off = 0;
    char* b = reinterpret_cast<char*>(base);
    auto take = [&](int64_t floats) { float* p = reinterpret_cast<float*>(b + off); off += fc_round_up_ll(floats * 4, 256); return p; };
    w.e0 = take((int64_t)B * Nc * f->d_in);
    w.e1 = take((int64_t)B * N * f->d_in);
    w.extra = take(B);
    w.eps = take((int64_t)B * N * (f->D - f->d_in));
    w.lp = take((int64_t)B * N);
    w.stats = take(2);
    w.inner = b + off;
    w.inner_bytes = fc_inner_loop_workspace_bytes(e, f, B, N, Nc);
    off += w.inner_bytes;
    w.total = off;
    return w;
}
}  // namespace

extern "C" int64_t fc_inner_loop_host_workspace_bytes(const fc_embedder* e, const fc_flow* f, int B, int N, int Nc) {
    if (!e || !f || B <= 0 || N <= 0 || Nc <= 0) return FC_ERR_INVALID_ARG;
    return carve_host(e, f, B, N, Nc, nullptr).total;
}

extern "C" int fc_inner_loop_host(const fc_embedder* e, const fc_flow* f, const float* e0_h, const float* e1_h,
                                  const float* extra_h, const float* eps_h, float* lp_h, float* stats_h, int B, int N,
                                  int Nc, void* workspace, int64_t workspace_bytes, int precision, fc_stream_t stream_) {
    FC_REQUIRE(e && f && e0_h && e1_h && eps_h && lp_h);
    FC_REQUIRE((f->extra != 0) == (extra_h != nullptr));
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255)) return FC_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream_;
    HostWs w = carve_host(e, f, B, N, Nc, workspace);
    if (w.total > workspace_bytes) return FC_ERR_WORKSPACE;
    FC_CUDA_OK(cudaMemcpyAsync(w.e0, e0_h, (size_t)B * Nc * f->d_in * 4, cudaMemcpyHostToDevice, s));
    FC_CUDA_OK(cudaMemcpyAsync(w.e1, e1_h, (size_t)B * N * f->d_in * 4, cudaMemcpyHostToDevice, s));
    if (extra_h) FC_CUDA_OK(cudaMemcpyAsync(w.extra, extra_h, (size_t)B * 4, cudaMemcpyHostToDevice, s));
    FC_CUDA_OK(cudaMemcpyAsync(w.eps, eps_h, (size_t)B * N * (f->D - f->d_in) * 4, cudaMemcpyHostToDevice, s));
    int rc = fc_inner_loop(e, f, w.e0, w.e1, extra_h ? w.extra : nullptr, w.eps, w.lp, w.stats, B, N, Nc, w.inner,
                           w.inner_bytes, precision, stream_);
    if (rc) return rc;
    FC_CUDA_OK(cudaMemcpyAsync(lp_h, w.lp, (size_t)B * N * 4, cudaMemcpyDeviceToHost, s));
    if (stats_h) FC_CUDA_OK(cudaMemcpyAsync(stats_h, w.stats, 2 * 4, cudaMemcpyDeviceToHost, s));
    FC_CUDA_OK(cudaStreamSynchronize(s));
    return FC_OK;
}
