// Elementwise tails of the transforms north_star names but no shipped config builds:
//   rational-quadratic spline coupling   reference models/spline_coupling.py:24-169, :187-227
//   exponential coupling                 reference models/exponential_coupling.py:44-77 (+ utils.py:300-327 `expm`)
//   CIF block pieces                     reference models/cif_block.py:71-111: ConditionalNormal sample / score
//                                        (models/distributions.py:128-153, models/augmenter.py:49-63, models/slice.py:31-44),
//                                        ActNorm between two Reverse permutations (folded to one per-column affine map)
// Each kernel reads the conditioner's raw output (the last GEMM of its MLP, plain store epilogue), updates the latent
// in place and keeps the row's log-det in registers until ONE owner thread adds it to the row's partial-sum slab
// (deterministic: no atomics).  They are bandwidth-bound passes over the parameters (spline: 100 B per latent element,
// exponential: n^2 + n floats per point).
#include "model.cuh"

namespace {

__device__ __forceinline__ float fc_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // F.softplus, threshold 20

constexpr int RQ_MAXB = 16;
constexpr float RQ_MIN_W = 1e-3f, RQ_MIN_H = 1e-3f, RQ_MIN_D = 1e-3f;   // spline_coupling.py:12-14

// knots of one axis: softmax -> min size + rescale -> cumulative sum -> [lo, hi] with exact ends (spline_coupling.py:87-94)
__device__ __forceinline__ void rq_knots(const float* __restrict__ u, int nb, float min_size, float lo, float hi, float (&cum)[RQ_MAXB + 1]) {
    float v[RQ_MAXB];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < RQ_MAXB; ++k) { v[k] = k < nb ? u[k] : -INFINITY; mx = fmaxf(mx, v[k]); }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < RQ_MAXB; ++k) { v[k] = k < nb ? expf(v[k] - mx) : 0.f; sum += v[k]; }
    const float inv = 1.0f / sum, room = 1.0f - min_size * (float)nb;
    float c = 0.f;
    cum[0] = lo;
#pragma unroll
    for (int k = 0; k < RQ_MAXB; ++k) {
        if (k < nb) { c += min_size + room * (v[k] * inv); cum[k + 1] = (k == nb - 1) ? hi : (hi - lo) * c + lo; }
        else cum[k + 1] = hi;
    }
}

// One warp per row, lanes stride over the row's n2 transformed elements; element j reads its 3*nb+1 parameters
// [widths nb | heights nb | derivatives nb+1] (the reference's reshape + split, spline_coupling.py:197-198).
__global__ void rq_spline_kernel(const float* __restrict__ params, int ldp, float* __restrict__ lat, int ldx, int col0,
                                 int n2, int nb, int M, float tail, float d_edge, float* __restrict__ part, int inverse) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const int stride = 3 * nb + 1;
    float ldj = 0.f;
    for (int j = lane; j < n2; j += 32) {
        float* xp = lat + (size_t)row * ldx + col0 + j;
        const float x = *xp;
        if (!(x >= -tail && x <= tail)) continue;            // linear tails: identity, log-det 0 (spline_coupling.py:35-48)
        const float* p = params + (size_t)row * ldp + (size_t)j * stride;
        float cw[RQ_MAXB + 1], ch[RQ_MAXB + 1];
        rq_knots(p, nb, RQ_MIN_W, -tail, tail, cw);
        rq_knots(p + nb, nb, RQ_MIN_H, -tail, tail, ch);
        // searchsorted (spline_coupling.py:17-19): count the knots <= x, the last knot moved up by 1e-6
        int bin = -1;
#pragma unroll
        for (int k = 0; k <= RQ_MAXB; ++k) {
            if (k <= nb) {
                float loc = inverse ? ch[k] : cw[k];
                if (k == nb) loc += 1e-6f;
                bin += (x >= loc) ? 1 : 0;
            }
        }
        float in_cw = 0.f, in_w = 1.f, in_ch = 0.f, in_h = 1.f, ud0 = 0.f, ud1 = 0.f;
#pragma unroll
        for (int k = 0; k < RQ_MAXB; ++k) {
            if (k == bin) { in_cw = cw[k]; in_w = cw[k + 1] - cw[k]; in_ch = ch[k]; in_h = ch[k + 1] - ch[k]; }
        }
        // derivatives (spline_coupling.py:42-45,96): the reference pads the nb+1 network outputs on BOTH sides, so knot 0
        // takes the tail constant (d_edge, see the launcher) and knot k >= 1 takes output k-1; the last output is never read
        if (bin > 0) ud0 = p[2 * nb + bin - 1];
        ud1 = p[2 * nb + bin];
        const float d = bin == 0 ? d_edge : RQ_MIN_D + fc_softplus(ud0);
        const float d1 = RQ_MIN_D + fc_softplus(ud1);
        const float delta = in_h / in_w;
        const float dd = d + d1 - 2.0f * delta;
        if (!inverse) {
            const float th = (x - in_cw) / in_w;
            const float tt = th * (1.0f - th);
            const float den = delta + dd * tt;
            *xp = in_ch + in_h * (delta * th * th + d * tt) / den;
            const float num = delta * delta * (d1 * th * th + 2.0f * delta * tt + d * (1.0f - th) * (1.0f - th));
            ldj += logf(num) - 2.0f * logf(den);
        } else {
            const float a = (x - in_ch) * dd + in_h * (delta - d);
            const float b = in_h * d - (x - in_ch) * dd;
            const float c = -delta * (x - in_ch);
            const float root = (2.0f * c) / (-b - sqrtf(b * b - 4.0f * a * c));
            *xp = root * in_w + in_cw;
        }
    }
    if (!inverse && part) {
        ldj = fc_warp_sum(ldj);
        if (lane == 0) part[row] += ldj;
    }
}

// ConditionalNormal with parameters [mean S | log_std S]: mode 0 draws z = mean + sigma*eps into lat[:, col0:col0+S] and adds
// -log q(z) (Augment.forward: ldj = -log q); mode 1 adds +log N(lat[:, col0:col0+S]; mean, sigma) (Slice.forward).
// sigma = min(exp(log_std), clamp)  (distributions.py:134-136).  One warp per row.
__global__ void cond_normal_kernel(const float* __restrict__ params, int ldp, float* __restrict__ lat, int ldx, int col0,
                                   int S, const float* __restrict__ eps, int M, float clamp, float* __restrict__ part, int mode) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* p = params + (size_t)row * ldp;
    float ldj = 0.f;
    for (int j = lane; j < S; j += 32) {
        const float mean = p[j];
        const float sigma = fminf(expf(p[S + j]), clamp);
        float* zp = lat + (size_t)row * ldx + col0 + j;
        float e;
        if (mode == 0) { e = eps[(size_t)row * S + j]; *zp = fmaf(sigma, e, mean); }
        else e = (*zp - mean) / sigma;
        const float nlq = fmaf(0.5f * e, e, logf(sigma)) + 0.91893853320467274178f;   // -log N
        ldj += mode == 0 ? nlq : -nlq;
    }
    ldj = fc_warp_sum(ldj);
    if (lane == 0) part[row] += ldj;
}

__global__ void col_affine_kernel(float* __restrict__ lat, int ldx, int cols, long long M, const float* __restrict__ sc,
                                  const float* __restrict__ bi) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * cols) return;
    const long long r = i / cols; const int c = (int)(i % cols);
    float* p = lat + r * ldx + c;
    *p = fmaf(*p, sc[c], bi[c]);
}

// y2 = expm(W) x2 + b, ldj = trace W, W = rescale*tanh(scale*w + shift) + reshift + 1e-8 (exponential_coupling.py:48-58).
// The reference materialises expm(W) per point (n^3 work per product); only its ACTION on one vector is needed, so this
// kernel applies the scaled Taylor polynomial to the vector: with s = ceil(log2 ||W||_inf) halvings,
// y = (sum_k (W/2^s)^k / k!)^(2^s) x, i.e. 2^s * m matrix-vector products (n^2 each) with W held in REGISTERS (TPR threads
// per matrix row, 16-byte broadcast loads of the vector from shared memory).  One CTA per point.
constexpr int EX_PER_MAX = 80;
template <int TPR>
__global__ void __launch_bounds__(1024) expm_action_kernel(const float* __restrict__ params, int ldp, float* __restrict__ lat,
                                                           int ldx, int col0, int n2, const float* __restrict__ sq,
                                                           float* __restrict__ part, long long row0, int inverse) {
    extern __shared__ __align__(16) float ex_sh[];
    __shared__ float red[33];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int i = t / TPR, pid = t % TPR;
    const bool active = i < n2;
    const int per = ((n2 + TPR - 1) / TPR + 3) & ~3;       // elements of the row this thread owns (multiple of 4)
    const int j0 = pid * per;
    const int vlen = TPR * per;
    float* cur = ex_sh; float* nxt = ex_sh + vlen;
    const long long prow = blockIdx.x;
    const long long row = row0 + prow;
    const float* p = params + prow * (long long)ldp;
    const float scale = sq[0], shift = sq[1], rescale = sq[2], reshift = sq[3];
    float w[EX_PER_MAX];
    float tr = 0.f, asum = 0.f;
#pragma unroll
    for (int e = 0; e < EX_PER_MAX; ++e) {
        const int j = j0 + e;
        float v = 0.f;
        if (active && e < per && j < n2) {
            v = fmaf(rescale, tanhf(fmaf(scale, p[(size_t)i * n2 + j], shift)), reshift) + 1e-8f;
            if (j == i) tr += v;
            asum += fabsf(v);
        }
        w[e] = inverse ? -v : v;
    }
    for (int o = TPR >> 1; o > 0; o >>= 1) asum += __shfl_xor_sync(0xffffffffu, asum, o);
    // block reductions: max row sum (infinity norm) and trace
    float nmax = fc_warp_max(asum), trs = fc_warp_sum(tr);
    if (lane == 0) { red[warp] = nmax; }
    __syncthreads();
    if (t == 0) { float m = 0.f; for (int k = 0; k < (int)((blockDim.x + 31) >> 5); ++k) m = fmaxf(m, red[k]); red[32] = m; }
    __syncthreads();
    float norm = red[32];
    __syncthreads();
    if (lane == 0) red[warp] = trs;
    // the vector: x2 (forward) or y2 - b (inverse)
    for (int k = t; k < vlen; k += blockDim.x) {
        float v = 0.f;
        if (k < n2) { v = lat[row * ldx + col0 + k]; if (inverse) v -= p[(size_t)n2 * n2 + k]; }
        cur[k] = v; nxt[k] = 0.f;
    }
    __syncthreads();
    if (t == 0 && part && !inverse) { float s = 0.f; for (int k = 0; k < (int)((blockDim.x + 31) >> 5); ++k) s += red[k]; part[row] += s; }
    int s = 0;
    while (norm > 1.0f && s < 24) { norm *= 0.5f; ++s; }
    const float sc2 = ldexpf(1.0f, -s);
#pragma unroll
    for (int e = 0; e < EX_PER_MAX; ++e) w[e] *= sc2;
    int m = 0;
    { float term = 1.f; while (term > 1e-9f && m < 24) { ++m; term *= norm / (float)m; } }
    for (int rep = 0; rep < (1 << s); ++rep) {
        float acc = active ? cur[i] : 0.f;
        for (int k = 1; k <= m; ++k) {
            float partial = 0.f;
            const float4* c4 = reinterpret_cast<const float4*>(cur + j0);
#pragma unroll
            for (int e = 0; e < EX_PER_MAX; e += 4) {
                if (e < per) {
                    const float4 c = c4[e >> 2];
                    partial = fmaf(w[e], c.x, partial); partial = fmaf(w[e + 1], c.y, partial);
                    partial = fmaf(w[e + 2], c.z, partial); partial = fmaf(w[e + 3], c.w, partial);
                }
            }
            for (int o = TPR >> 1; o > 0; o >>= 1) partial += __shfl_xor_sync(0xffffffffu, partial, o);
            const float tk = partial / (float)k;
            acc += tk;
            if (active && pid == 0) nxt[i] = tk;
            __syncthreads();
            float* sw = cur; cur = nxt; nxt = sw;
        }
        if (active && pid == 0) nxt[i] = acc;
        __syncthreads();
        float* sw = cur; cur = nxt; nxt = sw;
    }
    if (active && pid == 0) {
        float y = cur[i];
        if (!inverse) y += p[(size_t)n2 * n2 + i];
        lat[row * ldx + col0 + i] = y;
    }
}

}  // namespace

int fc_launch_rq_spline(const float* params, int ldp, float* lat, int ldx, int col0, int n2, int nb, int M, float* part,
                        int inverse, cudaStream_t s) {
    FC_REQUIRE(params && lat && nb >= 1 && nb <= RQ_MAXB && n2 > 0 && M > 0);
    // tail derivative, AS THE REFERENCE WRITES IT (spline_coupling.py:43): `math.log(math.exp((1 - min_derivative) - 1))`, i.e.
    // log(exp(-min_derivative)) = -1e-3 -- the parenthesis sits one token early compared with the neural-spline-flows original
    // (log(exp(1 - min_derivative) - 1), which would give an edge derivative of exactly 1); the edge knot's derivative is
    // therefore min_derivative + softplus(-1e-3) = 0.6936..., and parity follows the reference
    const float c = (float)log(exp((1.0 - 1e-3) - 1.0));
    const float d_edge = RQ_MIN_D + (c > 20.f ? c : log1pf(expf(c)));
    const int wpb = 4;
    FcProfScope prof(FC_CLS_OTHER, 0.0, 4.0 * M * ((double)n2 * (3 * nb + 3)), s);
    rq_spline_kernel<<<(M + wpb - 1) / wpb, wpb * 32, 0, s>>>(params, ldp, lat, ldx, col0, n2, nb, M, 3.0f, d_edge, part, inverse);
    fc_count_launch();
    FC_LAUNCH_OK();
    return FC_OK;
}

int fc_launch_cond_normal(const float* params, int ldp, float* lat, int ldx, int col0, int S, const float* eps, int M,
                          float clamp, float* part, int mode, cudaStream_t s) {
    FC_REQUIRE(params && lat && part && S > 0 && M > 0 && (mode == 1 || eps));
    const int wpb = 8;
    cond_normal_kernel<<<(M + wpb - 1) / wpb, wpb * 32, 0, s>>>(params, ldp, lat, ldx, col0, S, eps, M,
                                                                clamp > 0.f ? clamp : INFINITY, part, mode);
    fc_count_launch();
    FC_LAUNCH_OK();
    return FC_OK;
}

int fc_launch_col_affine(float* lat, int ldx, int cols, long long M, const float* sc, const float* bi, cudaStream_t s) {
    FC_REQUIRE(lat && sc && bi && cols > 0 && M > 0);
    const long long tot = M * cols;
    col_affine_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(lat, ldx, cols, M, sc, bi);
    fc_count_launch();
    FC_LAUNCH_OK();
    return FC_OK;
}

int fc_expm_max_n() { return 256; }

int fc_launch_expm_action(const float* params, int ldp, float* lat, int ldx, int col0, int n2, const float* squash4,
                          float* part, long long row0, int rows, int inverse, cudaStream_t s) {
    FC_REQUIRE(params && lat && squash4 && n2 > 0 && n2 <= fc_expm_max_n() && rows > 0);
    const int tpr = n2 <= 2 * EX_PER_MAX ? 2 : 4;
    const int per = ((n2 + tpr - 1) / tpr + 3) & ~3;
    FC_REQUIRE(per <= EX_PER_MAX);
    const int threads = fc_round_up(n2 * tpr, 32);
    const size_t smem = (size_t)2 * tpr * per * sizeof(float);
    FcProfScope prof(FC_CLS_OTHER, 0.0, 4.0 * rows * ((double)n2 * n2 + 3.0 * n2), s);
    if (tpr == 2) expm_action_kernel<2><<<rows, threads, smem, s>>>(params, ldp, lat, ldx, col0, n2, squash4, part, row0, inverse);
    else          expm_action_kernel<4><<<rows, threads, smem, s>>>(params, ldp, lat, ldx, col0, n2, squash4, part, row0, inverse);
    fc_count_launch();
    FC_LAUNCH_OK();
    return FC_OK;
}

// ------------------------------------------------------------------------------------------ op-level C ABI
// `unconstrained_rational_quadratic_spline(inputs, widths, heights, derivatives, inverse)` of the reference
// (models/spline_coupling.py:24-66; 'linear' tails, tail_bound 3): params row = n * (3*num_bins+1) floats, element j's
// [widths | heights | derivatives(num_bins+1)] contiguous (the reshape + split of :197-198); x [M][ldx] is transformed in
// place; forward adds each row's summed log|det| to logabsdet_rowsum[row] (caller zeroes it), inverse leaves it alone.
extern "C" int fc_rq_spline(const float* params, int ldp, float* x, int ldx, int n, int num_bins, int M,
                            float* logabsdet_rowsum, int inverse, fc_stream_t stream) {
    FC_REQUIRE(ldp >= n * (3 * num_bins + 1) && ldx >= n && (inverse || logabsdet_rowsum));
    return fc_launch_rq_spline(params, ldp, x, ldx, 0, n, num_bins, M, logabsdet_rowsum, inverse ? 1 : 0, (cudaStream_t)stream);
}

// `ExponentialCoupling`'s transform of x2 given its conditioner output (reference models/exponential_coupling.py:48-58 forward,
// :68-77 inverse): params row = [w n*n | b n]; W = rescale*tanh(scale*w + shift) + reshift + 1e-8 with squash4 = (scale, shift,
// rescale, reshift) on the device; forward x <- expm(W) x + b and trace_rowsum[row] += trace W (caller zeroes it); inverse
// x <- expm(-W)(x - b).
extern "C" int fc_expm_action(const float* params, int ldp, float* x, int ldx, int n, const float* squash4, float* trace_rowsum,
                              int M, int inverse, fc_stream_t stream) {
    FC_REQUIRE(ldp >= n * n + n && ldx >= n && (inverse || trace_rowsum));
    return fc_launch_expm_action(params, ldp, x, ldx, 0, n, squash4, trace_rowsum, 0, M, inverse ? 1 : 0, (cudaStream_t)stream);
}
