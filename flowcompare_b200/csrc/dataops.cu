// Data-side ops feeding the path: what the reference's loader does on the CPU between reading a voxel and
// calling `inner_loop`:
//   furthest point sub-sampling to n_samples / n_samples_context points over ALL point columns (xyz + rgb), first pick = point 0
//     reference dataloaders/ams_voxel_loader.py:298-307 -> torch_cluster.fps(x, batch, ratio, random_start=False)
//     (torch_cluster 1.5.9, environment.yml; absent here: its published algorithm -- squared Euclidean distance to the last
//     pick, running minimum, arg-max with the FIRST maximum winning -- is restated in oracle/dataops_ref.py)
//   joint zero-mean / unit-ball normalisation of the xyz columns of a cloud pair
//     reference utils.py:259-280 `co_unit_sphere` over `unit_sphere` (called by ams_voxel_loader.py:357-358)
#include "model.cuh"

namespace {

// One CTA per cloud, every thread keeps its points and their running minimum distances in registers (see fps_kernel in
// paconv.cu for the design); C columns per point, ties to the LOWEST index.  The squared distance is summed column by column
// with separate multiplies and adds (no FMA contraction): the exact arithmetic of the numpy / torch restatement, so the picked
// indices can be compared bit for bit.
template <int PT, int C>
__global__ void __launch_bounds__(1024) fps_nd_kernel(const float* __restrict__ pts, int ld, int n, int m,
                                                      int32_t* __restrict__ idx, float* __restrict__ out, int ldo) {
    __shared__ unsigned long long part[2][32];
    __shared__ float pick[2][C];
    const int b = blockIdx.x, tid = threadIdx.x, bs = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = bs >> 5;
    const float* x = pts + (size_t)b * n * ld;
    float p[PT][C], dist[PT];
#pragma unroll
    for (int i = 0; i < PT; ++i) {
        const int k = tid + i * bs;
        dist[i] = k < n ? 3.0e38f : -1.f;
#pragma unroll
        for (int c = 0; c < C; ++c) p[i][c] = k < n ? x[(size_t)k * ld + c] : 0.f;
    }
    if (tid < C) pick[0][tid] = x[tid];
    if (tid == 0) idx[(size_t)b * m] = 0;
    if (out && tid < C) out[(size_t)b * m * ldo + tid] = x[tid];
    __syncthreads();
    for (int j = 1; j < m; ++j) {
        float o[C];
#pragma unroll
        for (int c = 0; c < C; ++c) o[c] = pick[(j - 1) & 1][c];
        unsigned long long best = 0ull;
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            const int k = tid + i * bs;
            float d0 = p[i][0] - o[0];
            float s = __fmul_rn(d0, d0);
#pragma unroll
            for (int c = 1; c < C; ++c) { const float d = p[i][c] - o[c]; s = __fadd_rn(s, __fmul_rn(d, d)); }
            const float d2 = fminf(s, dist[i]);
            if (k < n) {
                dist[i] = d2;
                // key = (distance bits, ~index): distances are >= 0 so their bit patterns order like the values; lowest index wins ties
                const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)(~(unsigned)k);
                if (key > best) best = key;
            }
        }
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) { const unsigned long long v = __shfl_xor_sync(0xffffffffu, best, of); best = v > best ? v : best; }
        if (lane == 0) part[j & 1][warp] = best;
        __syncthreads();
        best = lane < nwarps ? part[j & 1][lane] : 0ull;
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) { const unsigned long long v = __shfl_xor_sync(0xffffffffu, best, of); best = v > best ? v : best; }
        const int sel = (int)(~(unsigned)best);
        // the owner of the pick publishes its coordinates for the next iteration (double buffered: one barrier per pick)
#pragma unroll
        for (int i = 0; i < PT; ++i)
            if (tid + i * bs == sel) {
#pragma unroll
                for (int c = 0; c < C; ++c) pick[j & 1][c] = p[i][c];
            }
        if (tid == 0) idx[(size_t)b * m + j] = sel;
        __syncthreads();
        if (out && tid < C) out[((size_t)b * m + j) * ldo + tid] = pick[j & 1][tid];
    }
}

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    __syncthreads();
    return t;
}

// One CTA per cloud pair: mean of the xyz columns over BOTH clouds, subtract, divide by the largest norm -- in that order and
// with a true division, as utils.py:259-267 does.  inverse[b] = (mean x, mean y, mean z, furthest_distance).
__global__ void co_unit_sphere_kernel(float* __restrict__ p0, int n0, int ld0, float* __restrict__ p1, int n1, int ld1,
                                      float* __restrict__ inverse) {
    __shared__ double sh[32];
    __shared__ float shf[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    float* a = p0 + (size_t)b * n0 * ld0;
    float* c = p1 + (size_t)b * n1 * ld1;
    const int n = n0 + n1;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int i = tid; i < n; i += blockDim.x) {
        const float* q = i < n0 ? a + (size_t)i * ld0 : c + (size_t)(i - n0) * ld1;
        sx += q[0]; sy += q[1]; sz += q[2];
    }
    const float mx = (float)(block_sum_d(sx, sh) / n), my = (float)(block_sum_d(sy, sh) / n), mz = (float)(block_sum_d(sz, sh) / n);
    float far2 = 0.f;
    for (int i = tid; i < n; i += blockDim.x) {
        float* q = i < n0 ? a + (size_t)i * ld0 : c + (size_t)(i - n0) * ld1;
        const float x = q[0] - mx, y = q[1] - my, z = q[2] - mz;
        q[0] = x; q[1] = y; q[2] = z;
        far2 = fmaxf(far2, __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    }
    far2 = fc_warp_max(far2);
    if ((tid & 31) == 0) shf[tid >> 5] = far2;
    __syncthreads();
    float m = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, shf[i]);
    const float far = sqrtf(m);
    for (int i = tid; i < n; i += blockDim.x) {     // each thread re-reads the points it wrote itself
        float* q = i < n0 ? a + (size_t)i * ld0 : c + (size_t)(i - n0) * ld1;
        q[0] = q[0] / far; q[1] = q[1] / far; q[2] = q[2] / far;
    }
    if (inverse && tid == 0) { float* o = inverse + (size_t)b * 4; o[0] = mx; o[1] = my; o[2] = mz; o[3] = far; }
}

template <int C>
int launch_fps_nd(const float* pts, int ld, int B, int n, int m, int32_t* idx, float* out, int ldo, cudaStream_t s) {
    int threads = fc_round_up(n, 32);
    if (threads > 1024) threads = 1024;
    const int pt = (n + threads - 1) / threads;
#define FC_CASE(P) fps_nd_kernel<P, C><<<B, threads, 0, s>>>(pts, ld, n, m, idx, out, ldo)
    if (pt <= 1) FC_CASE(1); else if (pt <= 2) FC_CASE(2); else if (pt <= 4) FC_CASE(4);
    else if (pt <= 8) FC_CASE(8); else if (pt <= 16) FC_CASE(16); else FC_CASE(32);
#undef FC_CASE
    fc_count_launch(); FC_LAUNCH_OK();
    return FC_OK;
}

}  // namespace

extern "C" int fc_fps_points(const float* pts, int ld, int B, int n, int C, int m, int32_t* idx_out, float* pts_out, int ld_out,
                             fc_stream_t stream) {
    FC_REQUIRE(pts && idx_out && B > 0 && B <= 65535 && n > 0 && n <= 32768 && m > 0 && m <= n && ld >= C && (!pts_out || ld_out >= C));
    cudaStream_t s = (cudaStream_t)stream;
    if (C == 3) return launch_fps_nd<3>(pts, ld, B, n, m, idx_out, pts_out, ld_out, s);
    if (C == 6) return launch_fps_nd<6>(pts, ld, B, n, m, idx_out, pts_out, ld_out, s);
    return FC_ERR_UNSUPPORTED;
}

extern "C" int fc_co_unit_sphere(float* points_0, int n0, int ld0, float* points_1, int n1, int ld1, int B, float* inverse_out,
                                 fc_stream_t stream) {
    FC_REQUIRE(points_0 && points_1 && n0 > 0 && n1 > 0 && ld0 >= 3 && ld1 >= 3 && B > 0 && B <= 65535);
    co_unit_sphere_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(points_0, n0, ld0, points_1, n1, ld1, inverse_out);
    fc_count_launch(); FC_LAUNCH_OK();
    return FC_OK;
}
