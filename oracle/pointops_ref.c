/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the three index-producing pointops kernels the PAConv
 * embedder uses.  The reference's own native module cannot be built here
 * (it includes <THC/THC.h>, removed from torch >= 1.11) and is CUDA-only, so these follow the algorithms of
 *   K2 furthest point sampling   lib/pointops/src/sampling/sampling_cuda_kernel.cu:58-168 (+ launcher :170-209)
 *   K1 heap kNN                  lib/pointops/src/knnquery_heap/knnquery_heap_cuda_kernel.cu:21-89
 *   K5 three nearest neighbours  lib/pointops/src/interpolation/interpolation_cuda_kernel.cu:134-176
 * including their tie behaviour (strict comparisons, heap order, block-strided arg-max), so that the CUDA
 * kernels in flowcompare_b200/csrc/paconv.cu can be checked bit-for-bit.  Distances use the contraction nvcc
 * applies to `dx*dx + dy*dy + dz*dz`: fmaf(dz, dz, fmaf(dx, dx, dy*dy)) -- read off the SASS of the reference's own
 * kernels compiled for sm_100a (FMUL dy*dy; FFMA dx,dx; FFMA dz,dz in all three).  Pinned: on a GPU box
 * tests/test_pointops_gpu.py compares the CUDA kernels AND this restatement with the reference's compiled kernels
 * (oracle/_ref/libpointops_ref.so, built by oracle/build_ref_pointops.sh from the reference's unmodified sources),
 * indices and distances bit for bit, including clouds with exact duplicate points.                              */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static float sqdist3(const float* a, const float* b) {
    const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* threads per block the reference launcher picks: 2^floor(log2 n) via double log, clipped to [1,1024] */
int fc_oracle_fps_block(int n) {
    const int p = (int)(log((double)n) / log(2.0));
    int t = 1 << p;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

/* xyz [n][3] -> idx [m]; start at index 0; running min distance; block arg-max with the reference's
 * tie rule: within a thread the first (lowest k) maximum, across threads the lower thread id. */
int fc_oracle_fps(const float* xyz, int n, int m, int32_t* idx) {
    if (m <= 0) return 0;
    const int bs = fc_oracle_fps_block(n);
    float* temp = (float*)malloc(sizeof(float) * (size_t)n);
    float* bv = (float*)malloc(sizeof(float) * (size_t)bs);
    int* bi = (int*)malloc(sizeof(int) * (size_t)bs);
    if (!temp || !bv || !bi) { free(temp); free(bv); free(bi); return -1; }
    for (int k = 0; k < n; ++k) temp[k] = 1e10f;
    int old = 0;
    idx[0] = 0;
    for (int j = 1; j < m; ++j) {
        for (int tid = 0; tid < bs; ++tid) {
            float best = -1.0f; int besti = 0;
            for (int k = tid; k < n; k += bs) {
                const float d = sqdist3(xyz + 3 * (size_t)k, xyz + 3 * (size_t)old);
                const float d2 = d < temp[k] ? d : temp[k];
                temp[k] = d2;
                if (d2 > best) { best = d2; besti = k; }
            }
            bv[tid] = best; bi[tid] = besti;
        }
        for (int half = bs / 2; half >= 1; half /= 2)
            for (int tid = 0; tid < half; ++tid) {
                const float v1 = bv[tid], v2 = bv[tid + half];
                if (v2 > v1) { bv[tid] = v2; bi[tid] = bi[tid + half]; }
            }
        old = bi[0];
        idx[j] = old;
    }
    free(temp); free(bv); free(bi);
    return 0;
}

static void sift_down(float* d, int32_t* ix, int size) {
    int root = 0;
    for (;;) {
        int child = 2 * root + 1;
        if (child >= size) return;
        if (child + 1 < size && d[child + 1] > d[child]) ++child;
        if (d[root] > d[child]) return;
        const float td = d[root]; d[root] = d[child]; d[child] = td;
        const int32_t ti = ix[root]; ix[root] = ix[child]; ix[child] = ti;
        root = child;
    }
}

/* xyz [n][3], queries [m][3] -> idx [m][k]: max-heap of the k best (strict < replaces the root), then
 * heap-sorted ascending; slots never filled (k > n) keep index 0. */
int fc_oracle_knn_heap(const float* xyz, const float* queries, int n, int m, int k, int32_t* idx) {
    if (k < 1 || k > 100) return -1;
    float d[100]; int32_t ix[100];
    for (int q = 0; q < m; ++q) {
        for (int i = 0; i < k; ++i) { d[i] = 1e10f; ix[i] = 0; }
        for (int i = 0; i < n; ++i) {
            const float d2 = sqdist3(queries + 3 * (size_t)q, xyz + 3 * (size_t)i);
            if (d2 < d[0]) { d[0] = d2; ix[0] = i; sift_down(d, ix, k); }
        }
        for (int i = k - 1; i > 0; --i) {
            const float td = d[0]; d[0] = d[i]; d[i] = td;
            const int32_t ti = ix[0]; ix[0] = ix[i]; ix[i] = ti;
            sift_down(d, ix, i);
        }
        for (int i = 0; i < k; ++i) idx[(size_t)q * k + i] = ix[i];
    }
    return 0;
}

/* unknown [n][3], known [m][3] -> dist2 [n][3] (squared), idx [n][3]; strict < insertion (lower index wins ties) */
int fc_oracle_three_nn(const float* unknown, const float* known, int n, int m, float* dist2, int32_t* idx) {
    for (int p = 0; p < n; ++p) {
        double b1 = 1e40, b2 = 1e40, b3 = 1e40;
        int i1 = 0, i2 = 0, i3 = 0;
        for (int k = 0; k < m; ++k) {
            const float d = sqdist3(unknown + 3 * (size_t)p, known + 3 * (size_t)k);
            if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
            else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
            else if (d < b3) { b3 = d; i3 = k; }
        }
        dist2[3 * (size_t)p + 0] = (float)b1; dist2[3 * (size_t)p + 1] = (float)b2; dist2[3 * (size_t)p + 2] = (float)b3;
        idx[3 * (size_t)p + 0] = i1; idx[3 * (size_t)p + 1] = i2; idx[3 * (size_t)p + 2] = i3;
    }
    return 0;
}
