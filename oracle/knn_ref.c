/* TEST INFRASTRUCTURE ONLY -- bit-exact CPU restatement ("oracle B", SURVEY.md hard-part 2) of the
 * canonical arithmetic of the kNN kernels in flowcompare_b200/csrc/knn.cu.
 *
 * The algebraic form is the reference's:
 *   mode 0: reference models/pytorch_gcn.py:13-20   pd = -xx_j - (-2 * <x_i,x_j>) - xx_i, k largest
 *   mode 1: reference knn.py:40-52                  diss = (qq_i + tt_j) - 2 * <q_i,t_j>, k smallest
 * but the reference leaves the summation order of the inner product to BLAS and the order of exactly
 * tied neighbours to torch.topk; here both are pinned: sequential fmaf over the feature index, ties by
 * lower index (stable).  Build: gcc -O2 -ffp-contract=off (see Makefile).  Pinned against the unmodified
 * reference `knn` in tests/test_knn_oracle.py (set equality up to rounding ties, checked in fp64).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct { float key; int32_t idx; } fc_pair;

static int cmp_desc(const void* a, const void* b) {
    const fc_pair* x = (const fc_pair*)a;
    const fc_pair* y = (const fc_pair*)b;
    if (x->key > y->key) return -1;
    if (x->key < y->key) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

static float sq_chain(const float* v, int C) {
    float acc = 0.0f;
    for (int c = 0; c < C; ++c) acc = fmaf(v[c], v[c], acc);
    return acc;
}

/* q [Nq][ldq], t [Nt][ldt]; idx [Nq][k]; returns 0 on success */
int fc_oracle_knn(const float* q, int ldq, const float* t, int ldt, int Nq, int Nt, int C, int k, int mode,
                  int32_t* idx) {
    if (k > Nt || k < 1) return -1;
    fc_pair* buf = (fc_pair*)malloc(sizeof(fc_pair) * (size_t)Nt);
    float* tt = (float*)malloc(sizeof(float) * (size_t)Nt);
    if (!buf || !tt) { free(buf); free(tt); return -2; }
    for (int j = 0; j < Nt; ++j) tt[j] = sq_chain(t + (size_t)j * ldt, C);
    for (int i = 0; i < Nq; ++i) {
        const float* qi = q + (size_t)i * ldq;
        const float qq = sq_chain(qi, C);
        for (int j = 0; j < Nt; ++j) {
            const float* tj = t + (size_t)j * ldt;
            float dot = 0.0f;
            for (int c = 0; c < C; ++c) dot = fmaf(qi[c], tj[c], dot);
            float key;
            if (mode == 0) {
                volatile float inner = -2.0f * dot;
                volatile float t1 = (-tt[j]) - inner;
                key = t1 - qq;
            } else {
                volatile float s = qq + tt[j];
                volatile float d = s - 2.0f * dot;
                key = -d;
            }
            buf[j].key = key;
            buf[j].idx = j;
        }
        qsort(buf, (size_t)Nt, sizeof(fc_pair), cmp_desc);
        for (int e = 0; e < k; ++e) idx[(size_t)i * k + e] = buf[e].idx;
    }
    free(buf); free(tt);
    return 0;
}
