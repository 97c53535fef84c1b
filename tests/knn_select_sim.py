"""TEST INFRASTRUCTURE -- a lane-by-lane numpy model of `knn_select_kernel` (flowcompare_b200/csrc/knn.cu): one warp per query
row, 1280-key segments, lower bound from the 96 lane-local top-3 keys, prefix-sum compaction of the survivors, sort, carried
list; and the exact-bisection path taken when more than SEL_CAP keys survive (masses of equal keys).  It follows the kernel
statement by statement (same per-lane registers, ballots and prefix sums) so that the ALGORITHM -- in particular the order of
equal keys -- can be checked on a machine without a GPU against a plain (key descending, index ascending) sort.
Used by tests/test_kernel_constants.py; nothing in flowcompare_b200/ imports it."""
import re
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def constants():
    src = open(os.path.join(ROOT, "flowcompare_b200", "csrc", "knn.cu")).read()
    r = int(re.search(r"constexpr int SEL_R = (\d+);", src).group(1))
    cap = int(re.search(r"constexpr int SEL_CAP = (\d+);", src).group(1))
    return r, cap


def ordered_words(x):
    """knn_ord: monotone float -> unsigned (0 is below every float)."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    neg = (u & 0x80000000) != 0
    return np.where(neg, (~u) & 0xFFFFFFFF, u | 0x80000000).astype(np.uint64)


def select(keys, k, stats=None):
    """keys [Nt] float32 -> the kernel's k indices for this row."""
    SEL_R, SEL_CAP = constants()
    SEG = 32 * SEL_R
    Nt = len(keys)
    words = ordered_words(keys)
    lanes = np.arange(32)
    sv = np.zeros(SEL_CAP, dtype=np.uint64)
    have = 0
    for seg0 in range(0, Nt, SEG):
        e = seg0 + np.arange(SEL_R)[:, None] * 32 + lanes[None, :]                 # [i, lane]
        r = np.where(e < Nt, words[np.minimum(e, Nt - 1)], 0).astype(np.uint64)
        m1 = np.zeros(32, np.uint64); m2 = m1.copy(); m3 = m1.copy()
        for i in range(SEL_R):
            a = np.minimum(m1, r[i]); m1 = np.maximum(m1, r[i])
            c = np.minimum(m2, a); m2 = np.maximum(m2, a)
            m3 = np.maximum(m3, c)
        w0 = np.where(lanes < have, sv[lanes], 0).astype(np.uint64)
        w1 = np.where(32 + lanes < have, sv[32 + lanes], 0).astype(np.uint64)
        l0, l1 = w0 >> np.uint64(32), w1 >> np.uint64(32)
        L = 0
        for bit in range(31, -1, -1):
            cand = L | (1 << bit)
            if int((m1 >= cand).sum() + (m2 >= cand).sum() + (m3 >= cand).sum()) >= k:
                L = cand
        if have >= k:
            L = max(L, int(sv[k - 1] >> np.uint64(32)))
        thr = L - 1 if L else 0
        cnt_l = (l0 > thr).astype(int) + (l1 > thr).astype(int) + (r > thr).sum(axis=0)
        tot = int(cnt_l.sum())
        new = np.zeros(SEL_CAP, dtype=np.uint64)
        word = (r << np.uint64(32)) | ((~e.astype(np.uint64)) & np.uint64(0xFFFFFFFF))
        if tot <= SEL_CAP:
            off = np.cumsum(cnt_l) - cnt_l
            for lane in range(32):
                o = int(off[lane])
                if l0[lane] > thr: new[o] = w0[lane]; o += 1
                if l1[lane] > thr: new[o] = w1[lane]; o += 1
                for i in range(SEL_R):
                    if r[i, lane] > thr: new[o] = word[i, lane]; o += 1
            base = tot
            if stats is not None: stats["compaction"] = stats.get("compaction", 0) + 1
        else:
            Tk = 0
            for bit in range(31, -1, -1):
                cand = Tk | (1 << bit)
                if int((l0 >= cand).sum() + (l1 >= cand).sum() + (r >= cand).sum()) >= k:
                    Tk = cand
            need = k - int((l0 > Tk).sum() + (l1 > Tk).sum() + (r > Tk).sum())
            base = 0

            def push(gt, eq, w):
                nonlocal need, base
                before = np.cumsum(eq) - eq                      # popc(meq & lt)
                take = gt | (eq & (before < need))
                need = max(0, need - int(eq.sum()))
                pos = np.cumsum(take) - take
                for lane in range(32):
                    if take[lane]: new[base + int(pos[lane])] = w[lane]
                base += int(take.sum())

            push(l0 > Tk, (l0 == Tk) & (lanes < have), w0)
            push(l1 > Tk, (l1 == Tk) & (32 + lanes < have), w1)
            for i in range(SEL_R):
                push(r[i] > Tk, (r[i] == Tk) & (e[i] < Nt), word[i])
            if stats is not None: stats["bisection"] = stats.get("bisection", 0) + 1
        n_sorted = 64 if base <= 64 else (128 if base <= 128 else 256)
        v = np.zeros(n_sorted, dtype=np.uint64)
        v[:base] = new[:base]
        v = np.sort(v)[::-1]                                     # the bitonic network: descending 64-bit words
        sv = new
        sv[:n_sorted] = v
        have = min(base, k)
    out = np.zeros(k, dtype=np.int64)
    for p in range(k):
        if p < have:
            out[p] = int((~sv[p]) & np.uint64(0xFFFFFFFF))
    return out


def reference(keys, k):
    """k largest keys, ties by lower index."""
    w = ordered_words(keys)
    order = np.lexsort((np.arange(len(keys)), -w.astype(np.int64)))
    return order[:k]
