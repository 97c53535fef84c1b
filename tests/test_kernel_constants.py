"""CPU checks of constants / index arithmetic that live inside the CUDA sources (no GPU, no library calls):
the branch-free GELU's polynomial and the N-tile partition of the tcgen05 GEMM."""
import math
import os
import re

import numpy as np

from flowcompare_b200 import packing

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gelu_coefficients():
    src = open(os.path.join(ROOT, "flowcompare_b200", "csrc", "common.cuh")).read()
    body = src[src.index("float fc_gelu_erf_fast(float x)"):]
    body = body[:body.index("return")]
    clamp = float(re.search(r"fminf\(fabsf\(x\), ([0-9.eE+-]+)f\)", body).group(1))
    first = float(re.search(r"float r = ([0-9.eE+-]+)f;", body).group(1))
    rest = [float(v) for v in re.findall(r"r = fmaf\(r, u, ([0-9.eE+-]+)f\);", body)]
    return clamp, [first] + rest


def test_fast_gelu_polynomial_matches_exact_gelu():
    """fp32 emulation of fc_gelu_erf_fast (csrc/common.cuh): relu(x) - |x| * 2^(u r(u) - 1) vs 0.5 x (1 + erf(x/sqrt2))."""
    clamp, coef = _gelu_coefficients()
    assert len(coef) == 8 and 6.0 < clamp < 7.0
    x = np.linspace(-12.0, 12.0, 400001).astype(np.float32)
    u = np.minimum(np.abs(x), np.float32(clamp))
    r = np.full_like(u, np.float32(coef[0]))
    for c in coef[1:]:
        r = r * u + np.float32(c)
    e = np.exp2((r * u - np.float32(1.0)).astype(np.float64))
    got = -np.abs(x).astype(np.float64) * e + np.maximum(x, 0).astype(np.float64)
    want = np.array([0.5 * v * (1.0 + math.erf(v / math.sqrt(2.0))) for v in x.astype(np.float64)])
    assert np.abs(got - want).max() < 1e-7


def _tiles(N, bnmax=96):
    """Mirror of gemm_tc.cu: n_tiles = ceil(N / bnmax); widths are multiples of 16 that differ by at most 16."""
    n_tiles = (N + bnmax - 1) // bnmax
    units = (N + 15) // 16
    base, rem = divmod(units, n_tiles)
    n0 = [16 * (t * base + min(t, rem)) for t in range(n_tiles)]
    bn = [16 * (base + (1 if t < rem else 0)) for t in range(n_tiles)]
    return n0, bn


def test_gemm_tile_partition_covers_n_without_padding_columns():
    for N in (16, 64, 96, 100, 128, 256, 300, 512, 588, 1024, 97, 193):
        n0, bn = _tiles(N)
        assert len(bn) == packing.tc_n_tiles(N)
        assert max(bn) <= packing.tc_bn(N) <= 96 and min(bn) >= 16 and all(b % 16 == 0 for b in bn)
        assert max(bn) - min(bn) <= 16
        assert n0[0] == 0 and all(n0[i] + bn[i] == n0[i + 1] for i in range(len(bn) - 1))
        assert n0[-1] + bn[-1] == (N + 15) // 16 * 16                    # exactly N rounded up to 16: no padded tile columns
        # the TMA box is tc_bn(N) rows from n0: it must stay inside the packed weight rows (n_tiles * tc_bn)
        assert n0[-1] + packing.tc_bn(N) <= packing.tc_n_tiles(N) * packing.tc_bn(N)
