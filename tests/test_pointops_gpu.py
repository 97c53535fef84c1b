"""GPU tests of the op-level pointops entry points (fc_fps / fc_knn_heap / fc_three_nn / fc_three_interpolate /
fc_group_points, through flowcompare_b200.pointops -> ctypes -> C ABI) against

  * the REFERENCE's own compiled kernels (oracle/_ref/libpointops_ref.so: the reference's unmodified
    lib/pointops/src/*/*_cuda_kernel.cu built by oracle/build_ref_pointops.sh) -- bit-exact indices, and
  * the CPU restatement oracle/pointops_ref.c (which the PAConv goldens were generated with), so that restatement is
    pinned to the compiled reference too.

/root/reference is not needed at run time: the .so travels with the repo snapshot.  If it is absent the reference
comparisons are skipped and only the restatement is checked.
"""
import time

import pytest
import torch

from flowcompare_b200 import configs, pointops as fpo, spec
from oracle import pointops_refcuda as ref
from oracle import port_paconv

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libpointops_ref.so not built (needs /root/reference once)")


def _cloud(B, n, seed, duplicates=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, n, 3, generator=g) * 2 - 1
    if duplicates:   # the reference's loader oversamples small voxels: exact duplicate points (utils.py:362-370)
        x[:, n // 2:] = x[:, : n - n // 2]
    return x.contiguous()


FPS_CASES = [(3, 1250, 312, False), (2, 312, 78, False), (2, 78, 19, False), (2, 19, 4, False), (1, 4096, 1024, False),
             (2, 5000, 1250, False), (2, 1250, 312, True), (1, 2500, 700, True), (1, 33, 33, False)]


@needs_ref
@pytest.mark.parametrize("B,n,m,dup", FPS_CASES)
def test_fps_bit_exact_vs_reference_kernel(B, n, m, dup):
    x = _cloud(B, n, n + m, dup).to(DEV)
    want = ref.furthestsampling(x, m)
    got, nx = fpo.furthestsampling(x, m, return_xyz=True)
    assert torch.equal(got, want)
    assert torch.equal(nx, torch.gather(x, 1, got.long().unsqueeze(-1).expand(-1, -1, 3)))


@pytest.mark.parametrize("B,n,m,dup", FPS_CASES[:4] + FPS_CASES[6:7])
def test_fps_matches_cpu_restatement(B, n, m, dup):
    x = _cloud(B, n, n + m, dup)
    assert torch.equal(fpo.furthestsampling(x.to(DEV), m).cpu(), port_paconv.furthestsampling(x, m))


def test_fps_lowest_index_tie_rule():
    """tie_block < 0: greedy farthest point where equal distances (exact duplicates) go to the LOWEST index."""
    import numpy as np
    x = _cloud(1, 2500, 5, duplicates=True)
    got = fpo.furthestsampling(x.to(DEV), 600, tie_block=-1).cpu()[0].tolist()
    xc = x[0].numpy()
    d = np.full(2500, 1e10, dtype=np.float32)
    old, want = 0, [0]
    for _ in range(599):
        dx = (xc - xc[old]).astype(np.float32)
        # fmaf(dz, dz, fmaf(dy, dy, dx * dx)) emulated through float64 (products of two floats are exact there)
        t = (dx[:, 0] * dx[:, 0]).astype(np.float32)
        t = (dx[:, 1].astype(np.float64) * dx[:, 1] + t).astype(np.float32)
        t = (dx[:, 2].astype(np.float64) * dx[:, 2] + t).astype(np.float32)
        d = np.minimum(d, t)
        old = int(np.argmax(d))          # first maximum = lowest index
        want.append(old)
    assert got == want


KNN_CASES = [(2, 1250, 312, 32, False), (2, 312, 78, 32, False), (2, 78, 19, 32, False), (2, 19, 4, 32, False),
             (2, 1250, 312, 32, True), (1, 700, 700, 16, False), (1, 3000, 129, 40, False)]


@needs_ref
@pytest.mark.parametrize("B,n,m,k,dup", KNN_CASES)
def test_knn_heap_bit_exact_vs_reference_kernel(B, n, m, k, dup):
    x = _cloud(B, n, 3 * n + k, dup).to(DEV)
    q = x[:, torch.randperm(n, generator=torch.Generator().manual_seed(n))[:m]].contiguous()
    want_idx, want_d2 = ref.knnquery_heap(k, x, q)
    got_idx, got_d2 = fpo.knnquery_heap(k, x, q, return_dist2=True)
    assert torch.equal(got_idx, want_idx)      # including the heap's order of exactly equal distances and the k > n padding
    assert torch.equal(got_d2, want_d2)


@pytest.mark.parametrize("B,n,m,k,dup", KNN_CASES[:5])
def test_knn_heap_matches_cpu_restatement(B, n, m, k, dup):
    x = _cloud(B, n, 3 * n + k, dup)
    q = x[:, torch.randperm(n, generator=torch.Generator().manual_seed(n))[:m]].contiguous()
    assert torch.equal(fpo.knnquery_heap(k, x.to(DEV), q.to(DEV)).cpu(), port_paconv.knnquery_heap(k, x, q))


NN_CASES = [(2, 1250, 312, False), (2, 312, 78, False), (2, 19, 4, False), (1, 7, 2, False), (2, 1250, 312, True), (1, 5000, 1250, False)]


@needs_ref
@pytest.mark.parametrize("B,n,m,dup", NN_CASES)
def test_three_nn_bit_exact_vs_reference_kernel(B, n, m, dup):
    u = _cloud(B, n, n + 7 * m, dup).to(DEV)
    kn = u[:, :m].contiguous() if dup else _cloud(B, m, m, False).to(DEV)
    want_d2, want_idx = ref.nearestneighbor(u, kn)
    got_d2, got_idx = fpo.nearestneighbor(u, kn, squared=True)
    assert torch.equal(got_idx, want_idx)
    assert torch.equal(got_d2, want_d2)


@pytest.mark.parametrize("B,n,m,dup", NN_CASES[:3])
def test_three_nn_matches_cpu_restatement(B, n, m, dup):
    u, kn = _cloud(B, n, n + 7 * m), _cloud(B, m, m)
    d, i = port_paconv.nearestneighbor(u, kn)
    gd2, gi = fpo.nearestneighbor(u.to(DEV), kn.to(DEV), squared=True)
    assert torch.equal(gi.cpu(), i) and torch.equal(torch.sqrt(gd2.cpu()), d)    # same sqrt implementation on both sides


@needs_ref
def test_interpolation_and_grouping_vs_reference_kernels():
    g = torch.Generator().manual_seed(3)
    B, c, m, n, k = 2, 67, 312, 1250, 32
    feats = torch.randn(B, c, m, generator=g).to(DEV)
    u, kn = _cloud(B, n, 1).to(DEV), _cloud(B, m, 2).to(DEV)
    d, idx = fpo.nearestneighbor(u, kn)
    w = 1.0 / (d + 1e-8)
    w = (w / w.sum(dim=2, keepdim=True)).contiguous()
    assert torch.equal(fpo.interpolation(feats, idx, w), ref.interpolation(feats, idx, w))
    gidx = torch.randint(0, m, (B, 100, k), generator=g, dtype=torch.int32).to(DEV)
    assert torch.equal(fpo.grouping(feats, gidx), ref.grouping(feats, gidx))
    fidx = torch.randint(0, m, (B, 50), generator=g, dtype=torch.int32).to(DEV)
    assert torch.equal(fpo.gathering(feats, fidx), ref.gathering(feats, fidx))


@needs_ref
def test_pointops_timing_vs_reference_kernels(capsys):
    """Not an assertion on speed -- prints the device time of each op next to the reference's own kernel compiled for
    sm_100a, at the PAConv embedder's level-0 shapes and 128 clouds (the bench batch).  scripts/pointops_bench.py writes the
    same table to profiles/."""
    from scripts.pointops_bench import table
    rows = table(B=128)
    with capsys.disabled():
        for r in rows:
            print("  ", r)
    assert all(r["match"] for r in rows)
