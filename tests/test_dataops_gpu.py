"""GPU parity of the data-side ops: furthest point sub-sampling and the joint unit-ball normalisation."""
import numpy as np
import pytest
import torch

from flowcompare_b200 import dataops
from oracle import dataops_ref
from oracle.make_dataops_golden import inputs
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,C,m", [(1250, 6, 1024), (5000, 6, 1250), (700, 3, 64), (33, 6, 33), (20000, 6, 1250)])
def test_fps_subsample_matches_oracle_bit_exact(n, C, m):
    """Indices bit-exact against the restatement of torch_cluster.fps (oracle/dataops_ref.py), incl. duplicated points."""
    g = torch.Generator().manual_seed(n + m)
    pts = torch.rand(n, C, generator=g)
    pts[n // 2:n // 2 + 5] = pts[:5]                       # exact duplicates: ties in the arg-max
    want = dataops_ref.fps(pts.numpy(), m)
    got, idx = dataops.fps_subsample(pts.cuda(), m, return_index=True)
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), want)
    assert torch.equal(got.cpu(), pts[torch.from_numpy(want)])


def test_fps_subsample_batched_equals_single():
    g = torch.Generator().manual_seed(3)
    pts = torch.rand(4, 900, 6, generator=g).cuda()
    out, idx = dataops.fps_subsample(pts, 256, return_index=True)
    for b in range(4):
        o1, i1 = dataops.fps_subsample(pts[b], 256, return_index=True)
        assert torch.equal(idx[b], i1) and torch.equal(out[b], o1)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_co_unit_sphere_matches_reference_golden(seed):
    """Against the UNMODIFIED reference's utils.co_unit_sphere (tests/golden/dataops.pt): 1e-6 absolute on unit-ball coordinates
    (the joint mean is accumulated in another order), colours untouched, inverse = (mean, furthest distance)."""
    gold = load_golden("dataops")[seed]
    p0, p1 = inputs(seed)
    a, b, inv = dataops.co_unit_sphere(p0.cuda(), p1.cuda(), return_inverse=True)
    assert (a.cpu() - gold["points_0"]).abs().max().item() < 1e-6
    assert (b.cpu() - gold["points_1"]).abs().max().item() < 1e-6
    assert torch.equal(a.cpu()[:, 3:], p0[:, 3:]) and torch.equal(b.cpu()[:, 3:], p1[:, 3:])
    assert abs(inv["furthest_distance"].item() - gold["furthest_distance"].item()) < 1e-6 * gold["furthest_distance"].item() + 1e-7
    assert (inv["mean"].cpu() - gold["mean"]).abs().max().item() < 1e-5
    joint = torch.cat((a, b))[:, :3]
    assert abs(joint.norm(dim=-1).max().item() - 1.0) < 1e-6 and joint.mean(0).abs().max().item() < 1e-6


def test_co_unit_sphere_batched_equals_single():
    g = torch.Generator().manual_seed(9)
    p0, p1 = torch.rand(3, 500, 6, generator=g).cuda() * 5, torch.rand(3, 400, 6, generator=g).cuda() * 5
    a, b = dataops.co_unit_sphere(p0, p1)
    for i in range(3):
        ai, bi = dataops.co_unit_sphere(p0[i], p1[i])
        assert torch.equal(a[i], ai) and torch.equal(b[i], bi)
