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


def test_dataops_write_only_their_outputs():
    """Guard bands around the outputs of fc_fps_points and fc_co_unit_sphere (no sanitizer on the GPU pool)."""
    from flowcompare_b200 import lib as fclib
    lib = fclib.load()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(1)
    B, n, C, m, G = 3, 777, 6, 100, 1024
    pts = torch.rand(B, n, C, generator=g).cuda()
    widx = torch.full((G + B * m + G,), -7, dtype=torch.int32, device="cuda")
    wout = torch.full((G + B * m * C + G,), -12345.0, device="cuda")
    p0 = pts.clone()
    assert lib.fc_fps_points(pts.data_ptr(), C, B, n, C, m, widx[G:].data_ptr(), wout[G:].data_ptr(), C, st) == 0
    torch.cuda.synchronize()
    assert torch.equal(pts, p0)
    assert (widx[:G] == -7).all() and (widx[G + B * m:] == -7).all() and int(widx[G:G + B * m].min()) >= 0 and int(widx[G:G + B * m].max()) < n
    assert (wout[:G] == -12345.0).all() and (wout[G + B * m * C:] == -12345.0).all()
    n0, n1 = 300, 200
    wa = torch.full((G + B * n0 * C + G,), -12345.0, device="cuda"); a = wa[G:G + B * n0 * C].view(B, n0, C); a.copy_(torch.rand(B, n0, C, generator=g).cuda())
    wb = torch.full((G + B * n1 * C + G,), -12345.0, device="cuda"); b = wb[G:G + B * n1 * C].view(B, n1, C); b.copy_(torch.rand(B, n1, C, generator=g).cuda())
    winv = torch.full((G + B * 4 + G,), -12345.0, device="cuda")
    rgb_a, rgb_b = a[..., 3:].clone(), b[..., 3:].clone()
    assert lib.fc_co_unit_sphere(a.data_ptr(), n0, C, b.data_ptr(), n1, C, B, winv[G:].data_ptr(), st) == 0
    torch.cuda.synchronize()
    for w, cnt in ((wa, B * n0 * C), (wb, B * n1 * C), (winv, B * 4)):
        assert (w[:G] == -12345.0).all() and (w[G + cnt:] == -12345.0).all()
    assert torch.equal(a[..., 3:], rgb_a) and torch.equal(b[..., 3:], rgb_b)
