"""GPU parity of the transforms north_star names but no shipped config selects: rational-quadratic
spline coupling, exponential coupling, CIF block (Augment / Reverse / AffineCoupling / ActNorm / Slice), Permuter,
FullCombiner, ExponentialCombiner, ReLU conditioners, the identity augmenter.  Goldens are the UNMODIFIED reference's outputs
(tests/golden/a18_*.pt, oracle/make_golden.py); tolerances are north_star's: 1e-3 nats per point, 1e-4 relative on the mean."""
import math

import pytest
import torch

from flowcompare_b200 import configs, engine, lib as fclib
from oracle import port
from oracle.make_golden import A18, fixture_inputs
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "fp16x3"])
@pytest.mark.parametrize("name", A18)
def test_a18_inner_loop_matches_reference_golden(name, precision):
    gold = load_golden(name)
    cfg, fsd, esd, batch = fixture_inputs(name)
    e = engine.FlowCompareB200((fsd, esd), cfg, device="cuda:0", precision=precision)
    extra = None if batch["extra_context"] is None else batch["extra_context"].cuda()
    eps_cif = batch["eps_cif"].cuda() if "eps_cif" in batch else None
    loss, lp, bpd = e.inner_loop((batch["extract_0"].cuda(), batch["extract_1"].cuda(), extra), eps=batch["eps"].cuda(),
                                 eps_cif=eps_cif)
    d = (lp.cpu() - gold["log_prob"]).abs()
    print(f"{name} {precision}: max {d.max().item():.3e} mean {d.mean().item():.3e}")
    assert d.max().item() < 1e-3
    assert abs(loss.item() - gold["loss"].item()) / abs(gold["loss"].item()) < 1e-4
    assert abs(bpd.item() - gold["bpd"].item()) / abs(gold["bpd"].item()) < 1e-4
    e.close()


@pytest.mark.parametrize("n,nb,M", [(150, 8, 257), (12, 8, 64), (40, 16, 33), (7, 3, 5)])
def test_rq_spline_op_matches_port(n, nb, M):
    """fc_rq_spline against the port of `unconstrained_rational_quadratic_spline` (reference models/spline_coupling.py:24-169),
    forward (values + summed log|det|), inverse, tails, and the forward -> inverse round trip."""
    lib = fclib.load()
    g = torch.Generator().manual_seed(n * 100 + nb)
    x = torch.rand(M, n, generator=g) * 7.2 - 3.6
    x[0, 0], x[0, n - 1] = 3.0, -3.0                       # the interval's end points belong to the spline
    pr = torch.randn(M, n, 3 * nb + 1, generator=g) * 1.5
    want_y, want_lad = port.rq_spline(x, pr[..., :nb], pr[..., nb:2 * nb], pr[..., 2 * nb:])
    # fp64 evaluation of the same function: bins as narrow as 1e-3 of the interval make the fp32 result itself uncertain at the
    # 1e-4 level, so the kernel is held to (a small multiple of) the fp32 port's own distance from the fp64 values
    p64 = pr.double()
    true_y, true_lad = port.rq_spline(x.double(), p64[..., :nb], p64[..., nb:2 * nb], p64[..., 2 * nb:])
    port_y_err = (want_y.double() - true_y).abs().max().item()
    port_l_err = (want_lad.double().sum(-1) - true_lad.sum(-1)).abs().max().item()
    st = torch.cuda.current_stream().cuda_stream
    ldp = (n * (3 * nb + 1) + 3) // 4 * 4
    P = torch.zeros(M, ldp, device="cuda"); P[:, :n * (3 * nb + 1)] = pr.reshape(M, -1).cuda()
    X = x.cuda().contiguous()
    ldj = torch.zeros(M, device="cuda")
    assert lib.fc_rq_spline(P.data_ptr(), ldp, X.data_ptr(), n, n, nb, M, ldj.data_ptr(), 0, st) == 0
    y_err = (X.cpu().double() - true_y).abs().max().item()
    l_err = (ldj.cpu().double() - true_lad.sum(-1)).abs().max().item()
    print(f"spline n={n} nb={nb}: |y - fp64| {y_err:.3e} (port {port_y_err:.3e})  |ldj - fp64| {l_err:.3e} (port {port_l_err:.3e})")
    assert y_err < max(2e-5, 3 * port_y_err)
    assert l_err < max(1e-4, 3 * port_l_err)
    outside = x.abs() > 3
    assert torch.equal(X.cpu()[outside], x[outside])
    assert lib.fc_rq_spline(P.data_ptr(), ldp, X.data_ptr(), n, n, nb, M, 0, 1, st) == 0
    want_x, _ = port.rq_spline(want_y, pr[..., :nb], pr[..., nb:2 * nb], pr[..., 2 * nb:], inverse=True)
    print(f"   inverse: |x - x0| {(X.cpu() - x).abs().max().item():.3e}   port: {(want_x - x).abs().max().item():.3e}")
    assert (X.cpu() - x).abs().max().item() < max(1e-4, 3 * (want_x - x).abs().max().item())   # forward -> inverse round trip


@pytest.mark.parametrize("n,M,amp", [(150, 40, 0.05), (150, 8, 0.6), (12, 100, 1.0), (3, 7, 2.0), (200, 6, 0.05), (64, 16, 0.3)])
def test_expm_action_op_matches_matrix_exp(n, M, amp):
    """fc_expm_action against expm(W) x + b with torch.matrix_exp in float64 (reference models/exponential_coupling.py:48-58:
    the reference forms the matrix exponential; only its action on x2 is computed here), trace, and the inverse."""
    lib = fclib.load()
    g = torch.Generator().manual_seed(n + M)
    w = torch.randn(M, n * n, generator=g) * amp
    b = torch.randn(M, n, generator=g)
    x = torch.randn(M, n, generator=g)
    sq = torch.tensor([0.13, 0.01, 1.02, -0.002])
    W = (sq[2].double() * torch.tanh(sq[0].double() * w.double() + sq[1].double()) + sq[3].double() + 1e-8).reshape(M, n, n)
    want = torch.matmul(torch.matrix_exp(W), x.double().unsqueeze(-1)).squeeze(-1) + b.double()
    ldp = (n * n + n + 3) // 4 * 4
    P = torch.zeros(M, ldp, device="cuda"); P[:, :n * n] = w.cuda(); P[:, n * n:n * n + n] = b.cuda()
    X = x.cuda().contiguous()
    tr = torch.zeros(M, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    assert lib.fc_expm_action(P.data_ptr(), ldp, X.data_ptr(), n, n, sq.cuda().data_ptr(), tr.data_ptr(), M, 0, st) == 0
    scale = want.abs().max().item()
    assert (X.cpu().double() - want).abs().max().item() < 2e-5 * max(1.0, scale)
    assert (tr.cpu().double() - W.diagonal(dim1=-2, dim2=-1).sum(-1)).abs().max().item() < 1e-4
    assert lib.fc_expm_action(P.data_ptr(), ldp, X.data_ptr(), n, n, sq.cuda().data_ptr(), 0, M, 1, st) == 0
    assert (X.cpu() - x).abs().max().item() < 1e-3 * max(1.0, scale)


def test_cif_flow_requires_its_noise():
    """fc_flow_log_prob on a CIF flow (no eps_cif) is an argument error, not a silent zero-noise pass."""
    cfg, fsd, esd, batch = fixture_inputs("a18_cif")
    e = engine.FlowCompareB200((fsd, esd), cfg, device="cuda:0", precision="fp32")
    ctx = e.embed(batch["extract_0"].cuda())
    x = batch["extract_1"].cuda().contiguous()
    B, N = x.shape[:2]
    out = torch.empty(B, N, device="cuda")
    nbytes = e.lib.fc_flow_workspace_bytes(e._flow["handle"], B, N, ctx.shape[1])
    ws = e._workspace(nbytes)
    rc = e.lib.fc_flow_log_prob(e._flow["handle"], x.data_ptr(), ctx.data_ptr(), 0, batch["eps"].cuda().data_ptr(), out.data_ptr(),
                                B, N, ctx.shape[1], ws, nbytes, 0, torch.cuda.current_stream().cuda_stream)
    assert rc == -1   # FC_ERR_INVALID_ARG
    assert e.lib.fc_flow_cif_noise_dim(e._flow["handle"]) == 8
    e.close()


def _guarded(rows, ld, guard=2048):
    whole = torch.full((guard + rows * ld + guard,), -12345.0, device="cuda")
    return whole, whole[guard:guard + rows * ld].view(rows, ld), guard


@pytest.mark.parametrize("n,M", [(150, 131), (7, 5)])
def test_transform_ops_write_only_their_outputs(n, M):
    """Guard bands around every buffer fc_rq_spline / fc_expm_action write (compute-sanitizer is not available on the GPU
    pool): nothing outside the n written columns of the M rows, the row-sum vectors and nothing of the parameters changes."""
    lib = fclib.load()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(n)
    nb, ldx = 8, n + 3
    ldp = (n * (3 * nb + 1) + 3) // 4 * 4
    wp, P, gp = _guarded(M, ldp)
    P.copy_(torch.randn(M, ldp, generator=g).cuda())
    wx, X, gx = _guarded(M, ldx)
    X.copy_((torch.rand(M, ldx, generator=g) * 6 - 3).cuda())
    wl, L, gl = _guarded(1, M)
    L.zero_()
    p0, x0 = wp.clone(), wx.clone()
    assert lib.fc_rq_spline(P.data_ptr(), ldp, X.data_ptr(), ldx, n, nb, M, L.data_ptr(), 0, st) == 0
    torch.cuda.synchronize()
    assert torch.equal(wp, p0)
    assert torch.equal(wx[:gx], x0[:gx]) and torch.equal(wx[gx + M * ldx:], x0[gx + M * ldx:])
    assert torch.equal(X[:, n:], x0[gx:gx + M * ldx].view(M, ldx)[:, n:])          # the row padding
    assert (wl[:gl] == -12345.0).all() and (wl[gl + M:] == -12345.0).all() and torch.isfinite(L).all()
    # exponential coupling
    ldp2 = (n * n + n + 3) // 4 * 4
    wp2, P2, _ = _guarded(M, ldp2)
    P2.copy_((torch.randn(M, ldp2, generator=g) * 0.2).cuda())
    wx2, X2, gx2 = _guarded(M, ldx)
    X2.copy_(torch.randn(M, ldx, generator=g).cuda())
    wt, T, gt = _guarded(1, M)
    T.zero_()
    sq = torch.tensor([0.125, 0.0, 1.0, 0.0], device="cuda")
    p20, x20 = wp2.clone(), wx2.clone()
    assert lib.fc_expm_action(P2.data_ptr(), ldp2, X2.data_ptr(), ldx, n, sq.data_ptr(), T.data_ptr(), M, 0, st) == 0
    torch.cuda.synchronize()
    assert torch.equal(wp2, p20)
    assert torch.equal(wx2[:gx2], x20[:gx2]) and torch.equal(wx2[gx2 + M * ldx:], x20[gx2 + M * ldx:])
    assert torch.equal(X2[:, n:], x20[gx2:gx2 + M * ldx].view(M, ldx)[:, n:])
    assert (wt[:gt] == -12345.0).all() and (wt[gt + M:] == -12345.0).all() and torch.isfinite(X2[:, :n]).all()
