"""CPU tests: the oracle port against the reference's golden outputs (and against the live reference
when /root/reference is present), the kNN C oracle, and the pack-time algebra (arena interpreter)."""
import os

import pytest
import torch

from flowcompare_b200 import configs, packing, spec
from oracle import knn_ref, port, refload
from oracle.make_golden import FIXTURES, fixture_inputs
from tests import arena_sim
from tests.conftest import load_golden, voxel_label_mismatch

torch.set_grad_enabled(False)

TINY = ["tiny_dgcnn_attn", "tiny_dgcnn_attn_extra", "tiny_dgcnn_global"]


def _port_forward(name):
    cfg, fsd, esd, batch = fixture_inputs(name)
    dcfg = configs.derive(cfg)
    loss, lp, bpd = port.inner_loop((batch["extract_0"], batch["extract_1"], batch["extra_context"]), fsd, esd, dcfg,
                                    batch["eps"])
    return cfg, fsd, esd, batch, loss, lp, bpd


@pytest.mark.parametrize("name", TINY + ["mid_dgcnn_attn"])
def test_port_matches_reference_golden(name):
    """tolerance: north_star's 1e-3 nats absolute per point, 1e-4 relative on the mean."""
    gold = load_golden(name)
    cfg, fsd, esd, batch, loss, lp, bpd = _port_forward(name)
    assert lp.shape == gold["log_prob"].shape
    assert (lp - gold["log_prob"]).abs().max().item() < 1e-3
    assert abs(loss.item() - gold["loss"].item()) / abs(gold["loss"].item()) < 1e-4
    assert abs(bpd.item() - gold["bpd"].item()) / abs(gold["bpd"].item()) < 1e-4


def test_port_embedding_matches_golden():
    gold = load_golden("tiny_dgcnn_attn")
    cfg, fsd, esd, batch = fixture_inputs("tiny_dgcnn_attn")
    emb, idxs = port.dgcnn_embed(esd, batch["extract_0"], cfg["n_neighbors"])
    assert (emb - gold["embedding"]).abs().max().item() < 1e-5
    assert torch.equal(idxs[0].to(torch.int32), gold["knn_idx_layer1"])


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_state_dict_layout_matches_reference():
    models, mi = refload.load()
    for label in ["dgcnn_attn", "dgcnn_attn_extra", "dgcnn_global"]:
        cfg = configs.get_config(label, n_flow_layers=2)
        md = mi.initialize_flow(dict(cfg), "cpu", "test")
        fsd, esd = spec.random_state_dicts(cfg, seed=0)
        rf, re_ = md["flow"].state_dict(), md["input_embedder"].state_dict()
        assert list(rf.keys()) == list(fsd.keys())
        assert list(re_.keys()) == list(esd.keys())
        assert all(tuple(rf[k].shape) == tuple(fsd[k].shape) for k in rf)
        assert all(tuple(re_[k].shape) == tuple(esd[k].shape) for k in re_)


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_port_matches_live_reference_with_reference_init():
    """Reference's own (unperturbed) init, live forward, eps injected."""
    from oracle.make_golden import run_reference
    cfg = configs.get_config("dgcnn_attn_extra", n_flow_layers=2, sample_size=64, n_samples_context=80)
    models, mi = refload.load()
    torch.manual_seed(5)
    md = mi.initialize_flow(dict(cfg), "cpu", "test")
    fsd, esd = md["flow"].state_dict(), md["input_embedder"].state_dict()
    batch = spec.synthetic_batch(cfg, 2, seed=9)
    out = run_reference(cfg, fsd, esd, batch)
    _, lp, _ = port.inner_loop((batch["extract_0"], batch["extract_1"], batch["extra_context"]), fsd, esd,
                               configs.derive(cfg), batch["eps"])
    assert (lp - out["log_prob"]).abs().max().item() < 1e-3


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("label", ["dgcnn_attn_extra", "dgcnn_global"])
def test_port_sampling_pass_matches_live_reference(label):
    """SURVEY 8f rank 1 (oracle side only so far): `Flow.sample` / `make_sample` (reference models/transform.py:79-84,
    model_initialization.py:231-245) against port.flow_sample, with the base draw injected; and the round trip
    inverse(forward(x)) = x of the port."""
    import einops
    cfg = configs.get_config(label, n_flow_layers=3, sample_size=64, n_samples_context=80)
    models, mi = refload.load()
    fsd, esd = spec.random_state_dicts(cfg, seed=3)
    torch.manual_seed(0)
    md = mi.initialize_flow(dict(cfg), "cpu", "test")
    md["flow"].load_state_dict(fsd)
    md["input_embedder"].load_state_dict(esd)
    dcfg = configs.derive(cfg)
    batch = spec.synthetic_batch(cfg, 1, seed=4)
    n_points = 50
    z0 = torch.randn(1, n_points, cfg["latent_dim"], generator=torch.Generator().manual_seed(8))

    class Injected(torch.nn.Module):        # stands in for sample_distrib: returns the injected base draw
        def sample(self, num_samples, context=None, n_points=None):
            return z0.clone()

    with torch.no_grad():
        extra = batch["extra_context"]
        want = mi.make_sample(n_points, batch["extract_0"], md, dcfg, sample_distrib=Injected(), extra_context=extra)
        if dcfg["global"]:
            emb, _ = port.dgcnn_embed_global(esd, batch["extract_0"][:, :, :6], cfg["n_neighbors"])
            ctx = einops.repeat(emb, "b e -> b p e", p=n_points)
        else:
            ctx, _ = port.dgcnn_embed(esd, batch["extract_0"][:, :, :6], cfg["n_neighbors"])
        ex = None if extra is None else einops.repeat(extra, "b c -> b n c", n=n_points)
        got = port.flow_sample(fsd, dcfg, z0, ctx, ex)
        assert got.shape == (1, n_points, 6)
        assert (got.squeeze() - want).abs().max().item() < 1e-3
        # round trip through the port: sample -> forward reproduces the latent's first columns' pre-image
        x = got
        eps = torch.randn(1, n_points, cfg["latent_dim"] - 6, generator=torch.Generator().manual_seed(9))
        trace = []
        lp = port.flow_log_prob(fsd, dcfg, x, ctx, ex, eps, trace=trace)
        assert torch.isfinite(lp).all()
        x_back = port.flow_sample(fsd, dcfg, trace[-1][1], ctx, ex)     # inverse(forward(x)) == x
        assert (x_back - x).abs().max().item() < 1e-4


def test_inverse_actnorm_lu_fold_matches_port():
    """packing.fold_inverse_actnorm_lu (pack-time algebra of the sampling pass) against the port's two inverse steps."""
    cfg = configs.get_config("dgcnn_attn", n_flow_layers=4)
    fsd, _ = spec.random_state_dicts(cfg, seed=6)
    folds = packing.fold_inverse_actnorm_lu(fsd, cfg)
    assert len(folds) == 3
    sd64 = port.to_dtype(fsd, torch.float64)
    g = torch.Generator().manual_seed(1)
    zp = torch.randn(2, 17, cfg["latent_dim"], generator=g, dtype=torch.float64)
    t = 1
    for layer, (w_off, wdiag, bias) in enumerate(folds):
        t += 1                                  # the coupling layer
        want = port.actnorm_inverse(sd64, t, port.linear_lu_inverse(sd64, t + 1, zp, cfg["linear_lu_eps"]))
        got = zp * wdiag + zp @ w_off.t() + bias
        assert (got - want).abs().max().item() < 1e-9
        # and it undoes the forward pair
        fwd, _ = port.linear_lu_forward(sd64, t + 1, port.actnorm_forward(sd64, t, got)[0], cfg["linear_lu_eps"])
        assert (fwd - zp).abs().max().item() < 1e-9
        t += 2


# ----------------------------------------------------------------------------- kNN oracle
def _tie_safe_mismatches(x, idx_a, idx_b, k):
    """Rows where the neighbour SETS differ must be rounding ties: the fp64 distances of the symmetric
    difference are equal to ~1e-6 relative."""
    bad = 0
    xd = x.double()
    for b in range(x.shape[0]):
        d = torch.cdist(xd[b], xd[b]) ** 2
        for i in range(x.shape[1]):
            sa, sb = set(idx_a[b, i].tolist()), set(idx_b[b, i].tolist())
            if sa != sb:
                da = sorted(d[i, list(sa - sb)].tolist())
                db = sorted(d[i, list(sb - sa)].tolist())
                if any(abs(p - q) > 1e-5 * max(1.0, abs(p)) for p, q in zip(da, db)):
                    bad += 1
    return bad


@pytest.mark.parametrize("C,N", [(6, 300), (64, 257), (128, 200)])
def test_knn_c_oracle_vs_reference_form(C, N):
    g = torch.Generator().manual_seed(C)
    x = torch.randn(2, N, C, generator=g)
    k = 40
    a = knn_ref.knn_self(x, k)
    b = port.knn_reference_form(x.permute(0, 2, 1), k)
    assert (a == b).float().mean().item() > 0.999
    assert _tie_safe_mismatches(x, a, b, k) == 0


def test_knn_c_oracle_ties_lower_index_first():
    x = torch.rand(1, 64, 6)
    x[0, 32:] = x[0, :32]  # exact duplicates
    idx = knn_ref.knn_self(x, 4)
    for i in range(32):
        assert idx[0, i, 0].item() == i and idx[0, i, 1].item() == i + 32
        assert idx[0, i + 32, 0].item() == i and idx[0, i + 32, 1].item() == i + 32


def test_knn_query_oracle_vs_reference_form():
    g = torch.Generator().manual_seed(3)
    q, t = torch.randn(100, 3, generator=g), torch.randn(500, 3, generator=g)
    diss = (q ** 2).sum(-1).view(-1, 1) + (t ** 2).sum(-1).view(1, -1) - 2 * q @ t.t()   # knn.py:44-49
    ref = diss.topk(8, dim=1, largest=False).indices
    got = knn_ref.knn_query(q, t, 8)
    assert (got == ref).float().mean().item() > 0.995


# ----------------------------------------------------------------------------- packing algebra
@pytest.mark.parametrize("name", TINY)
def test_packed_flow_algebra_matches_port(name):
    cfg, fsd, esd, batch = fixture_inputs(name)
    dcfg = configs.derive(cfg)
    packed = packing.pack_flow(fsd, cfg)
    if dcfg["global"]:
        ctx, _ = port.dgcnn_embed_global(esd, batch["extract_0"], cfg["n_neighbors"])
        ctx_port = ctx.unsqueeze(1).expand(-1, batch["extract_1"].shape[1], -1)
    else:
        ctx, _ = port.dgcnn_embed(esd, batch["extract_0"], cfg["n_neighbors"])
        ctx_port = ctx
    extra = batch["extra_context"]
    ex_port = None if extra is None else extra.unsqueeze(1).expand(-1, batch["extract_1"].shape[1], -1)
    want = port.flow_log_prob(fsd, dcfg, batch["extract_1"], ctx_port, ex_port, batch["eps"])
    got = arena_sim.flow_log_prob(packed, batch["extract_1"], ctx, extra, batch["eps"])
    assert (got - want).abs().max().item() < 1e-3


def test_packed_tf32x3_copies_match_port():
    """The TF32 hi/lo weight copies + a 3xTF32 evaluation of every GEMM (what csrc/gemm_tc.cu computes)
    stay within north_star's 1e-3 nats of the fp32 oracle."""
    name = "tiny_dgcnn_attn_extra"
    cfg, fsd, esd, batch = fixture_inputs(name)
    dcfg = configs.derive(cfg)
    packed = packing.pack_flow(fsd, cfg)
    ctx, _ = port.dgcnn_embed(esd, batch["extract_0"], cfg["n_neighbors"])
    extra = batch["extra_context"]
    ex_port = extra.unsqueeze(1).expand(-1, batch["extract_1"].shape[1], -1)
    want = port.flow_log_prob(fsd, dcfg, batch["extract_1"], ctx, ex_port, batch["eps"])
    arena_sim.USE_TC = True
    try:
        got = arena_sim.flow_log_prob(packed, batch["extract_1"], ctx, extra, batch["eps"])
    finally:
        arena_sim.USE_TC = False
    assert (got - want).abs().max().item() < 1e-3


@pytest.mark.parametrize("name", ["tiny_dgcnn_attn", "tiny_dgcnn_global"])
def test_packed_embedder_algebra_matches_port(name):
    cfg, fsd, esd, batch = fixture_inputs(name)
    packed = packing.pack_embedder(esd, cfg)
    got, idxs = arena_sim.dgcnn_embed(packed, batch["extract_0"], knn_ref.knn_self)
    if configs.derive(cfg)["global"]:
        want, _ = port.dgcnn_embed_global(esd, batch["extract_0"], cfg["n_neighbors"], idx_list=idxs)
    else:
        want, _ = port.dgcnn_embed(esd, batch["extract_0"], cfg["n_neighbors"], idx_list=idxs)
    assert (got - want).abs().max().item() < 2e-5


def test_change_score_port_properties():
    g = torch.Generator().manual_seed(0)
    lp10 = torch.randn(3, 200, generator=g) * 5 - 20
    lp00 = torch.randn(3, 200, generator=g) - 10
    lp10[0, 3] = float("-inf")
    ch = port.log_prob_to_change(lp10, lp00, 1.0)
    assert ch.min() >= 0 and ch.max() <= 1 and torch.isfinite(ch).all()


# ----------------------------------------------------------------------------- change score
def test_port_change_score_matches_reference_golden():
    """oracle/port.py: log_prob_to_change against the UNMODIFIED reference's outputs (tests/golden/change_score.pt, made by
    oracle/make_change_golden.py from test_flow.py:241-275), incl. -inf clamping over the whole tensor and both masks."""
    gold = load_golden("change_score")
    assert len(gold["cases"]) == 9
    for c in gold["cases"]:
        got = port.log_prob_to_change(c["lp10"].clone(), c["lp00"].clone(), c["multiple"], c["hard_cutoff"])
        assert torch.equal(got == 0, c["change"] == 0)
        assert (got - c["change"]).abs().max().item() < 1e-6


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_port_change_score_matches_live_reference():
    tf = refload.load_test_flow()
    g = torch.Generator().manual_seed(77)
    lp10 = torch.randn(5, 300, generator=g) * 4 - 20
    lp00 = torch.randn(5, 300, generator=g) - 11
    lp10[2, 17] = float("-inf")
    lp00[0, 3] = float("-inf")
    for multiple, cut in ((5.4, None), (2.0, -21.0)):
        want = tf.log_prob_to_change(lp10.clone(), lp00.clone(), multiple, cut)
        got = port.log_prob_to_change(lp10.clone(), lp00.clone(), multiple, cut)
        assert torch.equal(got == 0, want == 0) and (got - want).abs().max().item() < 1e-6
    # the reference clamps its ARGUMENTS in place (clamp_infs mutates); the port (and fc_change_score) do not
    a = lp10.clone()
    tf.log_prob_to_change(a, lp00.clone(), 5.4)
    assert not torch.isinf(a).any() and torch.isinf(lp10).any()


# ----------------------------------------------------------------------------- sampling pass goldens
@pytest.mark.parametrize("name", ["tiny_dgcnn_attn", "tiny_dgcnn_attn_extra", "tiny_dgcnn_global", "a18_spline", "a18_spline_extra300",
                                  "a18_expo", "a18_expo_global300", "a18_cif", "a18_cif_spline_clamp", "a18_permute_relu",
                                  "a18_fullcombiner_noactnorm", "a18_expcombiner_global", "a18_identity_augmenter"])
def test_port_sampling_pass_matches_reference_golden(name):
    """oracle/port.py: flow_sample (every coupling's, permuter's and the CIF block's `.inverse`) against the UNMODIFIED
    reference's make_sample outputs (tests/golden/sample_*.pt)."""
    from oracle.make_sample_golden import base_draw, cif_draws
    cfg, fsd, esd, batch = fixture_inputs(name)
    dcfg = configs.derive(cfg)
    gold = load_golden("sample_" + name)
    P = gold["n_points"]
    z = base_draw(name, cfg, batch["extract_0"].shape[0])
    if dcfg["global"]:
        emb, _ = port.dgcnn_embed_global(esd, batch["extract_0"], cfg["n_neighbors"])
        ctx = emb.unsqueeze(1).expand(-1, P, -1)
    else:
        ctx, _ = port.dgcnn_embed(esd, batch["extract_0"], cfg["n_neighbors"])
    extra = batch["extra_context"]
    ex = None if extra is None else extra.unsqueeze(1).expand(-1, P, -1)
    got = port.flow_sample(fsd, dcfg, z, ctx, ex, eps_cif=cif_draws(name, cfg, batch["extract_0"].shape[0]))
    assert (got.squeeze() - gold["x"]).abs().max().item() < 1e-4


# --------------------------------------------------------------------------- transforms no shipped config selects
from oracle.make_golden import A18  # noqa: E402


def _a18_port_inputs(name):
    cfg, fsd, esd, batch = fixture_inputs(name)
    dcfg = configs.derive(cfg)
    B, N = batch["extract_1"].shape[:2]
    if dcfg["global"]:
        g, _ = port.dgcnn_embed_global(esd, batch["extract_0"], cfg["n_neighbors"])
        ctx_port, ctx = g.unsqueeze(1).expand(B, N, g.shape[1]), g
    else:
        ctx, _ = port.dgcnn_embed(esd, batch["extract_0"], cfg["n_neighbors"])
        ctx_port = ctx
    extra = batch["extra_context"]
    ex_port = None if extra is None else extra.unsqueeze(1).expand(B, N, 1)
    return cfg, dcfg, fsd, batch, ctx_port, ctx, ex_port, extra


@pytest.mark.parametrize("name", A18)
def test_port_a18_transforms_match_reference_golden(name):
    """oracle/port.py: rq_spline / spline and exponential couplings / CIF block / Permuter, FullCombiner, ExponentialCombiner /
    ReLU conditioners / identity augmenter against the UNMODIFIED reference's outputs (tests/golden/a18_*.pt)."""
    gold = load_golden(name)
    cfg, dcfg, fsd, batch, ctx_port, _, ex_port, _ = _a18_port_inputs(name)
    lp = port.flow_log_prob(fsd, dcfg, batch["extract_1"], ctx_port, ex_port, batch["eps"], eps_cif=batch.get("eps_cif"))
    assert (lp - gold["log_prob"]).abs().max().item() < 1e-3
    assert abs(-lp.mean().item() - gold["loss"].item()) / abs(gold["loss"].item()) < 1e-4


@pytest.mark.parametrize("name", A18)
def test_packed_a18_algebra_matches_port(name):
    """pack_flow for the same variants (Reverse folded into the CIF block's affine coupling, every permuter as one matrix with
    ActNorm, un-interleaved spline / exponential conditioner outputs), interpreted on the CPU, against the port."""
    cfg, dcfg, fsd, batch, ctx_port, ctx, ex_port, extra = _a18_port_inputs(name)
    want = port.flow_log_prob(fsd, dcfg, batch["extract_1"], ctx_port, ex_port, batch["eps"], eps_cif=batch.get("eps_cif"))
    packed = packing.pack_flow(fsd, cfg, "tf32")
    got = arena_sim.flow_log_prob(packed, batch["extract_1"], ctx, extra, batch["eps"], batch.get("eps_cif"))
    assert (got - want).abs().max().item() < 2e-4


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_a18_state_dict_layout_matches_reference():
    models, mi = refload.load()
    from oracle.make_golden import FIXTURES
    for name in A18:
        label, over, *_ = FIXTURES[name]
        cfg = configs.get_config(label, **dict(over, n_flow_layers=2))
        md = mi.initialize_flow(dict(cfg), "cpu", "test")
        fsd, _ = spec.random_state_dicts(cfg, seed=0)
        rf = md["flow"].state_dict()
        assert list(rf.keys()) == list(fsd.keys()), name
        assert all(tuple(rf[k].shape) == tuple(fsd[k].shape) and rf[k].dtype == fsd[k].dtype for k in rf), name


def test_rq_spline_port_inverse_round_trip():
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(64, 40, generator=g) * 8 - 4)           # some elements in the linear tails
    uw, uh, ud = torch.randn(64, 40, 8, generator=g), torch.randn(64, 40, 8, generator=g), torch.randn(64, 40, 9, generator=g)
    y, lad = port.rq_spline(x, uw, uh, ud)
    xb, _ = port.rq_spline(y, uw, uh, ud, inverse=True)
    assert (xb - x).abs().max().item() < 1e-3                # fp32: bins with slope ~1e-3 amplify the rounding of y
    y64, _ = port.rq_spline(x.double(), uw.double(), uh.double(), ud.double())
    xb64, _ = port.rq_spline(y64, uw.double(), uh.double(), ud.double(), inverse=True)
    assert (xb64 - x.double()).abs().max().item() < 1e-9
    outside = x.abs() > 3
    assert torch.equal(y[outside], x[outside]) and (lad[outside] == 0).all()


def test_fps_oracle_is_furthest_point_sampling():
    """oracle/dataops_ref.py: fps (restating torch_cluster.fps as called by reference dataloaders/ams_voxel_loader.py:298-307):
    starts at point 0, every pick maximises the distance to the already selected set (checked in float64), no repeats."""
    import numpy as np
    from oracle import dataops_ref
    pts = np.random.RandomState(1).rand(400, 6).astype(np.float32)
    idx = dataops_ref.fps(pts, 40)
    assert idx[0] == 0 and len(set(idx.tolist())) == 40
    x = pts.astype(np.float64)
    for j in range(1, 40):
        d = ((x[:, None, :] - x[None, idx[:j], :]) ** 2).sum(-1).min(1)
        assert d[idx[j]] >= d.max() * (1 - 1e-6)


class _PortEngine:
    """Stands in for FlowCompareB200 UNDER the product's drop-in adapters (engine._FlowAdapter / _EmbedderAdapter) on a machine
    without a GPU: same method signatures, the product's own argument normalisation (`FlowCompareB200._prep_context`), the CPU
    port doing the arithmetic.  Test infrastructure: lets the reference's unmodified callers be run on the adapters here."""

    def __init__(self, fsd, esd, dcfg):
        self.fsd, self.esd, self.cfg = fsd, esd, dcfg
        self.device = torch.device("cpu")
        self.is_global, self.has_extra = bool(dcfg["global"]), bool(dcfg["extra_z_value_context"])
        self.d_in = dcfg["input_dim"]

    def embed(self, extract_0):
        f = port.dgcnn_embed_global if self.is_global else port.dgcnn_embed
        return f(self.esd, extract_0[:, :, :self.d_in], self.cfg["n_neighbors"])[0]

    def _expand(self, context, extra_context, B, P):
        from flowcompare_b200.engine import FlowCompareB200
        ctx, _, extra = FlowCompareB200._prep_context(self, context, extra_context, B)
        if self.is_global:
            ctx = ctx.unsqueeze(1).expand(-1, P, -1)
        return ctx, (None if extra is None else extra.reshape(B, 1, 1).expand(-1, P, -1))

    def log_prob(self, x, context, extra_context=None, eps=None, eps_cif=None):
        ctx, ex = self._expand(context, extra_context, x.shape[0], x.shape[1])
        return port.flow_log_prob(self.fsd, self.cfg, x[..., :self.d_in], ctx, ex, eps, eps_cif=eps_cif)

    def sample(self, n_points, context, extra_context=None, z=None, seed=None, eps_cif=None):
        ctx, ex = self._expand(context, extra_context, context.shape[0], n_points)
        return port.flow_sample(self.fsd, self.cfg, z, ctx, ex, eps_cif=eps_cif)


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", TINY)
def test_reference_callers_run_unmodified_on_the_adapters(name):
    """The reference's OWN `inner_loop` and `make_sample` (model_initialization.py:206-245), imported unmodified, are handed a
    models_dict made of the product's adapter classes (what `accelerate` returns; the engine under them is the CPU stand-in
    above because there is no GPU here): the calls they make -- `input_embedder(extract_0)`,
    `flow.log_prob(x, context=, extra_context=)` with the einops-repeated [B,N,1] extra context and [B,P,E] global embedding,
    `flow.sample(num_samples=1, n_points=, context=, sample_distrib=, extra_context=)` -- bind to the adapters' signatures and
    give the goldens the same functions produced on the reference's own modules."""
    from flowcompare_b200.engine import _EmbedderAdapter, _FlowAdapter
    from oracle.make_sample_golden import Injected, base_draw
    _, mi = refload.load()
    cfg, fsd, esd, batch = fixture_inputs(name)
    dcfg = configs.derive(cfg)
    eng = _PortEngine(fsd, esd, dcfg)
    md = {"parameters": [], "flow": _FlowAdapter(eng), "input_embedder": _EmbedderAdapter(eng)}
    md["flow"].eps = batch["eps"]
    loss, lp, bpd = mi.inner_loop((batch["extract_0"], batch["extract_1"], batch["extra_context"]), md, dcfg)
    gold = load_golden(name)
    assert (lp - gold["log_prob"]).abs().max().item() < 1e-3
    assert abs(loss.item() - gold["loss"].item()) / abs(gold["loss"].item()) < 1e-4
    assert abs(bpd.item() - gold["bpd"].item()) / abs(gold["bpd"].item()) < 1e-4
    gs = load_golden("sample_" + name)
    z = base_draw(name, cfg, batch["extract_0"].shape[0])
    x = mi.make_sample(gs["n_points"], batch["extract_0"], md, dcfg, sample_distrib=Injected(z), extra_context=batch["extra_context"])
    assert x.shape == gs["x"].shape and (x - gs["x"]).abs().max().item() < 1e-4


def test_evaluate_on_batches_host_logic(monkeypatch):
    """engine.FlowCompareB200.evaluate_on_batches (the loop of reference test_flow.py:147-226) with the CUDA calls under it
    replaced by the CPU port: stacking the 1|0 and 0|0 passes, the change means and the running nats equal the reference's loop
    written out with separate port calls."""
    from flowcompare_b200 import engine as eng_mod
    cfg, fsd, esd, batch = fixture_inputs("tiny_dgcnn_attn_extra")
    dcfg = configs.derive(cfg)
    e = object.__new__(eng_mod.FlowCompareB200)             # no device, no library: only the host logic is exercised
    e.d_in, e.device, e.has_extra = dcfg["input_dim"], torch.device("cpu"), True
    e.inner_loop = lambda b, eps=None: port.inner_loop((b[0], b[1], b[2].reshape(-1, 1)), fsd, esd, dcfg, eps)
    monkeypatch.setattr(eng_mod, "log_prob_to_change", port.log_prob_to_change)
    B = batch["extract_0"].shape[0]
    data, epss = [], []
    for i in range(2):
        pair, ee = [], []
        for j in range(2):
            b = spec.synthetic_batch(cfg, B, seed=60 + 2 * i + j)
            pair.append((b["extract_0"], b["extract_1"], b["extra_context"]))
            ee.append(b["eps"])
        data.append(tuple(pair))
        epss.append(torch.cat(ee, dim=0))
    nats_avg, means = e.evaluate_on_batches(data, multiple=1.0, eps=epss)
    want_nats, want_means = 0.0, []
    for i, (b10, b00) in enumerate(data):
        _, lp10, nats = port.inner_loop(b10, fsd, esd, dcfg, epss[i][:B])
        _, lp00, _ = port.inner_loop(b00, fsd, esd, dcfg, epss[i][B:])
        change = port.log_prob_to_change(lp10, lp00, 1.0)
        want_means.extend((change > 0).float().mean(dim=-1).tolist())
        want_nats = (want_nats * i + nats.item()) / (i + 1)
    assert len(means) == 2 * B and any(m > 0 for m in means)
    assert all(abs(a - b) <= 2.0 / batch["extract_1"].shape[1] for a, b in zip(means, want_means))   # CPU BLAS is not batch invariant to the bit
    assert abs(nats_avg - want_nats) < 1e-5 * abs(want_nats)


@pytest.mark.parametrize("case", ["loader_default", "fine", "single_layer", "planar"])
def test_voxelize_oracle_matches_reference(case):
    """oracle/dataops_ref.voxelize (centres + nearest-centre labels through the canonical kNN oracle) and the product's host-side
    centre grid (flowcompare_b200.dataops.voxel_centers) against the UNMODIFIED reference's utils.voxelize (utils.py:446-454;
    tests/golden/voxelize.pt): centres bit-exact incl. their order; labels equal except where the two candidate centres are
    equidistant within the rounding noise of the reference's own fp32 formula (see tests/conftest.voxel_label_mismatch)."""
    from flowcompare_b200 import dataops
    from oracle import dataops_ref
    from oracle.make_voxelize_golden import inputs
    gold = load_golden("voxelize")[case]
    pos, start, end, size = inputs(case)
    labels, centers = dataops_ref.voxelize(pos, start, end, size)
    assert torch.equal(centers, gold["centers"])
    assert torch.equal(dataops.voxel_centers(start, end, size), gold["centers"])
    assert voxel_label_mismatch(pos, centers, labels, gold["labels"]) <= 5e-3
    if refload.available():
        refload.load()
        import utils
        l2, c2 = utils.voxelize(pos, start=start, end=end, size=size)
        assert torch.equal(l2, gold["labels"]) and torch.equal(c2, gold["centers"])


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_voxel_centers_equal_live_reference_on_random_boxes():
    """flowcompare_b200.dataops.voxel_centers against the centres the UNMODIFIED utils.voxelize builds (utils.py:448-451), 60 random
    boxes in 2 and 3 dimensions incl. extents that are exact multiples of the voxel size and boxes thinner than one voxel:
    bit-identical values in the same order."""
    from flowcompare_b200 import dataops
    refload.load()
    import utils
    g = torch.Generator().manual_seed(99)
    for case in range(60):
        D = 2 + case % 2
        size = torch.rand(D, generator=g) * 3 + 0.25
        start = (torch.rand(D, generator=g) - 0.5) * 200
        n = torch.randint(0, 7, (D,), generator=g).float()
        extent = (n + 1) * size if case % 3 == 0 else (n + torch.rand(D, generator=g)) * size   # exact multiples every third case
        end = start + extent
        pos = start + torch.rand(5, D, generator=g) * extent.clamp_min(1e-3)
        _, want = utils.voxelize(pos, start=start, end=end, size=size)
        got = dataops.voxel_centers(start, end, size)
        assert got.shape == want.shape and torch.equal(got, want), (case, start, end, size)


@pytest.mark.parametrize("kind", ["LinearLU", "random_permute", "FullCombiner", "ExponentialCombiner"])
def test_permuter_fold_and_its_inverse(kind):
    """packing.permuter_matrix / permuter_inverse_matrix (the pack-time form of every permuter, reference model_initialization.py:
    117-131): the matrix reproduces the port's forward transform and log-det, and inverse @ forward = I."""
    cfg = configs.tiny_config("dgcnn_attn", permuter_type=kind, latent_dim=24, cif_latent_dim=24)
    dcfg = configs.derive(cfg)
    fsd, _ = spec.random_state_dicts(cfg, seed=5)
    t = 3     # transforms.1 coupling, .2 ActNorm, .3 permuter
    W, ldj = packing.permuter_matrix(fsd, t, dcfg)
    Winv = packing.permuter_inverse_matrix(fsd, t, dcfg)
    assert (Winv @ W - torch.eye(24, dtype=torch.float64)).abs().max().item() < 1e-9
    x = torch.randn(2, 5, 24, generator=torch.Generator().manual_seed(1))
    y, l = port.permuter_forward(fsd, t, x, dcfg)
    assert (y.double() - x.double() @ W.t()).abs().max().item() < 1e-5
    assert abs(float(l.reshape(-1)[0]) - ldj) < 1e-4
    assert (port.permuter_inverse(fsd, t, y, dcfg) - x).abs().max().item() < 1e-4
