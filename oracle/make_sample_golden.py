"""TEST INFRASTRUCTURE ONLY -- golden outputs of the UNMODIFIED reference's generative pass.

    python -m oracle.make_sample_golden      # needs /root/reference; writes tests/golden/sample_<fixture>.pt

`make_sample(n_points, extract_0, models_dict, config, sample_distrib, extra_context)` (reference
model_initialization.py:231-245 -> Flow.sample, models/transform.py:79-84) on the tiny fixtures of oracle/make_golden.py,
with the base draw injected through `sample_distrib` (an object with the reference's `.sample(num_samples, n_points=)`),
so the only randomness of the pass is a stored input.  The base draw z is regenerated from its seed (spec._randn)."""
import os
import sys

import torch

from flowcompare_b200 import configs, spec
from oracle import refload
from oracle.make_golden import GOLDEN_DIR, fixture_inputs

SAMPLE_FIXTURES = {"tiny_dgcnn_attn": 40, "tiny_dgcnn_attn_extra": 33, "tiny_dgcnn_global": 64, "tiny_paconv_attn": 50,
                   "mid_dgcnn_attn": 48,
                   # the transforms no shipped config selects (oracle/make_golden.py: A18): their `.inverse` paths
                   "a18_spline": 40, "a18_spline_extra300": 24, "a18_expo": 40, "a18_expo_global300": 24, "a18_cif": 40,
                   "a18_cif_spline_clamp": 36, "a18_permute_relu": 40, "a18_fullcombiner_noactnorm": 40,
                   "a18_expcombiner_global": 40, "a18_identity_augmenter": 40}


def base_draw(name, cfg, B):
    n_points = SAMPLE_FIXTURES[name]
    g = torch.Generator().manual_seed(4242 + n_points)
    return spec._randn((B, n_points, cfg["latent_dim"]), g) * 0.6      # sample_dist = Normal(0, 0.6)


def cif_draws(name, cfg, B):
    """[L, B, n_points, S] draws of the CIF blocks' `Slice.inverse` (reference models/slice.py:46-58), indexed by forward layer; None
    for flows without CIF blocks."""
    S = cfg["cif_latent_dim"] - cfg["latent_dim"]
    if S <= 0:
        return None
    n_points = SAMPLE_FIXTURES[name]
    g = torch.Generator().manual_seed(777 + n_points)
    return spec._randn((cfg["n_flow_layers"], B, n_points, S), g)


class Injected(torch.nn.Module):
    def __init__(self, z):
        super().__init__()
        self.z = z

    def sample(self, num_samples, context=None, n_points=None):
        assert num_samples == 1 and n_points == self.z.shape[1]
        return self.z.clone()


def main(argv):
    torch.set_grad_enabled(False)
    models, mi = refload.load()
    for name, n_points in SAMPLE_FIXTURES.items():
        if len(argv) > 1 and argv[1] not in name:
            continue
        cfg, fsd, esd, batch = fixture_inputs(name)
        if cfg["input_embedder"] == "PAConv":
            from oracle import port_paconv
            port_paconv.patch_reference_pointops()
        torch.manual_seed(0)
        md = mi.initialize_flow(dict(cfg), "cpu", "test")
        md["flow"].load_state_dict(fsd)
        md["input_embedder"].load_state_dict(esd)
        dcfg = configs.derive(cfg)
        z = base_draw(name, cfg, batch["extract_0"].shape[0])
        ec = cif_draws(name, cfg, batch["extract_0"].shape[0])
        import torch.distributions.normal as tdn
        it = iter(reversed(list(ec))) if ec is not None else iter(())       # the reference inverts (and draws) last block first
        orig = tdn._standard_normal
        tdn._standard_normal = lambda shape, dtype, device: next(it).to(dtype).reshape(shape)
        try:
            x = mi.make_sample(n_points, batch["extract_0"], md, dcfg, sample_distrib=Injected(z), extra_context=batch["extra_context"])
        finally:
            tdn._standard_normal = orig
        path = os.path.join(GOLDEN_DIR, f"sample_{name}.pt")
        torch.save({"x": x.clone(), "n_points": n_points,
                    "meta": {"fixture": name, "generator": "oracle/make_sample_golden.py", "source": "unmodified reference make_sample, CPU fp32"}}, path)
        print(f"{name}: x {tuple(x.shape)} mean |x| {x.abs().mean().item():.4f} -> {path}")


if __name__ == "__main__":
    main(sys.argv)
