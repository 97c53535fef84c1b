"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.pt from the UNMODIFIED reference.

Run in the authoring container (needs /root/reference):

    python -m oracle.make_golden            # all fixtures
    python -m oracle.make_golden tiny       # only the small ones

The reference ships no golden vectors, so the pin is the reference's
own output on seeded inputs: weights come from `flowcompare_b200.spec.random_state_dicts`
(loaded into the reference modules with `load_state_dict`, reference
model_initialization.py:18-23), inputs from `spec.synthetic_batch`, and the single RNG
draw of the forward (reference models/distributions.py:148-153) is replaced by the
fixture's `eps` by patching `torch.distributions.normal._standard_normal`.
Weights and inputs are NOT stored (they are regenerated from the seeds; torch's CPU
generator is deterministic for a fixed build) -- only the reference's outputs are.
"""
import os
import sys

import torch

from flowcompare_b200 import configs, spec
from oracle import refload

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

_T = dict(n_flow_layers=3, sample_size=96, n_samples_context=128)
# name -> (arch label, config overrides, batch, weight seed, input seed)
FIXTURES = {
    "tiny_dgcnn_attn": ("dgcnn_attn", dict(n_flow_layers=3, sample_size=96, n_samples_context=128), 2, 11, 21),
    "tiny_dgcnn_attn_extra": ("dgcnn_attn_extra", dict(n_flow_layers=3, sample_size=96, n_samples_context=128), 2, 12, 22),
    "tiny_dgcnn_global": ("dgcnn_global", dict(n_flow_layers=3, sample_size=96, n_samples_context=128), 2, 13, 23),
    "tiny_paconv_attn": ("paconv_attn", dict(n_flow_layers=3, sample_size=96, n_samples_context=320), 2, 14, 24),
    "tiny_paconv_attn_extra": ("paconv_attn_extra", dict(n_flow_layers=2, sample_size=64, n_samples_context=272), 2, 16, 26),
    "mid_dgcnn_attn": ("dgcnn_attn", dict(n_flow_layers=12), 1, 15, 25),
    "full_dgcnn_attn": ("dgcnn_attn", {}, 1, 1, 3),
    "full_dgcnn_attn_extra": ("dgcnn_attn_extra", {}, 1, 2, 4),
    "full_dgcnn_global": ("dgcnn_global", {}, 1, 3, 5),
    "full_paconv_attn": ("paconv_attn", {}, 1, 4, 6),
    "full_paconv_attn_extra": ("paconv_attn_extra", {}, 1, 5, 7),
    # transforms north_star names but no shipped YAML selects, built by editing the config keys
    # `initialize_flow` reads (reference model_initialization.py:95-131, models/cif_block.py:30-46)
    "a18_spline": ("dgcnn_attn", dict(_T, flow_type="RationalQuadraticSplineCoupling", latent_dim=24, cif_latent_dim=24), 2, 31, 41),
    "a18_spline_extra300": ("dgcnn_attn_extra", dict(_T, n_flow_layers=2, flow_type="RationalQuadraticSplineCoupling"), 1, 32, 42),
    "a18_expo": ("dgcnn_attn", dict(_T, flow_type="ExponentialCoupling", latent_dim=24, cif_latent_dim=24), 2, 33, 43),
    "a18_expo_global300": ("dgcnn_global", dict(_T, n_flow_layers=2, sample_size=48, flow_type="ExponentialCoupling",
                                                coupling_expm_algo="original"), 1, 34, 44),
    "a18_cif": ("dgcnn_attn", dict(_T, latent_dim=24, cif_latent_dim=32), 2, 35, 45),
    "a18_cif_spline_clamp": ("dgcnn_attn", dict(_T, latent_dim=24, cif_latent_dim=30, clamp_dist=0.9,
                                                flow_type="RationalQuadraticSplineCoupling"), 2, 36, 46),
    "a18_permute_relu": ("dgcnn_attn", dict(_T, permuter_type="random_permute", coupling_block_nonlinearity="RELU",
                                            latent_dim=24, cif_latent_dim=24), 2, 37, 47),
    "a18_fullcombiner_noactnorm": ("dgcnn_attn", dict(_T, permuter_type="FullCombiner", act_norm=False, latent_dim=24,
                                                      cif_latent_dim=24), 2, 38, 48),
    "a18_expcombiner_global": ("dgcnn_global", dict(_T, permuter_type="ExponentialCombiner", latent_dim=24, cif_latent_dim=24), 2, 39, 49),
    "a18_identity_augmenter": ("dgcnn_attn", dict(_T, latent_dim=6, cif_latent_dim=6), 2, 40, 50),
}
A18 = [k for k in FIXTURES if k.startswith("a18_")]


def fixture_inputs(name):
    """(config, flow_sd, emb_sd, batch dict) for a fixture -- regenerated from seeds."""
    label, over, B, wseed, iseed = FIXTURES[name]
    cfg = configs.get_config(label, **over)
    fsd, esd = spec.random_state_dicts(cfg, seed=wseed)
    batch = spec.synthetic_batch(cfg, B, seed=iseed)
    return cfg, fsd, esd, batch


def run_reference(cfg, fsd, esd, batch):
    models, mi = refload.load()
    import torch.distributions.normal as tdn
    if cfg["input_embedder"] == "PAConv":
        from oracle import port_paconv
        port_paconv.patch_reference_pointops()
    with torch.no_grad():
        torch.manual_seed(0)
        md = mi.initialize_flow(dict(cfg), "cpu", "test")
        md["flow"].load_state_dict(fsd)
        md["input_embedder"].load_state_dict(esd)
        dcfg = configs.derive(cfg)
        orig = tdn._standard_normal
        # every RNG draw of the forward, in the order the reference makes them: the augmenter's (absent for the identity
        # augmenter), then one per CIF block
        draws = ([batch["eps"]] if dcfg["latent_dim"] > dcfg["input_dim"] else []) + list(batch.get("eps_cif", []))
        it = iter(draws)
        tdn._standard_normal = lambda shape, dtype, device: next(it).to(dtype).reshape(shape)
        try:
            loss, lp, bpd = mi.inner_loop((batch["extract_0"], batch["extract_1"], batch["extra_context"]), md, dcfg)
        finally:
            tdn._standard_normal = orig
        emb = md["input_embedder"](batch["extract_0"])
        out = {"log_prob": lp.clone(), "loss": loss.clone(), "bpd": bpd.clone(), "embedding": emb.clone()}
        if cfg["input_embedder"].startswith("DGCNN"):
            from models.pytorch_gcn import knn
            out["knn_idx_layer1"] = knn(batch["extract_0"].permute(0, 2, 1), cfg["n_neighbors"]).to(torch.int32)
    return out


def main(argv):
    only = argv[1] if len(argv) > 1 else ""
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name in FIXTURES:
        if only and not name.startswith(only):
            continue
        if FIXTURES[name][0].startswith("paconv"):
            try:
                from oracle import port_paconv  # noqa: F401
            except ImportError:
                print(f"skip {name}: PAConv oracle not built yet")
                continue
        cfg, fsd, esd, batch = fixture_inputs(name)
        out = run_reference(cfg, fsd, esd, batch)
        out["meta"] = {"fixture": name, "spec": FIXTURES[name], "torch": torch.__version__,
                       "generator": "oracle/make_golden.py", "source": "unmodified reference, CPU fp32"}
        if name.startswith("full") or name.startswith("mid"):
            out["embedding"] = out["embedding"][:, ::8].clone()  # keep fixtures small
            out["embedding_stride"] = 8
        path = os.path.join(GOLDEN_DIR, f"{name}.pt")
        torch.save(out, path)
        print(f"{name}: log_prob mean {out['log_prob'].mean().item():.6f} -> {path} ({os.path.getsize(path)} B)")


if __name__ == "__main__":
    main(sys.argv)
