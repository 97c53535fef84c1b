"""Benchmark configurations, keyed by ARCHITECTURE label.

The reference's YAML files do not match the README / BASELINE.json labels
(SURVEY.md A.1), so the five configs are defined by what they build, using exactly
the keys `initialize_flow` reads (reference `model_initialization.py:30-202`,
`config/*.yaml`).  A dict from here can be passed to the reference's
`initialize_flow` unchanged, and a dict loaded from a reference YAML / checkpoint
can be passed to `FlowCompareB200` unchanged.
"""
import copy

# keys shared by all shipped configs (reference config/*.yaml, identical in all five)
_COMMON = {
    "sample_size": 1024,
    "n_flow_layers": 115,
    "flow_type": "AffineCoupling",
    "hidden_dims": [512, 512, 512],
    "hidden_dims_embedder_out": [512] * 6,
    "permuter_type": "LinearLU",
    "input_dim": 6,
    "data_parallel": False,
    "coupling_block_nonlinearity": "GELU",
    "attn_dim": 512,
    "latent_dim": 300,
    "attn_input_dim": 256,
    "input_embedding_dim": 64,
    "cross_heads": 1,
    "cross_dim_head": 64,
    "attn_dropout": 0.0,
    "amp": False,
    "input_embedder": "DGCNNembedder",
    "n_neighbors": 40,
    "eps_expm": 1e-8,
    "augmenter_dist": "ConditionalNormal",
    "net_augmenter_dist_hidden_dims": [512, 512, 512],
    "pre_attention_mlp_hidden_dims": [256, 256, 256],
    "cif_dist": "ConditionalNormal",
    "net_cif_dist_hidden_dims": [64, 64],
    "cif_latent_dim": 300,
    "coupling_expm_algo": "torch",
    "act_norm": True,
    "cif_act_norm": True,
    "clamp_dist": 10.0,
    "num_bins_spline": 8,
    "linear_lu_eps": 1e-5,
    "affine_scale_fn": "sigmoid",
    "affine_cif_hidden": [256, 256, 256],
    "n_samples_context": 1250,
    "use_attn_augment": True,
    "extra_z_value_context": False,
    "batch_size": 20,
}


def _mk(**kw):
    c = copy.deepcopy(_COMMON)
    c.update(kw)
    return c


# label -> (config, reference yaml whose CONTENT matches)
CONFIGS = {
    # configs[0]: "good-surf (DGCNN Global embedder)"  == contents of config/helpful-sponge.yaml
    "dgcnn_global": _mk(input_embedder="DGCNNembedderGlobal", input_embedding_dim=124,
                        hidden_dims=[512] * 6, hidden_dims_embedder_out=[512] * 4, batch_size=25),
    # configs[1]: "summer-terrain (DGCNN Attention, no extra context)" == config/swept-energy.yaml
    "dgcnn_attn": _mk(),
    # configs[2]: "helpful-sponge (PAConv Attention)" == config/summer-terrain.yaml
    "paconv_attn": _mk(input_embedder="PAConv", batch_size=25),
    # configs[3]: "dulcet-universe (DGCNN Attention + extra context)" == config/dulcet-universe.yaml
    "dgcnn_attn_extra": _mk(extra_z_value_context=True),
    # (config/good-surf.yaml content: PAConv + extra context)
    "paconv_attn_extra": _mk(input_embedder="PAConv", extra_z_value_context=True, batch_size=25),
}

REFERENCE_YAML = {
    "dgcnn_global": "helpful-sponge",
    "dgcnn_attn": "swept-energy",
    "paconv_attn": "summer-terrain",
    "dgcnn_attn_extra": "dulcet-universe",
    "paconv_attn_extra": "good-surf",
}

BASELINE_LABEL = {
    "dgcnn_global": "good-surf (DGCNN Global embedder)",
    "dgcnn_attn": "summer-terrain (DGCNN Attention/Perceiver, no extra context)",
    "paconv_attn": "helpful-sponge (PAConv Attention embedder)",
    "dgcnn_attn_extra": "dulcet-universe (DGCNN Attention + extra context)",
}


def get_config(label: str, **overrides) -> dict:
    c = copy.deepcopy(CONFIGS[label])
    c.update(overrides)
    return c


def tiny_config(label: str = "dgcnn_attn", **overrides) -> dict:
    """Same architecture, few flow layers / points: for fixtures that must stay small."""
    c = get_config(label, n_flow_layers=3, sample_size=96, n_samples_context=128, n_neighbors=40)
    c.update(overrides)
    return c


def derive(config: dict) -> dict:
    """Derived keys exactly as `initialize_flow` adds them (model_initialization.py:33-45)."""
    c = dict(config)
    c["extra_context_dim"] = 1 if c["extra_z_value_context"] else 0
    c["using_extra_context"] = c["extra_context_dim"] > 0
    c["global"] = c["input_embedder"] in ("DGCNNembedderGlobal",)
    return c
