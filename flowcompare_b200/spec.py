"""Parameter layout of the hot path, and seeded synthetic weights / inputs.

The layout (keys, shapes) is the reference's own `state_dict` layout
(produced by reference `model_initialization.py:30-202`), so a
`state_dict` made here loads into the reference modules with `load_state_dict`, and a
reference checkpoint loads into `FlowCompareB200` unchanged.

There is no network and no checkpoint, so benchmarks/tests use seeded random weights
(`random_state_dicts`) and seeded synthetic cloud pairs (`synthetic_batch`) of the shapes
the reference's data loader emits.  torch's CPU generator is
deterministic for a fixed torch build, so the same seed gives the same tensors here and
on the GPU box; that is what lets full-size golden outputs be committed as a few KB.
"""
from collections import OrderedDict
import math

import numpy as np
import torch

from .configs import derive


# ----------------------------------------------------------------------------- shapes
def _mlp_shapes(prefix, in_dim, sizes, out_dim, out):
    # reference models/nets.py:8-17 (registration order: in_layer, out_layer, layers)
    out[f"{prefix}.in_layer.weight"] = (sizes[0], in_dim)
    out[f"{prefix}.in_layer.bias"] = (sizes[0],)
    out[f"{prefix}.out_layer.weight"] = (out_dim, sizes[-1])
    out[f"{prefix}.out_layer.bias"] = (out_dim,)
    for i in range(len(sizes) - 1):
        out[f"{prefix}.layers.{i}.weight"] = (sizes[i + 1], sizes[i])
        out[f"{prefix}.layers.{i}.bias"] = (sizes[i + 1],)


def _attn_shapes(prefix, cfg, out):
    # reference models/perceiver.py:89-119 : PreNorm(norm, fn=AttentionControlledOut(attention, lin))
    inner = cfg["cross_heads"] * cfg["cross_dim_head"]
    out[f"{prefix}.fn.attention.to_q.weight"] = (inner, cfg["attn_input_dim"])
    out[f"{prefix}.fn.attention.to_kv.weight"] = (2 * inner, cfg["input_embedding_dim"])
    out[f"{prefix}.fn.lin.weight"] = (cfg["attn_dim"], inner)
    out[f"{prefix}.fn.lin.bias"] = (cfg["attn_dim"],)
    out[f"{prefix}.norm.weight"] = (cfg["attn_input_dim"],)
    out[f"{prefix}.norm.bias"] = (cfg["attn_input_dim"],)


def _coupling_shapes(p, cfg, D, ex, out):
    """PreConditionApplier(flow_type, CouplingPreconditionerAttn | Global): reference models/cif_block.py:42-46,
    model_initialization.py:95-108 (parameters of a module precede those of its sub-modules)."""
    half = D // 2
    if not cfg["global"]:
        _attn_shapes(f"{p}.pre_conditioner.attn", cfg, out)
        _mlp_shapes(f"{p}.pre_conditioner.pre_attention_mlp", half,
                    cfg["pre_attention_mlp_hidden_dims"], cfg["attn_input_dim"], out)
        ctx_dim = cfg["attn_dim"] + ex
    else:
        ctx_dim = cfg["input_embedding_dim"] + ex
    kind = cfg["flow_type"]
    if kind == "AffineCoupling":                       # models/affine_coupling.py:17
        out_dim = (D - half) * 2
    elif kind == "RationalQuadraticSplineCoupling":    # models/spline_coupling.py:179
        out_dim = (cfg["num_bins_spline"] * 3 + 1) * half
    else:                                              # models/exponential_coupling.py:22-31
        for leaf in ("scale", "shift", "rescale", "reshift"):
            out[f"{p}.transform.{leaf}"] = (1,)
        out_dim = (D - half) ** 2 + (D - half)
    _mlp_shapes(f"{p}.transform.nn", half + ctx_dim, cfg["hidden_dims"], out_dim, out)


def _cif_shapes(p, cfg, out):
    """CIFblock, reference models/cif_block.py:50-68 (registration order; augmenter and slicer share ONE net)."""
    D, D2 = cfg["latent_dim"], cfg["cif_latent_dim"]
    S = D2 - D
    out[f"{p}.act_norm.shift"] = (1, D2)
    out[f"{p}.act_norm.log_scale"] = (1, D2)
    out[f"{p}.act_norm.initialized"] = (1,)
    _mlp_shapes(f"{p}.augmenter.noise_dist.net", D, cfg["net_cif_dist_hidden_dims"], S * 2, out)
    _mlp_shapes(f"{p}.affine_cif.nn", S, cfg["affine_cif_hidden"], D * 2, out)
    _coupling_shapes(f"{p}.flow", cfg, D, 0, out)
    _mlp_shapes(f"{p}.slicer.noise_dist.net", D, cfg["net_cif_dist_hidden_dims"], S * 2, out)
    out[f"{p}.reverse.permutation"] = (D2,)
    out[f"{p}.reverse.inv_permutation"] = (D2,)


def flow_param_shapes(config) -> "OrderedDict[str, tuple]":
    """Key -> shape of `models_dict['flow'].state_dict()` for the supported architectures."""
    cfg = derive(config)
    _check_supported(cfg)
    D, d_in, ex = cfg["latent_dim"], cfg["input_dim"], cfg["extra_context_dim"]
    out = OrderedDict()
    out["base_dist.buffer"] = (1,)
    out["sample_dist.loc"] = (1,)
    out["sample_dist.scale"] = (1,)
    out["sample_dist.std_normal.buffer"] = (1,)
    if D > d_in:
        # transforms.0 = AugmentAttentionPreconditioner (model_initialization.py:73-81); latent_dim == input_dim makes it
        # an IdentityTransform without parameters (:93-94)
        _mlp_shapes("transforms.0.augment.noise_dist.net", cfg["attn_dim"] + d_in + ex,
                    cfg["net_augmenter_dist_hidden_dims"], (D - d_in) * 2, out)
        _attn_shapes("transforms.0.attn", cfg, out)
        _mlp_shapes("transforms.0.pre_attn_mlp", d_in, cfg["hidden_dims"], cfg["attn_input_dim"], out)
    L = cfg["n_flow_layers"]
    cif = D < cfg["cif_latent_dim"]
    t = 1
    for layer in range(L):
        p = f"transforms.{t}"
        if cif:
            _cif_shapes(p, cfg, out)
        else:
            _coupling_shapes(p, cfg, D, ex, out)
        t += 1
        if layer != L - 1:
            if cfg["act_norm"]:
                out[f"transforms.{t}.shift"] = (1, D)
                out[f"transforms.{t}.log_scale"] = (1, D)
                out[f"transforms.{t}.initialized"] = (1,)
                t += 1
            kind = cfg["permuter_type"]
            if kind == "LinearLU":
                ntri = (D - 1) * D // 2
                out[f"transforms.{t}.lower_entries"] = (ntri,)
                out[f"transforms.{t}.upper_entries"] = (ntri,)
                out[f"transforms.{t}.unconstrained_upper_diag"] = (D,)
            elif kind == "random_permute":
                out[f"transforms.{t}.permutation"] = (D,)
                out[f"transforms.{t}.inv_permutation"] = (D,)
            else:   # FullCombiner / ExponentialCombiner (models/permuters.py:15-52)
                out[f"transforms.{t}.w"] = (D, D)
                if kind == "ExponentialCombiner":
                    for leaf in ("scale", "shift", "rescale", "reshift"):
                        out[f"transforms.{t}.{leaf}"] = (1,)
            t += 1
    return out


_DGCNN_CONVS = [(64, 12), (64, 128), (128, 128), (256, 256)]


def embedder_param_shapes(config) -> "OrderedDict[str, tuple]":
    cfg = derive(config)
    out = OrderedDict()
    name = cfg["input_embedder"]
    if name in ("DGCNNembedder", "DGCNNembedderGlobal"):
        # reference models/pytorch_gcn.py:52-79 / :112-141; each BN is registered twice
        chans = [64, 64, 128, 256, 512]
        for i, c in enumerate(chans):
            _bn_shapes(f"bn{i + 1}", c, out)
        convs = list(_DGCNN_CONVS)
        if name == "DGCNNembedderGlobal":
            convs[0] = (64, cfg["input_dim"] * 2)
        for i, (co, ci) in enumerate(convs):
            out[f"conv{i + 1}.0.weight"] = (co, ci, 1, 1)
            _bn_shapes(f"conv{i + 1}.1", co, out)
        out["conv5.0.weight"] = (512, 512, 1)
        _bn_shapes("conv5.1", 512, out)
        mlp_in = 1024 if name == "DGCNNembedderGlobal" else 512
        _mlp_shapes("out_mlp", mlp_in, cfg["hidden_dims_embedder_out"], cfg["input_embedding_dim"], out)
    elif name == "PAConv":
        from .paconv_spec import paconv_param_shapes
        return paconv_param_shapes(cfg)
    else:
        raise NotImplementedError(f"input_embedder {name!r}")
    return out


def _bn_shapes(prefix, c, out):
    out[f"{prefix}.weight"] = (c,)
    out[f"{prefix}.bias"] = (c,)
    out[f"{prefix}.running_mean"] = (c,)
    out[f"{prefix}.running_var"] = (c,)
    out[f"{prefix}.num_batches_tracked"] = ()


def _check_supported(cfg):
    """Everything `initialize_flow` (reference model_initialization.py:30-202) can build AND `inner_loop` can run: the
    StandardNormal / context-free augmenters raise a TypeError inside the reference's own `Flow.log_prob`
    (`Augment.forward()` takes no `extra_context`), `affine_scale_fn='exp'` and ELU conditioners are not built here."""
    if cfg["flow_type"] not in ("AffineCoupling", "RationalQuadraticSplineCoupling", "ExponentialCoupling"):
        raise NotImplementedError(f"flow_type {cfg['flow_type']!r}")
    if cfg["flow_type"] == "AffineCoupling" and cfg["affine_scale_fn"] != "sigmoid":
        raise NotImplementedError("affine_scale_fn: only 'sigmoid' (all shipped configs) is built")
    if cfg["permuter_type"] not in ("LinearLU", "random_permute", "FullCombiner", "ExponentialCombiner"):
        raise NotImplementedError(f"permuter_type {cfg['permuter_type']!r}")
    if cfg["latent_dim"] > cfg["cif_latent_dim"]:
        raise ValueError("Augment dim smaller than main latent!")           # as reference models/cif_block.py:48
    if cfg["latent_dim"] < cfg["cif_latent_dim"] and (cfg["using_extra_context"] or cfg["global"]):
        raise NotImplementedError("CIF block with extra context / a global embedding: the reference raises too "
                                  "(models/cif_block.py:33-36)")
    if cfg["latent_dim"] < cfg["input_dim"]:
        raise ValueError("Latent dim < Input dim")                          # as reference model_initialization.py:96
    if cfg["latent_dim"] > cfg["input_dim"] and (cfg["augmenter_dist"] != "ConditionalNormal" or not cfg["use_attn_augment"]):
        raise NotImplementedError("only the attention-conditioned ConditionalNormal augmenter runs in the reference")
    if cfg["coupling_block_nonlinearity"] not in ("GELU", "RELU"):
        raise NotImplementedError("conditioner non-linearity: GELU (all shipped configs) or RELU")


# ----------------------------------------------------------------------------- weights
def random_state_dicts(config, seed: int = 0, perturb: bool = True):
    """(flow_state_dict, embedder_state_dict) with realistic seeded values.

    Linear / conv weights follow nn.Linear's default scale U(+-1/sqrt(fan_in)).  The
    reference initialises ActNorm, LinearLU and BN statistics to identity, which would
    hide bugs in those kernels, so `perturb=True` (default) randomises them mildly
    (SURVEY.md 8c "weight realism"), including NEGATIVE BN gammas (they break any
    max/scale reordering in the EdgeConv fold).
    """
    g = torch.Generator().manual_seed(1000003 * seed + 12345)
    flow = OrderedDict()
    for k, shp in flow_param_shapes(config).items():
        flow[k] = _draw(k, shp, g, perturb, config)
    for k in list(flow):
        if k.endswith(".inv_permutation") and flow[k] is None:
            flow[k] = torch.argsort(flow[k[:-len("inv_permutation")] + "permutation"])
        if ".slicer.noise_dist.net." in k:   # the CIF block's slicer scores with the augmenter's own net (cif_block.py:58)
            flow[k] = flow[k.replace(".slicer.", ".augmenter.")]
    emb = OrderedDict()
    for k, shp in embedder_param_shapes(config).items():
        emb[k] = _draw(k, shp, g, perturb, config)
    # the reference registers each DGCNN BN twice (bnX and convX.1 are the same module)
    for k in list(emb):
        if k.startswith("conv") and ".1." in k:
            i = k[4]
            emb[k] = emb[f"bn{i}." + k.split(".1.", 1)[1]]
    return flow, emb


def _randn(shape, g):
    """Approximately N(0,1) from 12 uniforms (Irwin-Hall).  Only integer->float conversion and IEEE adds
    are involved, so the values are bit-identical on every CPU (torch.randn's vectorised Box-Muller differs
    between AVX2/AVX-512/scalar builds), which keeps the seeded fixtures reproducible on the GPU box."""
    shape = tuple(shape) if not isinstance(shape, int) else (shape,)
    u = torch.rand((12,) + shape, generator=g, dtype=torch.float32)
    acc = u[0]
    for i in range(1, 12):
        acc = acc + u[i]
    return acc - 6.0


def _uniform(shape, bound, g):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound


def _draw(key, shape, g, perturb, config):
    leaf = key.rsplit(".", 1)[-1]
    if key in ("base_dist.buffer", "sample_dist.loc", "sample_dist.std_normal.buffer"):
        return torch.zeros(shape)
    if key == "sample_dist.scale":
        return torch.full(shape, 0.6)
    if leaf == "initialized":
        return torch.ones(shape)
    if leaf == "num_batches_tracked":
        return torch.tensor(1, dtype=torch.long)
    if leaf in ("permutation", "inv_permutation"):
        if key.endswith("reverse.permutation") or key.endswith("reverse.inv_permutation"):
            return torch.arange(shape[0] - 1, -1, -1)                       # models/permuters.py:84 (its own inverse)
        if leaf == "permutation":
            return torch.randperm(shape[0], generator=g)
        return None                                                         # filled in from `permutation` by the caller
    if len(shape) == 1 and shape[0] == 1 and leaf in ("scale", "shift", "rescale", "reshift"):
        # ExponentialCoupling / ExponentialCombiner squashing scalars (models/exponential_coupling.py:22-25)
        base = {"scale": 0.125, "shift": 0.0, "rescale": 1.0, "reshift": 0.0}[leaf]
        # shift / reshift move EVERY entry of the n x n generator: a rank-one part with eigenvalue n * offset, so their
        # perturbation shrinks with n (expm would otherwise amplify the latent by e^(n * 0.04) per layer)
        amp = 0.02 if leaf in ("scale", "rescale") else 0.25 / config["latent_dim"]
        return torch.full(shape, base) + (_randn(shape, g) * amp if perturb else 0.0)
    if leaf == "w":   # FullCombiner (orthogonal at init) / ExponentialCombiner (randn): any well-conditioned matrix will do
        if config["permuter_type"] == "ExponentialCombiner":
            return _randn(shape, g)
        return torch.eye(shape[0]) + _randn(shape, g) * (0.3 / math.sqrt(shape[0]))
    if leaf == "shift":
        return _randn(shape, g) * 0.05 if perturb else torch.zeros(shape)
    if leaf == "log_scale":  # slightly positive mean keeps |z| ~ 1 through 115 random layers
        return _randn(shape, g) * 0.03 + 0.004 if perturb else torch.zeros(shape)
    if leaf in ("lower_entries", "upper_entries"):
        D = config["latent_dim"]
        return _randn(shape, g) * (0.04 / math.sqrt(D)) if perturb else torch.zeros(shape)
    if leaf == "unconstrained_upper_diag":
        base = math.log(math.exp(1 - config["linear_lu_eps"]) - 1)
        t = torch.full(shape, base)
        return t + _randn(shape, g) * 0.05 if perturb else t
    if leaf == "running_mean":
        return _randn(shape, g) * 0.1 if perturb else torch.zeros(shape)
    if leaf == "running_var":
        return torch.rand(shape, generator=g) * 1.5 + 0.25 if perturb else torch.ones(shape)
    if leaf == "weightbank":
        fan_in = shape[0]
        return _uniform(shape, math.sqrt(3.0 / fan_in), g)
    is_norm = (".norm." in key) or key.startswith("bn") or ".bn." in key or ".1." in key \
        or "mlp_bns" in key
    if is_norm and len(shape) == 1:
        if leaf == "weight":
            w = 1.0 + _randn(shape, g) * 0.1 if perturb else torch.ones(shape)
            if perturb and (key.startswith("bn") or ".bn." in key):
                flip = torch.rand(shape, generator=g) < 0.1
                w = torch.where(flip, -w, w)
            return w
        return _randn(shape, g) * 0.05 if perturb else torch.zeros(shape)
    if leaf == "weight":
        fan_in = int(np.prod(shape[1:]))
        w = _uniform(shape, 1.0 / math.sqrt(fan_in), g)
        if config.get("flow_type") == "ExponentialCoupling" and key.endswith("transform.nn.out_layer.weight"):
            # the n x n generator of expm: keep its norm (hence the growth of |z| per layer) independent of n, so that
            # log-probs of a 150-wide coupling stay O(100) and an absolute 1e-3 tolerance is meaningful in fp32
            n2 = config["latent_dim"] - config["latent_dim"] // 2
            w = w * min(1.0, 8.0 / n2)
        return w
    if leaf == "bias":
        return _uniform(shape, 0.05, g)
    raise KeyError(f"no initialiser for {key}")


# ----------------------------------------------------------------------------- inputs
def synthetic_batch(config, batch: int, seed: int = 0, n_context=None, n_target=None,
                    duplicates: bool = False):
    """Seeded synthetic cloud pairs with the shapes/normalisation of the reference loader.

    context xyz ~ U([-1.1,1.1]^2 x [-2.1,2.1]), target xyz ~ U([-1,1]^2 x [-2,2])
    (voxel sizes, reference config/*.yaml:179-184), rgb ~ U[0,1], then joint zero-mean /
    unit-ball normalisation of xyz (reference utils.py:259-280 `co_unit_sphere`).
    Returns dict(extract_0[B,Nc,6], extract_1[B,N,6], extra_context[B,1]|None, eps[B,N,D-6] (+ eps_cif[L,B,N,S])).
    `eps` is the single standard-normal draw the reference makes per forward
    (models/distributions.py:148-153 via Normal.rsample), so it can be injected.
    """
    cfg = derive(config)
    Nc = n_context or cfg["n_samples_context"]
    N = n_target or cfg["sample_size"]
    g = torch.Generator().manual_seed(7919 * seed + 17)
    half0 = torch.tensor([1.1, 1.1, 2.1])
    half1 = torch.tensor([1.0, 1.0, 2.0])
    e0 = torch.rand(batch, Nc, 6, generator=g)
    e1 = torch.rand(batch, N, 6, generator=g)
    e0[..., :3] = (e0[..., :3] * 2 - 1) * half0
    e1[..., :3] = (e1[..., :3] * 2 - 1) * half1
    if duplicates:  # the data loader oversamples small voxels -> exact duplicate points (utils.py:362-370)
        e0[:, Nc // 2:, :] = e0[:, : Nc - Nc // 2, :]
    joint = torch.cat((e0[..., :3], e1[..., :3]), dim=1)
    mean = joint.mean(dim=1, keepdim=True)
    joint = joint - mean
    far = joint.norm(dim=-1).amax(dim=1, keepdim=True).unsqueeze(-1)
    joint = joint / far
    e0[..., :3] = joint[:, :Nc]
    e1[..., :3] = joint[:, Nc:]
    extra = torch.rand(batch, 1, generator=g) if cfg["using_extra_context"] else None
    eps = _randn((batch, N, cfg["latent_dim"] - cfg["input_dim"]), g)
    out = {"extract_0": e0.contiguous(), "extract_1": e1.contiguous(), "extra_context": extra, "eps": eps}
    if cfg["latent_dim"] < cfg["cif_latent_dim"]:
        # one more draw per CIF block (reference models/cif_block.py:74 -> distributions.py:148-153), in list order
        out["eps_cif"] = _randn((cfg["n_flow_layers"], batch, N, cfg["cif_latent_dim"] - cfg["latent_dim"]), g)
    return out
