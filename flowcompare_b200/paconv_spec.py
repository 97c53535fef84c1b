"""Parameter layout of the PAConv embedder, `PointNet2SSGSeg(c=input_dim-3, k=E, out_mlp_dims)` with the
default `args={}` the reference passes (reference model_initialization.py:166-167,
models/scene_seg_PAConv/model/pointnet2/pointnet2_paconv_seg.py:29-57): 4 SA levels of 3 PAConv layers
(m=8 weight-bank kernels, ScoreNet hidden [16]), 4 FP levels of SharedMLP, out MLP 128 -> E."""
from collections import OrderedDict

M_KERNELS = 8
SCORE_HIDDEN = 16
NSAMPLE = 32


def sa_mlps(c):
    # use_xyz=True adds 3 to the first width (pointnet2_paconv_modules.py:93-94)
    return [[c + 3, 32, 32, 64], [64 + 3, 64, 64, 128], [128 + 3, 128, 128, 256], [256 + 3, 256, 256, 512]]


def fp_mlps(c):
    return [[128 + c, 128, 128, 128], [256 + 64, 256, 128], [256 + 128, 256, 256], [512 + 256, 256, 256]]


def _bn(prefix, n, out):
    out[f"{prefix}.weight"] = (n,)
    out[f"{prefix}.bias"] = (n,)
    out[f"{prefix}.running_mean"] = (n,)
    out[f"{prefix}.running_var"] = (n,)
    out[f"{prefix}.num_batches_tracked"] = ()


def paconv_param_shapes(cfg):
    from .spec import _mlp_shapes
    c = cfg["input_dim"] - 3
    out = OrderedDict()
    for i, widths in enumerate(sa_mlps(c)):
        for j in range(len(widths) - 1):
            p = f"SA_modules.{i}.mlps.0.layer{j}"
            cin, cout = widths[j], widths[j + 1]
            out[f"{p}.weightbank"] = (2 * cin, M_KERNELS * cout)
            _bn(f"{p}.bn", cout, out)
            out[f"{p}.scorenet.mlp_convs_hidden.0.weight"] = (SCORE_HIDDEN, 3, 1, 1)
            out[f"{p}.scorenet.mlp_convs_hidden.1.weight"] = (M_KERNELS, SCORE_HIDDEN, 1, 1)
            out[f"{p}.scorenet.mlp_convs_hidden.1.bias"] = (M_KERNELS,)
            _bn(f"{p}.scorenet.mlp_bns_hidden.0", SCORE_HIDDEN, out)
            _bn(f"{p}.scorenet.mlp_bns_hidden.1", M_KERNELS, out)   # exists in the state_dict, unused (last_bn=False)
    for i, widths in enumerate(fp_mlps(c)):
        for j in range(len(widths) - 1):
            p = f"FP_modules.{i}.mlp.layer{j}"
            out[f"{p}.conv.weight"] = (widths[j + 1], widths[j], 1, 1)
            _bn(f"{p}.bn.bn", widths[j + 1], out)
    _mlp_shapes("out_mlp", 128, cfg["hidden_dims_embedder_out"], cfg["input_embedding_dim"], out)
    return out
