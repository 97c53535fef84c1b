"""Packs the PAConv embedder (`PointNet2SSGSeg`) state_dict for csrc/paconv.cu (walk order = fc_paconv_create).

Per PAConv layer: the ScoreNet parameters with its first BatchNorm folded in (w0[16][3], b0[16], w1[8][16], b1[8]) and
ONE linear over the expanded input X'[(m, c)] = s_m * x_c whose weight is the weight bank re-laid-out as
W[o][(m, c)] = a_o * WB[c, m*Cout + o] with the layer's eval-mode BatchNorm (a, b) folded in
(reference model/pointnet2/paconv.py:100-102,144-149).  FP layers: conv weight with BN folded.
"""
import numpy as np
import torch

from .packing import EMB_MAGIC, TC_FORMATS, Arena, _bn_fold, _d, _pack_plain_mlp
from .paconv_spec import M_KERNELS, NSAMPLE, fp_mlps, sa_mlps


def latched_npoints(n_context):
    """npoint of each SA level = N//4, latched on the reference's first forward (pointnet2_paconv_modules.py:37-38);
    here it is latched at pack time from config['n_samples_context']."""
    out, n = [], n_context
    for _ in range(4):
        n = n // 4
        out.append(n)
    return out


def pack_paconv(sd, cfg, tc_format="tf32"):
    c = cfg["input_dim"] - 3
    ar = Arena(tc_format)
    for i, widths in enumerate(sa_mlps(c)):
        for j in range(len(widths) - 1):
            p = f"SA_modules.{i}.mlps.0.layer{j}"
            cin, cout = widths[j], widths[j + 1]
            a0, b0 = _bn_fold(sd, f"{p}.scorenet.mlp_bns_hidden.0")
            w0 = _d(sd, f"{p}.scorenet.mlp_convs_hidden.0.weight")[:, :, 0, 0] * a0[:, None]     # [16,3]
            w1 = _d(sd, f"{p}.scorenet.mlp_convs_hidden.1.weight")[:, :, 0, 0]                   # [8,16]
            b1 = _d(sd, f"{p}.scorenet.mlp_convs_hidden.1.bias")
            ar.vector(torch.cat((w0.reshape(-1), b0, w1.reshape(-1), b1)))
            a, b = _bn_fold(sd, f"{p}.bn")
            wb = _d(sd, f"{p}.weightbank")                                                      # [2cin, m*cout]
            assert wb.shape == (2 * cin, M_KERNELS * cout)
            # W[o, m*2cin + c] = a_o * WB[c, m*cout + o]
            W = wb.reshape(2 * cin, M_KERNELS, cout).permute(2, 1, 0).reshape(cout, M_KERNELS * 2 * cin) * a[:, None]
            ar.linear(W, b, M_KERNELS * 2 * cin)
    for i, widths in enumerate(fp_mlps(c)):
        for j in range(len(widths) - 1):
            p = f"FP_modules.{i}.mlp.layer{j}"
            a, b = _bn_fold(sd, f"{p}.bn.bn")
            W = _d(sd, f"{p}.conv.weight")[:, :, 0, 0] * a[:, None]
            ar.linear(W, b, widths[j])
    out_hid, n_out = _pack_plain_mlp(ar, sd, "out_mlp", 128)
    arena, table = ar.finish()
    npts = latched_npoints(cfg["n_samples_context"])
    header = np.asarray([EMB_MAGIC, TC_FORMATS[tc_format], 2, cfg["input_dim"], 0, cfg["input_embedding_dim"], out_hid, n_out]
                        + npts + [NSAMPLE, M_KERNELS], dtype=np.int32)
    return header, table, arena
