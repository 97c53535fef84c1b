"""TEST INFRASTRUCTURE ONLY -- oracle for the PAConv (PointNet2SSGSeg) embedder.

Two things live here:

1. Pure-CPU stand-ins for the six `pointops` functions the embedder calls (SURVEY.md 8c "PAConv oracle"):
   the reference wrappers allocate `torch.cuda.*Tensor` and call `pointops_cuda`, which cannot be built
   against modern torch (<THC/THC.h>) and is CUDA-only.  `patch_reference_pointops()` swaps them into the
   imported reference module so that everything ABOVE them (SA / FP / PAConv / ScoreNet modules) is the
   reference's own unmodified code.  The index-producing ones call oracle/pointops_ref.c.
2. `paconv_embed`: a restatement of `PointNet2SSGSeg.forward` (reference
   models/scene_seg_PAConv/model/pointnet2/pointnet2_paconv_seg.py:63-82 and the modules it calls) on plain
   tensors + the reference state_dict, used as the movable oracle on the GPU box.
"""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

from oracle import knn_ref
from oracle.port import mlp


def _lib():
    lib = knn_ref._load()
    if not getattr(lib, "_pointops_ready", False):
        vp, ci = ctypes.c_void_p, ctypes.c_int
        lib.fc_oracle_fps.restype = ci
        lib.fc_oracle_fps.argtypes = [vp, ci, ci, vp]
        lib.fc_oracle_knn_heap.restype = ci
        lib.fc_oracle_knn_heap.argtypes = [vp, vp, ci, ci, ci, vp]
        lib.fc_oracle_three_nn.restype = ci
        lib.fc_oracle_three_nn.argtypes = [vp, vp, ci, ci, vp, vp]
        lib._pointops_ready = True
    return lib


def _np(t):
    return np.ascontiguousarray(t.detach().to(torch.float32).cpu().numpy())


# ----------------------------------------------------------------------------- pointops stand-ins
def furthestsampling(xyz, m):
    """K2: xyz [B,N,3] -> idx [B,m] int32 (reference lib/pointops/functions/pointops.py:47-62)."""
    x = _np(xyz)
    out = np.zeros((x.shape[0], m), dtype=np.int32)
    for b in range(x.shape[0]):
        assert _lib().fc_oracle_fps(x[b].ctypes.data, x.shape[1], m, out[b].ctypes.data) == 0
    return torch.from_numpy(out)


def gathering(features, idx):
    """K3: features [B,C,N], idx [B,m] -> [B,C,m]."""
    return torch.gather(features, 2, idx.long().unsqueeze(1).expand(-1, features.shape[1], -1))


def knnquery_heap(nsample, xyz, new_xyz):
    """K1: xyz [B,N,3], new_xyz [B,m,3] -> idx [B,m,nsample] int32."""
    x, q = _np(xyz), _np(new_xyz)
    out = np.zeros((x.shape[0], q.shape[1], nsample), dtype=np.int32)
    for b in range(x.shape[0]):
        assert _lib().fc_oracle_knn_heap(x[b].ctypes.data, q[b].ctypes.data, x.shape[1], q.shape[1], nsample,
                                        out[b].ctypes.data) == 0
    return torch.from_numpy(out)


def grouping(features, idx):
    """K4: features [B,C,N], idx [B,m,k] -> [B,C,m,k]."""
    B, C, N = features.shape
    m, k = idx.shape[1], idx.shape[2]
    flat = idx.long().reshape(B, 1, m * k).expand(-1, C, -1)
    return torch.gather(features, 2, flat).reshape(B, C, m, k)


def nearestneighbor(unknown, known):
    """K5 + wrapper: -> (sqrt(dist2) [B,n,3], idx [B,n,3] int32)."""
    u, kn = _np(unknown), _np(known)
    d2 = np.zeros((u.shape[0], u.shape[1], 3), dtype=np.float32)
    ix = np.zeros((u.shape[0], u.shape[1], 3), dtype=np.int32)
    for b in range(u.shape[0]):
        assert _lib().fc_oracle_three_nn(u[b].ctypes.data, kn[b].ctypes.data, u.shape[1], kn.shape[1],
                                        d2[b].ctypes.data, ix[b].ctypes.data) == 0
    return torch.sqrt(torch.from_numpy(d2)), torch.from_numpy(ix)


def interpolation(features, idx, weight):
    """K6: features [B,C,m], idx [B,n,3], weight [B,n,3] -> [B,C,n] = sum_3 w * f[idx]."""
    B, C, m = features.shape
    n = idx.shape[1]
    g = torch.gather(features, 2, idx.long().reshape(B, 1, n * 3).expand(-1, C, -1)).reshape(B, C, n, 3)
    w = weight.unsqueeze(1)
    return _fma3(w[..., 0], g[..., 0], w[..., 1], g[..., 1], w[..., 2], g[..., 2])


def _fma3(w0, p0, w1, p1, w2, p2):
    """`w0*p0 + w1*p1 + w2*p2` as the reference's compiled kernel evaluates it (interpolation_cuda_kernel.cu:194, SASS:
    FMUL w1*p1; FFMA w0,p0; FFMA w2,p2).  fp32 inputs: each fused step is emulated through float64, where the product of two
    floats is exact, so only the (rare) double rounding of the sum can differ; other dtypes use the plain expression."""
    if w0.dtype != torch.float32:
        return (w0 * p0 + w1 * p1) + w2 * p2
    t = (w1 * p1).double()
    t = (w0.double() * p0.double() + t).float().double()
    return (w2.double() * p2.double() + t).float()


def patch_reference_pointops():
    """Swap the CPU stand-ins into the imported reference module (oracle/refload.py must have run)."""
    import models.scene_seg_PAConv.lib.pointops.functions.pointops as po
    po.furthestsampling = furthestsampling
    po.gathering = gathering
    po.knnquery_heap = knnquery_heap
    po.grouping = grouping
    po.nearestneighbor = nearestneighbor
    po.interpolation = interpolation


# ----------------------------------------------------------------------------- restatement of the embedder
def _bn(x, sd, prefix):
    """eval-mode BatchNorm over the LAST dim of a channel-last tensor."""
    return (x - sd[f"{prefix}.running_mean"]) / torch.sqrt(sd[f"{prefix}.running_var"] + 1e-5) * sd[f"{prefix}.weight"] \
        + sd[f"{prefix}.bias"]


def paconv_layer(sd, prefix, feat, gxyz):
    """`PAConv.forward` (reference model/pointnet2/paconv.py:107-153), channel-last:
    feat [B,m,K,C], gxyz [B,m,K,3] -> [B,m,K,Cout].  score_input='identity', softmax scores, m=8 kernels,
    kernel_input='neighbor', centre = neighbour 0."""
    x = torch.cat((feat - feat[:, :, :1], feat), dim=-1)
    dxyz = gxyz - gxyz[:, :, :1]
    w0 = sd[f"{prefix}.scorenet.mlp_convs_hidden.0.weight"][:, :, 0, 0]           # [16,3], no bias
    h = F.relu(_bn(dxyz @ w0.t(), sd, f"{prefix}.scorenet.mlp_bns_hidden.0"))
    w1 = sd[f"{prefix}.scorenet.mlp_convs_hidden.1.weight"][:, :, 0, 0]           # [8,16] + bias, no BN (last_bn=False)
    s = torch.softmax(h @ w1.t() + sd[f"{prefix}.scorenet.mlp_convs_hidden.1.bias"], dim=-1)
    wb = sd[f"{prefix}.weightbank"]                                               # [2C, m*Cout]
    n_k = s.shape[-1]
    y = (x @ wb).reshape(*x.shape[:-1], n_k, -1)
    out = torch.einsum("bnkm,bnkmo->bnko", s, y)                                  # assign_score, paconv_util.py:52-56
    return F.relu(_bn(out, sd, f"{prefix}.bn"))


def sa_module(sd, i, xyz, feats, npoint, nsample=32):
    """`_PointNet2SAModuleBase.forward` (pointnet2_paconv_modules.py:20-61) + QueryAndGroup (pointops.py:557-594)."""
    fidx = furthestsampling(xyz, npoint).long()
    new_xyz = torch.gather(xyz, 1, fidx.unsqueeze(-1).expand(-1, -1, 3))
    idx = knnquery_heap(nsample, xyz, new_xyz).long()                             # [B,m,K]
    B, m, K = idx.shape
    gxyz = torch.gather(xyz, 1, idx.reshape(B, m * K, 1).expand(-1, -1, 3)).reshape(B, m, K, 3)
    gfeat = torch.gather(feats, 1, idx.reshape(B, m * K, 1).expand(-1, -1, feats.shape[-1])).reshape(B, m, K, -1)
    f = torch.cat((gxyz - new_xyz.unsqueeze(2), gfeat), dim=-1)
    j = 0
    while f"SA_modules.{i}.mlps.0.layer{j}.weightbank" in sd:
        f = paconv_layer(sd, f"SA_modules.{i}.mlps.0.layer{j}", f, gxyz)
        j += 1
    return new_xyz, f.max(dim=2)[0], fidx, idx


def fp_module(sd, i, unknown, known, unknown_feats, known_feats):
    """`PointNet2FPModule.forward` (pointnet2_paconv_modules.py:206-238) with SharedMLP (util/block.py:14-39)."""
    dist, idx = nearestneighbor(unknown, known)
    w = 1.0 / (dist + 1e-8)
    w = w / w.sum(dim=2, keepdim=True)
    g = torch.gather(known_feats, 1, idx.long().reshape(idx.shape[0], -1, 1).expand(-1, -1, known_feats.shape[-1]))
    g = g.reshape(idx.shape[0], idx.shape[1], 3, -1)
    interp = _fma3(w[..., 0:1], g[:, :, 0], w[..., 1:2], g[:, :, 1], w[..., 2:3], g[:, :, 2])
    f = torch.cat((interp, unknown_feats), dim=-1)
    j = 0
    while f"FP_modules.{i}.mlp.layer{j}.conv.weight" in sd:
        w_ = sd[f"FP_modules.{i}.mlp.layer{j}.conv.weight"][:, :, 0, 0]
        f = F.relu(_bn(f @ w_.t(), sd, f"FP_modules.{i}.mlp.layer{j}.bn.bn"))
        j += 1
    return f


def latched_npoints(n_context):
    """npoint of each SA level is latched to N//4 on the FIRST call (pointnet2_paconv_modules.py:37-38)."""
    out, n = [], n_context
    for _ in range(4):
        n = n // 4
        out.append(n)
    return out


def paconv_embed(sd, pts, config, return_debug=False):
    """pts [B,N,6] -> [B,N,E]."""
    xyz, feats = pts[..., :3].contiguous(), pts[..., 3:].contiguous()
    npoints = latched_npoints(config["n_samples_context"])
    l_xyz, l_feat, dbg = [xyz], [feats], []
    for i in range(4):
        nx, nf, fidx, idx = sa_module(sd, i, l_xyz[i], l_feat[i], npoints[i])
        l_xyz.append(nx)
        l_feat.append(nf)
        dbg.append((fidx, idx))
    for i in range(-1, -5, -1):
        l_feat[i - 1] = fp_module(sd, 4 + i, l_xyz[i - 1], l_xyz[i], l_feat[i - 1], l_feat[i])
    out = mlp(sd, "out_mlp", l_feat[0])
    return (out, dbg, l_xyz, l_feat) if return_debug else out
