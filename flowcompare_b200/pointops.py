"""Host-side mirror of the reference's `pointops` functions (models/scene_seg_PAConv/lib/pointops/functions/pointops.py)
for the ops the PAConv embedder and the data side use: same names, argument order and meaning, CUDA tensors in and out,
int32 indices.  Every function is one call into libflowcompare_b200.so (csrc/paconv.cu); there is no CPU path.

The reference functions take channel-major feature tensors [B, C, N]; these wrappers accept the same and transpose to the
library's point-major layout (the embedder itself never transposes: it stays point-major end to end).
"""
import torch

from . import lib as _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(t):
    assert t.is_cuda, "flowcompare_b200.pointops works on CUDA tensors only (no CPU fallback)"
    return t.to(torch.float32).contiguous()


def furthestsampling(xyz, m, tie_block=0, return_xyz=False):
    """`furthestsampling(xyz, m)` (pointops.py:47-62): xyz [B,n,3] -> idx [B,m] int32, first pick = point 0.
    tie_block: 0 = equal distances resolved exactly as the reference's launch does, < 0 = lowest index first."""
    lib = _lib.load()
    xyz = _f32(xyz)
    B, n, _ = xyz.shape
    idx = torch.empty(B, m, dtype=torch.int32, device=xyz.device)
    new_xyz = torch.empty(B, m, 3, dtype=torch.float32, device=xyz.device) if return_xyz else None
    with torch.cuda.device(xyz.device):
        _lib.check(lib.fc_fps(xyz.data_ptr(), B, n, m, tie_block, idx.data_ptr(), 0 if new_xyz is None else new_xyz.data_ptr(),
                              _stream()), "fc_fps")
    return (idx, new_xyz) if return_xyz else idx


def knnquery_heap(nsample, xyz, new_xyz=None, return_dist2=False):
    """`knnquery_heap(nsample, xyz, new_xyz)` (pointops.py:475-497): -> idx [B,m,nsample] int32 (ascending distance)."""
    lib = _lib.load()
    xyz = _f32(xyz)
    new_xyz = xyz if new_xyz is None else _f32(new_xyz)
    B, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.empty(B, m, nsample, dtype=torch.int32, device=xyz.device)
    d2 = torch.empty(B, m, nsample, dtype=torch.float32, device=xyz.device) if return_dist2 else None
    with torch.cuda.device(xyz.device):
        _lib.check(lib.fc_knn_heap(xyz.data_ptr(), new_xyz.data_ptr(), B, n, m, nsample, idx.data_ptr(),
                                   0 if d2 is None else d2.data_ptr(), _stream()), "fc_knn_heap")
    return (idx, d2) if return_dist2 else idx


def nearestneighbor(unknown, known, squared=False):
    """`nearestneighbor(unknown, known)` (pointops.py:96-115): -> (dist [B,n,3] = sqrt of the squared distances, idx [B,n,3]);
    squared=True returns the kernel's squared distances instead (what the reference's CUDA kernel itself writes)."""
    lib = _lib.load()
    unknown, known = _f32(unknown), _f32(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = torch.empty(B, n, 3, dtype=torch.float32, device=unknown.device)
    idx = torch.empty(B, n, 3, dtype=torch.int32, device=unknown.device)
    with torch.cuda.device(unknown.device):
        _lib.check(lib.fc_three_nn(unknown.data_ptr(), known.data_ptr(), B, n, m, d2.data_ptr(), idx.data_ptr(), _stream()),
                   "fc_three_nn")
    return (d2 if squared else torch.sqrt(d2)), idx


def interpolation(features, idx, weight):
    """`interpolation(features, idx, weight)` (pointops.py:121-140): features [B,c,m], idx / weight [B,n,3] -> [B,c,n]."""
    lib = _lib.load()
    f = _f32(features).transpose(1, 2).contiguous()          # [B,m,c] point-major
    B, m, c = f.shape
    n = idx.shape[1]
    idx = idx.to(torch.int32).contiguous()
    weight = _f32(weight)
    out = torch.empty(B, n, c, dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device):
        _lib.check(lib.fc_three_interpolate(f.data_ptr(), c, c, idx.data_ptr(), weight.data_ptr(), B, n, m, out.data_ptr(), c,
                                            _stream()), "fc_three_interpolate")
    return out.transpose(1, 2).contiguous()


def grouping(features, idx):
    """`grouping(features, idx)` (pointops.py:158-175): features [B,c,n], idx [B,m,k] -> [B,c,m,k]."""
    lib = _lib.load()
    f = _f32(features).transpose(1, 2).contiguous()          # [B,n,c]
    B, n, c = f.shape
    m, k = idx.shape[1], idx.shape[2]
    idx = idx.to(torch.int32).contiguous()
    out = torch.empty(B, m, k, c, dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device):
        _lib.check(lib.fc_group_points(f.data_ptr(), c, c, idx.data_ptr(), B, n, m, k, out.data_ptr(), c, _stream()),
                   "fc_group_points")
    return out.permute(0, 3, 1, 2).contiguous()


def gathering(features, idx):
    """`gathering(features, idx)` (pointops.py:68-90): features [B,c,n], idx [B,m] -> [B,c,m]."""
    return grouping(features, idx.unsqueeze(-1)).squeeze(-1)
