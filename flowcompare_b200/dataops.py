"""Host-side mirror of the data-side functions that feed `inner_loop`: same names, argument meaning and
return values as the reference's, CUDA tensors in and out, one call into libflowcompare_b200.so each (csrc/dataops.cu).
There is no CPU path."""
import torch

from . import lib as _lib
from .engine import get_knn


def _stream():
    return torch.cuda.current_stream().cuda_stream


def fps_subsample(points, n_samples, return_index=False):
    """`voxel[fps(voxel, zeros, ratio=n_samples / voxel.shape[0], random_start=False), :][:n_samples, :]` (reference
    dataloaders/ams_voxel_loader.py:298-307): furthest point sampling over all columns of `points` ([n, C] or [B, n, C],
    C = 3 or 6), starting at point 0.  Returns the selected rows (and their int32 indices)."""
    assert points.is_cuda, "flowcompare_b200.dataops works on CUDA tensors only (no CPU fallback)"
    single = points.dim() == 2
    p = (points.unsqueeze(0) if single else points).to(torch.float32).contiguous()
    B, n, C = p.shape
    m = min(int(n_samples), n)
    idx = torch.empty(B, m, dtype=torch.int32, device=p.device)
    out = torch.empty(B, m, C, dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        _lib.check(_lib.load().fc_fps_points(p.data_ptr(), C, B, n, C, m, idx.data_ptr(), out.data_ptr(), C, _stream()),
                   "fc_fps_points")
    if single:
        out, idx = out[0], idx[0]
    return (out, idx) if return_index else out


def co_unit_sphere(points_0, points_1, return_inverse=False):
    """`co_unit_sphere(points_0, points_1, return_inverse)` (reference utils.py:259-280): joint zero-mean / unit-ball
    normalisation of the xyz columns, [n, C] clouds or batches [B, n, C] of pairs.  Like the reference it works on copies of
    the caller's rows (its `torch.cat` copies) and returns the two normalised clouds; inverse = {'furthest_distance', 'mean'}."""
    assert points_0.is_cuda and points_1.is_cuda, "flowcompare_b200.dataops works on CUDA tensors only (no CPU fallback)"
    single = points_0.dim() == 2
    a = (points_0.unsqueeze(0) if single else points_0).to(torch.float32).contiguous().clone()
    b = (points_1.unsqueeze(0) if single else points_1).to(torch.float32).contiguous().clone()
    B = a.shape[0]
    inv = torch.empty(B, 4, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().fc_co_unit_sphere(a.data_ptr(), a.shape[1], a.shape[2], b.data_ptr(), b.shape[1], b.shape[2], B,
                                                 inv.data_ptr(), _stream()), "fc_co_unit_sphere")
    if single:
        a, b, inverse = a[0], b[0], {"furthest_distance": inv[0, 3], "mean": inv[0, :3]}
    else:
        inverse = {"furthest_distance": inv[:, 3], "mean": inv[:, :3]}
    return (a, b, inverse) if return_inverse else (a, b)


def voxel_centers(start, end, size):
    """The centre grid of `voxelize` (reference utils.py:448-451): per axis `arange(start + size/2, end + size/2, size)`, all
    combinations with the FIRST axis running fastest, [m, D] fp32.  Host-side like the reference's (its `torch.arange` lands on
    the CPU whatever device the bounds live on): a few hundred rows, built once per cluster."""
    start, end, size = (torch.as_tensor(v, dtype=torch.float32).cpu() for v in (start, end, size))
    lo, hi = start + size / 2, end + size / 2                     # fp32, as the reference's per-axis tensor arithmetic
    axes = [torch.arange(lo[d].item(), hi[d].item(), size[d].item()) for d in range(size.numel())]
    cols, inner = [], 1
    total = 1
    for a in axes:
        total *= a.numel()
    for a in axes:
        cols.append(a.repeat_interleave(inner).repeat(total // (inner * a.numel())))
        inner *= a.numel()
    return torch.stack(cols, dim=1)


def voxelize(pos, start, end, size):
    """`voxelize(pos, start, end, size)` (reference utils.py:446-454, called by dataloaders/ams_voxel_loader.py:204): the voxel
    centres of the box and, for every point of `pos` [n, D] (CUDA), the index of its nearest centre -- the reference's
    `get_knn(pos, centers, 1)`, here one launch of the bit-exact kNN query (csrc/knn.cu, fc_knn_query) instead of an n x m
    distance matrix and a top-k.  Returns (labels [n, 1] int64, centers [m, D]) on pos.device."""
    assert pos.is_cuda, "flowcompare_b200.dataops works on CUDA tensors only (no CPU fallback)"
    centers = voxel_centers(start, end, size).to(pos.device)
    return get_knn(pos, centers, 1), centers
