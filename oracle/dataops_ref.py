"""TEST INFRASTRUCTURE ONLY -- CPU restatements of the data-side ops; never imported by the product path.

`fps` restates torch_cluster.fps (torch_cluster 1.5.9 pinned by the reference's environment.yml; the package is NOT in this
image and not under /root/reference, so this is its published CPU algorithm, torch_cluster/csrc/cpu/fps_cpu.cpp: start at the
first point of the cloud when random_start=False, `dist = min(dist, ((x - x[cur])**2).sum(1))`, `cur = dist.argmax()` -- the
first maximum) as called by reference dataloaders/ams_voxel_loader.py:298-307.  Parity for this function is pinned to this
restatement only ("parity unpinned" against torch_cluster itself).

`co_unit_sphere` is not restated: the tests import the reference's own utils.co_unit_sphere when /root/reference is present
and otherwise use the golden file made from it (tests/golden/dataops.pt, oracle/make_dataops_golden.py).

`voxelize` restates reference utils.py:446-454 (voxel centres of a box + the nearest centre of every point through
knn.get_knn, knn.py:79-90) and is pinned to the unmodified reference's output (tests/golden/voxelize.pt,
oracle/make_voxelize_golden.py; tests/test_oracle.py::test_voxelize_oracle_matches_reference)."""
import numpy as np


def fps(points, m):
    """points [n, C] float32 -> indices [m] (int64) of the furthest-point sample that starts at point 0."""
    x = np.ascontiguousarray(points, dtype=np.float32)
    n = x.shape[0]
    dist = np.full(n, 3.0e38, dtype=np.float32)
    idx = np.zeros(m, dtype=np.int64)
    cur = 0
    for j in range(1, m):
        d = x - x[cur]
        sq = d * d
        s = sq[:, 0].copy()
        for c in range(1, x.shape[1]):      # column by column, left to right, fp32
            s = s + sq[:, c]
        dist = np.minimum(dist, s)
        cur = int(np.argmax(dist))          # first maximum
        idx[j] = cur
    return idx


def voxelize(pos, start, end, size):
    """reference utils.py:446-454.  pos [n, D], start / end / size [D] (CPU fp32 tensors) -> (labels [n, 1] int64, centers [m, D]).
    Centres: per axis `arange(start + size/2, end + size/2, size)` (fp32 tensor arithmetic for the bounds, torch's own arange
    for the steps), enumerated with the first axis fastest.  Labels: index of the nearest centre with the reference's kNN
    arithmetic and tie order (oracle/knn_ref.c mode 1 = knn.py:40-52)."""
    import torch
    from oracle import knn_ref
    D = len(size)
    axes = [torch.arange(start[i] + size[i] / 2, end[i] + size[i] / 2, size[i]) for i in range(D)]
    centers = torch.empty(int(np.prod([len(a) for a in axes])), D)
    row = 0
    for multi in np.ndindex(*[len(a) for a in axes[::-1]]):      # last axis slowest, first axis fastest
        for d in range(D):
            centers[row, d] = axes[d][multi[D - 1 - d]]
        row += 1
    labels = knn_ref.knn_query(pos, centers, 1)
    return labels, centers
