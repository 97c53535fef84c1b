"""TEST INFRASTRUCTURE ONLY -- the REFERENCE's own compiled pointops CUDA kernels, driven through ctypes.

`oracle/_ref/libpointops_ref.so` is built by oracle/build_ref_pointops.sh from the reference's unmodified
`lib/pointops/src/*/*_cuda_kernel.cu` (kernels + extern "C" raw-pointer launchers).  The functions below allocate
outputs exactly as the reference's Python wrappers do (lib/pointops/functions/pointops.py:47-62, :96-115, :121-140,
:158-175, :475-497) and call those launchers, so on a GPU box they ARE the reference kernels: the bit-exact pin for
the index ops of flowcompare_b200/csrc/paconv.cu and the "kernel to beat" for timing.  CUDA tensors only.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpointops_ref.so")
_lib = None
SYNC = True     # the launchers use the legacy default stream: synchronise around them unless a benchmark turns it off


def available() -> bool:
    return os.path.exists(LIB_PATH)


def load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(LIB_PATH)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        lib.furthestsampling_cuda_launcher.argtypes = [ci, ci, ci, vp, vp, vp]
        lib.gathering_forward_cuda_launcher.argtypes = [ci, ci, ci, ci, vp, vp, vp]
        lib.knnquery_heap_cuda_launcher.argtypes = [ci, ci, ci, ci, vp, vp, vp, vp, vp]
        lib.nearestneighbor_cuda_launcher_fast.argtypes = [ci, ci, ci, vp, vp, vp, vp]
        lib.interpolation_forward_cuda_launcher_fast.argtypes = [ci, ci, ci, ci, vp, vp, vp, vp]
        lib.grouping_forward_cuda_launcher_fast.argtypes = [ci, ci, ci, ci, ci, vp, vp, vp]
        for name in ("furthestsampling_cuda_launcher", "gathering_forward_cuda_launcher", "knnquery_heap_cuda_launcher",
                     "nearestneighbor_cuda_launcher_fast", "interpolation_forward_cuda_launcher_fast",
                     "grouping_forward_cuda_launcher_fast"):
            getattr(lib, name).restype = None
        _lib = lib
    return _lib


def _chk(*ts):
    for t in ts:
        assert t.is_cuda and t.is_contiguous()


def furthestsampling(xyz, m):
    """K2 (sampling_cuda_kernel.cu:58-209): xyz [B,n,3] -> idx [B,m] int32.  Launches on the legacy default stream."""
    _chk(xyz)
    b, n, _ = xyz.shape
    idx = torch.empty(b, m, dtype=torch.int32, device=xyz.device)
    temp = torch.full((b, n), 1e10, dtype=torch.float32, device=xyz.device)
    if SYNC:
        torch.cuda.synchronize()
    load().furthestsampling_cuda_launcher(b, n, m, xyz.data_ptr(), temp.data_ptr(), idx.data_ptr())
    if SYNC:
        torch.cuda.synchronize()
    return idx


def gathering(features, idx):
    """K3 (sampling_cuda_kernel.cu:6-41): features [B,c,n], idx [B,m] -> [B,c,m]."""
    _chk(features, idx)
    b, c, n = features.shape
    m = idx.shape[1]
    out = torch.empty(b, c, m, dtype=torch.float32, device=features.device)
    if SYNC:
        torch.cuda.synchronize()
    load().gathering_forward_cuda_launcher(b, c, n, m, features.data_ptr(), idx.data_ptr(), out.data_ptr())
    if SYNC:
        torch.cuda.synchronize()
    return out


def knnquery_heap(nsample, xyz, new_xyz):
    """K1 (knnquery_heap_cuda_kernel.cu:53-110): -> (idx [B,m,nsample] int32, dist2 [B,m,nsample])."""
    _chk(xyz, new_xyz)
    b, m, _ = new_xyz.shape
    n = xyz.shape[1]
    idx = torch.zeros(b, m, nsample, dtype=torch.int32, device=xyz.device)
    dist2 = torch.zeros(b, m, nsample, dtype=torch.float32, device=xyz.device)
    load().knnquery_heap_cuda_launcher(b, n, m, nsample, xyz.data_ptr(), new_xyz.data_ptr(), idx.data_ptr(), dist2.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream)
    if SYNC:
        torch.cuda.synchronize()
    return idx, dist2


def nearestneighbor(unknown, known):
    """K5 (interpolation_cuda_kernel.cu:134-213): -> (dist2 [B,n,3] SQUARED distances, idx [B,n,3] int32).
    (The reference's Python wrapper returns sqrt(dist2), pointops.py:112.)"""
    _chk(unknown, known)
    b, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.empty(b, n, 3, dtype=torch.float32, device=unknown.device)
    idx = torch.empty(b, n, 3, dtype=torch.int32, device=unknown.device)
    if SYNC:
        torch.cuda.synchronize()
    load().nearestneighbor_cuda_launcher_fast(b, n, m, unknown.data_ptr(), known.data_ptr(), dist2.data_ptr(), idx.data_ptr())
    if SYNC:
        torch.cuda.synchronize()
    return dist2, idx


def interpolation(features, idx, weight):
    """K6 (interpolation_cuda_kernel.cu:181-229): features [B,c,m], idx/weight [B,n,3] -> [B,c,n]."""
    _chk(features, idx, weight)
    b, c, m = features.shape
    n = idx.shape[1]
    out = torch.empty(b, c, n, dtype=torch.float32, device=features.device)
    if SYNC:
        torch.cuda.synchronize()
    load().interpolation_forward_cuda_launcher_fast(b, c, m, n, features.data_ptr(), idx.data_ptr(), weight.data_ptr(),
                                                    out.data_ptr())
    if SYNC:
        torch.cuda.synchronize()
    return out


def grouping(features, idx):
    """K4 (grouping_cuda_kernel.cu:60-95): features [B,c,n], idx [B,m,k] -> [B,c,m,k]."""
    _chk(features, idx)
    b, c, n = features.shape
    m, k = idx.shape[1], idx.shape[2]
    out = torch.empty(b, c, m, k, dtype=torch.float32, device=features.device)
    if SYNC:
        torch.cuda.synchronize()
    load().grouping_forward_cuda_launcher_fast(b, c, n, m, k, features.data_ptr(), idx.data_ptr(), out.data_ptr())
    if SYNC:
        torch.cuda.synchronize()
    return out
