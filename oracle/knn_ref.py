"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/knn_ref.c (bit-exact kNN oracle)."""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfc_oracle.so")
_lib = None


def build():
    src = [os.path.join(_HERE, f) for f in sorted(os.listdir(_HERE)) if f.endswith(".c")]
    if os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _SO] + src + ["-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.fc_oracle_knn.restype = ctypes.c_int
        _lib.fc_oracle_knn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return _lib


def knn_self(x, k):
    """x [B,N,C] fp32 (CPU) -> idx [B,N,k] int64; DGCNN form (mode 0)."""
    x = x.detach().to(torch.float32).contiguous()
    B, N, C = x.shape
    out = np.empty((B, N, k), dtype=np.int32)
    xn = x.numpy()
    for b in range(B):
        rc = _load().fc_oracle_knn(xn[b].ctypes.data, C, xn[b].ctypes.data, C, N, N, C, k, 0, out[b].ctypes.data)
        assert rc == 0
    return torch.from_numpy(out.astype(np.int64))


def knn_query(q, t, k):
    """q [Nq,D], t [Nt,D] -> idx [Nq,k] int64; knn.py form (mode 1)."""
    q = q.detach().to(torch.float32).contiguous().numpy()
    t = t.detach().to(torch.float32).contiguous().numpy()
    out = np.empty((q.shape[0], k), dtype=np.int32)
    rc = _load().fc_oracle_knn(q.ctypes.data, q.shape[1], t.ctypes.data, t.shape[1], q.shape[0], t.shape[0],
                               q.shape[1], k, 1, out.ctypes.data)
    assert rc == 0
    return torch.from_numpy(out.astype(np.int64))
