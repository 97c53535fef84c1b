"""CPU tests of the boundary: the library loads, exports every symbol the header declares, and the
ctypes signatures cover the header (no compute calls without a GPU)."""
import os
import re

from flowcompare_b200 import lib as fclib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "flowcompare_b200.h")).read()
    return sorted(set(re.findall(r"^FC_API [\w\s\*]+?\b(fc_\w+)\(", src, flags=re.M)))


def test_library_exports_every_declared_symbol():
    lib = fclib.load()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_ctypes_signatures_cover_header():
    assert sorted(fclib.SIGNATURES) == _header_symbols()


def test_version_and_error_string_callable_without_gpu():
    lib = fclib.load()
    assert lib.fc_version() >= 100
    assert isinstance(lib.fc_last_error(), bytes)
    assert lib.fc_launch_count() >= 0


def test_create_rejects_bad_header():
    import ctypes
    import numpy as np
    lib = fclib.load()
    header = np.zeros(19, dtype=np.int32)
    table = np.zeros(4, dtype=np.int64)
    arena = np.zeros(64, dtype=np.float32)
    h = ctypes.c_void_p()
    rc = lib.fc_flow_create(header.ctypes.data, 19, table.ctypes.data, 4, arena.ctypes.data, 64, h)
    assert rc == -6  # FC_ERR_MODEL (bad magic); nothing touches the device


def test_bench_reads_traffic_from_committed_profile():
    """bench.py's roofline.traffic comes from the committed ncu capture; make sure the file is there and parses."""
    import bench
    traffic, note = bench.ncu_traffic()
    assert isinstance(traffic, int) and traffic > 0, note


def test_bench_reads_the_measured_ffma_peak():
    import bench
    v = bench.ffma_peak()
    assert isinstance(v, float) and 50.0 < v < 80.0      # 148 SMs x 128 lanes x 2 flop x <= 1.965 GHz = 74.5 TFLOP/s nominal


def test_argument_validation_happens_before_any_gpu_work():
    """Invalid arguments are rejected with FC_ERR_INVALID_ARG / FC_ERR_UNSUPPORTED on the host, before a single CUDA
    call: these run without a GPU."""
    lib = fclib.load()
    INVALID, UNSUPPORTED = -1, -5
    assert lib.fc_knn_self(0, 6, 1, 10, 6, 4, 0, 0, 0) == INVALID                       # null input
    assert lib.fc_knn_self_ws(0, 6, 1, 10, 6, 4, 0, 0, 0, 0, 0) == INVALID               # null input (workspace form)
    assert lib.fc_knn_self_ws(8, 6, 1, 10, 6, 11, 8, 0, 8, 1 << 20, 0) == INVALID         # k > N
    assert lib.fc_knn_self_ws(8, 6, 1, 10, 6, 4, 8, 0, 0, 0, 0) == -4                     # FC_ERR_WORKSPACE: no workspace
    assert lib.fc_knn_self_ws(8, 6, 1, 10, 6, 4, 8, 0, 256, 16, 0) == -4                  # FC_ERR_WORKSPACE: too small
    assert lib.fc_knn_query_ws(8, 8, 10, 5, 3, 6, 8, 8, 1 << 20, 0) == INVALID            # k > Nt
    assert lib.fc_knn_query_ws(8, 8, 10, 5, 3, 2, 8, 0, 0, 0) == -4
    # the workspace covers the norms and the keys of one chunk of clouds, 256-byte slack for alignment
    assert lib.fc_knn_workspace_bytes(1, 1000, 315, 0) >= 4 * (1000 + 315 + 1000 * 316)
    assert lib.fc_knn_workspace_bytes(2, 1250, 1250, 1) >= 4 * (2 * 1250 + 2 * 1250 * 1252)
    assert lib.fc_knn_workspace_bytes(0, 10, 10, 1) == 0
    assert lib.fc_cross_attention(0, 64, 0, 128, 0, 64, 1, 8, 8, 64, 0.125, 0) == INVALID
    assert lib.fc_cross_attention_tf32x3(0, 64, 0, 128, 0, 64, 1, 8, 8, 64, 0.125, 0) == INVALID
    assert lib.fc_cross_attention_tc(0, 64, 0, 128, 0, 64, 1, 8, 8, 64, 0.125, 0, 0, 0) == INVALID   # no scratch
    assert lib.fc_cross_attention_tc_scratch_bytes(0, 8) == INVALID
    # 2 clouds x 1250 keys: hi/lo copies of k [B*Nc, 64] and of v^T [B, 64, round4(Nc)]
    assert lib.fc_cross_attention_tc_scratch_bytes(2, 1250) >= 4 * (2 * 2 * 1250 * 64 + 2 * 2 * 64 * 1252)
    assert lib.fc_flow_workspace_bytes(0, 1, 8, 8) == INVALID
    assert lib.fc_gemm(0, 0, 0, 0, 0, 0, 0, 4, 4, 4, 0, 0, 0) == INVALID
    msg = lib.fc_last_error()
    assert isinstance(msg, bytes)


def _header_prototypes():
    """{name: (return C type, [parameter C types])} parsed from the FC_API prototypes of the header."""
    src = open(os.path.join(ROOT, "include", "flowcompare_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    out = {}
    for m in re.finditer(r"FC_API\s+([\w\s\*]+?)\b(fc_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        types = []
        if params and params != "void":
            for p in params.split(","):
                p = " ".join(p.split())
                p = re.sub(r"\s*\b\w+$", "", p) if not p.endswith("*") else p      # drop the parameter name
                types.append(p.replace(" *", "*"))
        out[name] = (ret, types)
    return out


def _ctype_of(c_type):
    import ctypes
    t = c_type.replace("const ", "").strip()
    if t.endswith("*") and t.replace(" ", "") == "char*":
        return ctypes.c_char_p
    if t.endswith("**"):
        return ctypes.POINTER(ctypes.c_void_p)      # out-handles (fc_flow**, fc_embedder**)
    if t.endswith("*") or t in ("fc_stream_t",):
        return ctypes.c_void_p
    return {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "float": ctypes.c_float, "double": ctypes.c_double,
            "uint64_t": ctypes.c_uint64, "unsigned long long": ctypes.c_uint64, "long long": ctypes.c_int64}[t]


def test_ctypes_argument_types_match_the_header_one_by_one():
    """Every entry of lib.SIGNATURES has the header prototype's return type and, position by position, its parameter types
    (pointers and fc_stream_t -> c_void_p, int -> c_int, int64_t -> c_int64, float -> c_float ...): a drifted binding would
    pass wrong-sized arguments without any error on the GPU box."""
    protos = _header_prototypes()
    assert sorted(protos) == sorted(fclib.SIGNATURES)
    for name, (res, args) in fclib.SIGNATURES.items():
        ret, types = protos[name]
        assert (res is None) if ret == "void" else (_ctype_of(ret) is res), (name, ret, res)
        want = [_ctype_of(t) for t in types]
        assert want == list(args), (name, types, args)
