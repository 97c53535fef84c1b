"""Times the library's GEMM kernels (FFMA vs tcgen05 3xTF32) on the flow's shapes (run on the GPU box)."""
import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flowcompare_b200 import lib as fclib, packing

lib = fclib.load()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
shapes = [(150, 256, 1), (256, 256, 1), (256, 64, 0), (64, 128, 0), (214, 512, 1), (512, 512, 1), (512, 300, 0), (300, 300, 0)]
st = torch.cuda.current_stream().cuda_stream
for K, N, act in shapes:
    A = torch.randn(M, (K + 3) // 4 * 4, device="cuda")
    W = torch.randn(N, K) / math.sqrt(K)
    Wt = torch.zeros(packing.gemm_kpad(K), packing.gemm_ldw(N)); Wt[:K, :N] = W.t()
    rows, ldk = packing.tc_n_tiles(N) * packing.tc_bn(N), packing.tc_kpad(K)
    W32 = torch.zeros(rows, ldk); W32[:N, :K] = W
    hi = packing.tf32_round(W32); lo = packing.tf32_round(W32 - hi)
    h16, l16 = packing.f16_split(W32)
    Wt, hi, lo, h16, l16 = Wt.cuda(), hi.cuda(), lo.cuda(), h16.cuda(), l16.cuda()
    b = torch.randn(N, device="cuda")
    C = torch.empty(M, N, device="cuda")
    res = {}
    for name in ("ffma", "tc", "f16"):
        def run():
            if name == "ffma":
                return lib.fc_gemm(A.data_ptr(), A.shape[1], Wt.data_ptr(), Wt.shape[1], b.data_ptr(), C.data_ptr(), N, M, N, K, act, 0, st)
            if name == "f16":
                return lib.fc_gemm_f16x3(A.data_ptr(), A.shape[1], h16.data_ptr(), l16.data_ptr(), ldk, b.data_ptr(), C.data_ptr(), N, M, N, K, act, st)
            return lib.fc_gemm_tf32x3(A.data_ptr(), A.shape[1], hi.data_ptr(), lo.data_ptr(), ldk, b.data_ptr(), C.data_ptr(), N, M, N, K, act, st)
        for _ in range(3): assert run() == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        res[name] = (ms, 2.0 * M * N * K / ms / 1e9)
        if name != "ffma":
            res[name + "_err"] = (C.double() - (A[:, :K].double() @ W.double().cuda().t() + b.double())).abs().max().item() if act == 0 else float("nan")
    print(f"M={M} K={K:4d} N={N:4d} act={act}  ffma {res['ffma'][0]*1e3:8.1f} us {res['ffma'][1]:6.1f} TF/s | tf32x3 {res['tc'][0]*1e3:8.1f} us {res['tc'][1]:6.1f} TF/s err {res['tc_err']:.2e} | f16x3 {res['f16'][0]*1e3:8.1f} us {res['f16'][1]:6.1f} TF/s err {res['f16_err']:.2e}", flush=True)
