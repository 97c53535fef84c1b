"""Times the 3xFP16 / 3xTF32 tcgen05 GEMM on a list of shapes with the library named by FLOWCOMPARE_B200_LIB (variant A/B runs).
usage: python scripts/gemm_time.py [M] ; prints one line per (K, N, act)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flowcompare_b200 import lib as fclib, packing
lib = fclib.load()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
shapes = [(512, 512, 1), (256, 256, 1), (214, 512, 1), (512, 300, 0), (300, 300, 0), (150, 256, 1)]
st = torch.cuda.current_stream().cuda_stream
F16 = os.environ.get("FC_FMT", "fp16") == "fp16"
out = []
for K, N, act in shapes:
    A = torch.randn(M, (K + 3) // 4 * 4, device="cuda")
    rows, ldk = packing.tc_n_tiles(N) * packing.tc_bn(N), packing.tc_kpad(K)
    W32 = torch.zeros(rows, ldk); W32[:N, :K] = torch.randn(N, K) / math.sqrt(K)
    hi, lo = packing.f16_split(W32) if F16 else (packing.tf32_round(W32), packing.tf32_round(W32 - packing.tf32_round(W32)))
    hi, lo = hi.cuda(), lo.cuda()
    fn = lib.fc_gemm_f16x3 if F16 else lib.fc_gemm_tf32x3
    b = torch.randn(N, device="cuda"); C = torch.empty(M, N, device="cuda")
    run = lambda: fn(A.data_ptr(), A.shape[1], hi.data_ptr(), lo.data_ptr(), ldk, b.data_ptr(), C.data_ptr(), N, M, N, K, act, st)
    for _ in range(3): assert run() == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    kb = ((M + 127) // 128) * packing.tc_n_tiles(N) * ((K + 31) // 32) / 148
    out.append(f"K{K}N{N}: {us:6.1f}us {2.0*M*N*K/us/1e6:5.0f}TF {us*1e-6*1.9e9/kb:5.0f}cyc/kb")
print(os.environ.get("FLOWCOMPARE_B200_LIB", "main").split("lib_")[-1], " | ".join(out), flush=True)
