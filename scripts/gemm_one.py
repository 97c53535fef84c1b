"""One tcgen05 GEMM shape, few launches (for ncu)."""
import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flowcompare_b200 import lib as fclib, packing
lib = fclib.load()
M, K, N, act = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
st = torch.cuda.current_stream().cuda_stream
A = torch.randn(M, (K + 3) // 4 * 4, device="cuda")
W = torch.randn(N, K) / math.sqrt(K)
rows, ldk = packing.tc_n_tiles(N) * packing.tc_bn(N), packing.tc_kpad(K)
W32 = torch.zeros(rows, ldk); W32[:N, :K] = W
F16 = os.environ.get("FC_FMT", "fp16") == "fp16"
hi, lo = packing.f16_split(W32) if F16 else (packing.tf32_round(W32), packing.tf32_round(W32 - packing.tf32_round(W32)))
hi, lo = hi.cuda(), lo.cuda()
fn = lib.fc_gemm_f16x3 if F16 else lib.fc_gemm_tf32x3
b = torch.randn(N, device="cuda"); C = torch.empty(M, N, device="cuda")
for _ in range(4):
    assert fn(A.data_ptr(), A.shape[1], hi.data_ptr(), lo.data_ptr(), ldk, b.data_ptr(), C.data_ptr(), N, M, N, K, act, st) == 0
torch.cuda.synchronize()
print("ok")
