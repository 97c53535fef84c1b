"""Debug helper: one tcgen05 GEMM through the C ABI vs fp64 (run on the GPU box)."""
import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flowcompare_b200 import lib as fclib, packing

M, N, K = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else (256, 256, 64)
act = int(sys.argv[4]) if len(sys.argv) > 4 else 0
lib = fclib.load()
g = torch.Generator().manual_seed(0)
A = torch.randn(M, K, generator=g)
W = torch.randn(N, K, generator=g) / math.sqrt(K)
b = torch.randn(N, generator=g)
lda = (K + 3) // 4 * 4
Ad = torch.zeros(M, lda); Ad[:, :K] = A
rows, ldk = packing.tc_n_tiles(N) * packing.tc_bn(N), packing.tc_kpad(K)
W32 = torch.zeros(rows, ldk); W32[:N, :K] = W
hi = packing.tf32_round(W32); lo = packing.tf32_round(W32 - hi)
Ad, hi, lo, bd = Ad.cuda(), hi.cuda(), lo.cuda(), b.cuda()
C = torch.full((M, N), float("nan"), device="cuda")
print("launch", M, N, K, "BN", packing.tc_bn(N), "ldk", ldk, flush=True)
rc = lib.fc_gemm_tf32x3(Ad.data_ptr(), lda, hi.data_ptr(), lo.data_ptr(), ldk, bd.data_ptr(), C.data_ptr(), N, M, N, K, act,
                        torch.cuda.current_stream().cuda_stream)
print("rc", rc, lib.fc_last_error(), flush=True)
torch.cuda.synchronize()
ref = A.double() @ W.double().t() + b.double()
if act == 1: ref = torch.nn.functional.gelu(ref)
d = (C.cpu().double() - ref).abs()
print("max err", d.max().item(), "nan", int(torch.isnan(C).sum().item()), "mean err", d.nanmean().item())
if d.max().item() > 1e-3 or torch.isnan(C).any():
    bad = (d > 1e-3) | torch.isnan(d)
    rows_bad = bad.any(dim=1).nonzero().flatten()[:20].tolist()
    cols_bad = bad.any(dim=0).nonzero().flatten()[:40].tolist()
    print("bad rows (first)", rows_bad, "bad cols (first)", cols_bad, "frac bad", bad.float().mean().item())
    print("C[0,:8]", C[0, :8].tolist()); print("ref[0,:8]", ref[0, :8].tolist())
