"""Per-CTA phase cycle counts of the persistent tcgen05 GEMM.  Needs a timers build:
    scripts/build_variant.sh timers gemm_tc "-DTC_PHASE_TIMERS=1"
    FLOWCOMPARE_B200_LIB=flowcompare_b200/variants/lib_timers.so python scripts/tc_phases.py"""
import ctypes, math, os, sys
os.environ["FC_TC_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flowcompare_b200 import lib as fclib, packing

lib = fclib.load()
lib.fc_debug_tc_phases.argtypes = [ctypes.c_void_p]
shapes = [(65536, 512, 512, 1), (65536, 512, 512, 0), (65536, 256, 256, 1), (65536, 256, 256, 0), (65536, 256, 150, 1), (65536, 256, 150, 0), (65536, 300, 300, 0)]
for (M, N, K, act) in shapes:
    g = torch.Generator().manual_seed(0)
    A = torch.randn(M, (K + 3) // 4 * 4, generator=g).cuda()
    rows, ldk = packing.tc_n_tiles(N) * packing.tc_bn(N), packing.tc_kpad(K)
    W32 = torch.zeros(rows, ldk); W32[:N, :K] = torch.randn(N, K, generator=g) / math.sqrt(K)
    F16 = os.environ.get("FC_FMT", "fp16") == "fp16"
    if F16:
        hi, lo = packing.f16_split(W32)
    else:
        hi = packing.tf32_round(W32); lo = packing.tf32_round(W32 - hi)
    hi, lo, b = hi.cuda(), lo.cuda(), torch.randn(N, generator=g).cuda()
    C = torch.empty(M, N, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    def run():
        rc = (lib.fc_gemm_f16x3 if F16 else lib.fc_gemm_tf32x3)(A.data_ptr(), A.shape[1], hi.data_ptr(), lo.data_ptr(), ldk, b.data_ptr(), C.data_ptr(), N, M, N, K, act, st)
        assert rc == 0
    for _ in range(3): run()
    torch.cuda.synchronize()
    out = (ctypes.c_ulonglong * 16)()
    lib.fc_debug_tc_phases(out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    rc = lib.fc_debug_tc_phases(out)
    n = max(1, out[0])
    tiles = ((M + 127) // 128) * packing.tc_n_tiles(N)
    kb = tiles * ((K + 31) // 32) / n
    print(f"{'f16x3' if F16 else 'tf32x3'} M={M} N={N} K={K} act={act} ctas={out[0]} rc={rc} us={e0.elapsed_time(e1)*1e3:.1f} k-blocks/CTA={kb:.0f} | MMA warp: total {out[1]/n:.0f} "
          f"(/kb {out[1]/n/kb:.0f}) wait full {out[2]/n:.0f} conv {out[3]/n:.0f} acc_free {out[4]/n:.0f} | converter: total {out[5]/n:.0f} wait full {out[6]/n:.0f} "
          f"tfree {out[7]/n:.0f} | MMA probe {out[10]/n:.0f} issue {out[11]/n:.0f} | epilogue: total {out[8]/n:.0f} wait acc_full {out[9]/n:.0f}", flush=True)
