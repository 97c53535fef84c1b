"""Times the three cross-attention kernels alone (B clouds, N queries, Nc keys, d=64) through the C ABI."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flowcompare_b200 import lib as fclib

B, N, Nc = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else (64, 1024, 1250)
lib = fclib.load()
g = torch.Generator().manual_seed(0)
q = (torch.randn(B, N, 64, generator=g) * 2).cuda()
kv = torch.randn(B, Nc, 128, generator=g).cuda()
out = torch.empty(B, N, 64, device="cuda")
nbytes = lib.fc_cross_attention_tc_scratch_bytes(B, Nc)
scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")


def run(kind):
    if kind.startswith("tcgen05"):
        fn = lib.fc_cross_attention_tc_f16 if kind == "tcgen05_f16" else lib.fc_cross_attention_tc
        return fn(q.data_ptr(), 64, kv.data_ptr(), 128, out.data_ptr(), 64, B, N, Nc, 64, 0.125, scratch.data_ptr(), nbytes, st)
    fn = lib.fc_cross_attention_tf32x3 if kind == "mma" else lib.fc_cross_attention
    return fn(q.data_ptr(), 64, kv.data_ptr(), 128, out.data_ptr(), 64, B, N, Nc, 64, 0.125, st)


ref = None
for kind in ("mma", "tcgen05", "tcgen05_f16"):
    for _ in range(3):
        assert run(kind) == 0
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(kind); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    if ref is None: ref = out.clone()
    print(f"{kind:8s} B={B} N={N} Nc={Nc}: {us:8.1f} us  {4.0 * B * N * Nc * 64 / us / 1e6:7.1f} TFLOP/s  max|diff vs mma| {(out - ref).abs().max().item():.2e}", flush=True)

if os.environ.get("FC_ATTN_DEBUG") == "1":
    import ctypes
    lib.fc_debug_attn_phases.argtypes = [ctypes.c_void_p]
    o = (ctypes.c_ulonglong * 16)()
    lib.fc_debug_attn_phases(o)
    run("tcgen05"); torch.cuda.synchronize()
    lib.fc_debug_attn_phases(o)
    n = max(1, o[0]); nb = (Nc + 63) // 64
    names = ["ctas", "mma_total", "w_kfull", "w_sfree", "w_vfull", "w_ofree", "w_pready", "sm_w_sready", "sm_w_pfree", "sm_w_oready", "sm_total", "q_phase"]
    print("per-CTA cycles:", {k: int(o[i] / n) for i, k in enumerate(names) if i}, "blocks", nb, "-> per block", int(o[1] / n / nb))
