"""BASELINE.json configs[4]: point-count sweep (architecture dgcnn_attn, Nc = N in {2k..}) on one GPU.
Prints one JSON line per point count: pairs/s (CUDA events), per-class kernel times of one instrumented step."""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flowcompare_b200 import configs, engine, lib, spec

torch.set_grad_enabled(False)
points = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [2048, 4096, 8192]
clib = lib.load()
base = configs.get_config("dgcnn_attn")
fsd, esd = spec.random_state_dicts(base, seed=0)
eng = None
names = ["gemm_fp32_ffma", "gemm_tf32x3_tcgen05", "cross_attention", "knn", "edgeconv_gather_max", "other"]
for n in points:
    cfg = configs.get_config("dgcnn_attn", sample_size=n, n_samples_context=n)
    if eng is None:
        eng = engine.FlowCompareB200((fsd, esd), cfg, device="cuda:0", precision="tf32x3")
    B = max(1, min(16, 65536 // n))
    batch = spec.synthetic_batch(cfg, B, seed=n)
    e0, e1, eps = batch["extract_0"].cuda(), batch["extract_1"].cuda(), batch["eps"].cuda()
    for _ in range(2):
        eng.inner_loop((e0, e1, None), eps=eps)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 3
    ev0.record()
    for _ in range(steps):
        loss, lp, bpd = eng.inner_loop((e0, e1, None), eps=eps)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    clib.fc_profile_begin()
    eng.inner_loop((e0, e1, None), eps=eps)
    a, b, c = (ctypes.c_double * 6)(), (ctypes.c_double * 6)(), (ctypes.c_double * 6)()
    d = (ctypes.c_int64 * 6)()
    clib.fc_profile_end(a, b, c, d, 6)
    print(json.dumps({"n_points": n, "pairs_per_step": B, "ms_per_step": round(ms, 2), "pairs_per_s": round(B / ms * 1e3, 3),
                      "finite": bool(torch.isfinite(lp).all()), "mean_log_prob": float(lp.mean()),
                      "class_ms": {names[i]: round(a[i], 2) for i in range(6) if d[i]}}), flush=True)
