"""Authoring-container only: times the UNMODIFIED reference inner_loop (oracle/refload.py) next to oracle/port.py on
the same weights / inputs / eps, to show that the port used as the CPU baseline on the GPU box runs at the
reference's own speed.  (SURVEY.md 8d "CPU baseline timing": warm-up 1, best of 3, all host threads.)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flowcompare_b200 import configs, spec
from oracle import port, make_golden

torch.set_grad_enabled(False)
torch.set_num_threads(os.cpu_count())
label = sys.argv[1] if len(sys.argv) > 1 else "dgcnn_attn"
cfg = configs.get_config(label)
fsd, esd = spec.random_state_dicts(cfg, seed=0)
from oracle import refload
import torch.distributions.normal as tdn
models, mi = refload.load()
torch.manual_seed(0)
md = mi.initialize_flow(dict(cfg), "cpu", "test")          # built once; only inner_loop is timed
md["flow"].load_state_dict(fsd)
md["input_embedder"].load_state_dict(esd)
dcfg = configs.derive(cfg)
for B in (1, 4):
    batch = spec.synthetic_batch(cfg, B, seed=1)
    def ref():
        orig = tdn._standard_normal
        tdn._standard_normal = lambda shape, dtype, device: batch["eps"].to(dtype).reshape(shape)
        try:
            return mi.inner_loop((batch["extract_0"], batch["extract_1"], batch["extra_context"]), md, dcfg)
        finally:
            tdn._standard_normal = orig
    def prt():
        return port.inner_loop((batch["extract_0"], batch["extract_1"], batch["extra_context"]), fsd, esd, dcfg, batch["eps"])
    res = {}
    for name, fn in (("reference", ref), ("port", prt)):
        fn()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); out = fn(); best = min(best, time.perf_counter() - t0)
        res[name] = B / best
    print(f"{label} B={B} threads={torch.get_num_threads()}: reference {res['reference']:.2f} pairs/s, port {res['port']:.2f} pairs/s", flush=True)
