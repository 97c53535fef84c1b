"""Experiment: does running two half-batches on two streams (two host threads) beat one full batch?"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flowcompare_b200 import configs, engine, spec

torch.set_grad_enabled(False)
cfg = configs.get_config("dgcnn_attn")
fsd, esd = spec.random_state_dicts(cfg, seed=0)
nstream = int(sys.argv[1]) if len(sys.argv) > 1 else 2
Btot = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = 4
engs = [engine.FlowCompareB200((fsd, esd), cfg, device="cuda:0", precision="tf32x3") for _ in range(nstream)]
streams = [torch.cuda.Stream() for _ in range(nstream)]
Bh = Btot // nstream
batches = []
for i in range(nstream):
    b = spec.synthetic_batch(cfg, Bh, seed=i)
    batches.append((b["extract_0"].cuda(), b["extract_1"].cuda(), b["eps"].cuda()))
torch.cuda.synchronize()


def work(i, n):
    with torch.cuda.stream(streams[i]):
        e0, e1, eps = batches[i]
        for _ in range(n):
            engs[i].inner_loop((e0, e1, None), eps=eps)


def run(n):
    ths = [threading.Thread(target=work, args=(i, n)) for i in range(nstream)]
    for t in ths: t.start()
    for t in ths: t.join()
    torch.cuda.synchronize()


run(2)
t0 = time.perf_counter()
run(steps)
dt = time.perf_counter() - t0
print(f"streams={nstream} B_total={Btot} ms/step={dt / steps * 1e3:.1f} pairs/s={Btot * steps / dt:.1f}", flush=True)
