"""Where does the fp32 noise of a full-depth forward come from?  (CPU, oracle port; authoring container or GPU-box host)

Runs oracle/port.py on a full 115-layer fixture in fp64 (truth), in fp32, and in fp64 with ONE op family demoted to
fp32 at a time.  Prints max / mean |log_prob - fp64| per variant.  Used to decide where the CUDA path must be more
careful than plain fp32."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from flowcompare_b200 import configs
from oracle import port
from oracle.make_golden import fixture_inputs

torch.set_grad_enabled(False)
name = sys.argv[1] if len(sys.argv) > 1 else "full_dgcnn_attn"
cfg, fsd, esd, batch = fixture_inputs(name)
dcfg = configs.derive(cfg)
ctx32, _ = port.dgcnn_embed(esd, batch["extract_0"], cfg["n_neighbors"])
ex = batch["extra_context"]
N = batch["extract_1"].shape[1]
ex3 = None if ex is None else ex.unsqueeze(1).expand(-1, N, -1)

def run(dtype, patch=None):
    sd = port.to_dtype(fsd, dtype)
    saved = {}
    if patch:
        for k, v in patch.items():
            saved[k] = getattr(*k)
            setattr(*k, v)
    try:
        t0 = time.time()
        out = port.flow_log_prob(sd, dcfg, batch["extract_1"].to(dtype), ctx32.to(dtype), None if ex3 is None else ex3.to(dtype), batch["eps"].to(dtype))
        return out.double(), time.time() - t0
    finally:
        for k, v in saved.items():
            setattr(*k, v)

truth, t = run(torch.float64)
print("fp64 s", t)
f32, t = run(torch.float32)
d = (f32 - truth).abs()
print(f"all fp32: max {d.max():.3e} mean {d.mean():.3e}  ({t:.1f}s)")

def demote(fn):
    def w(*a, **k):
        a2 = [x.float() if torch.is_tensor(x) and x.is_floating_point() else x for x in a]
        k2 = {kk: (x.float() if torch.is_tensor(x) and x.is_floating_point() else x) for kk, x in k.items()}
        return fn(*a2, **k2).double()
    return w

variants = {
    "linear(fp32)": {(F, "linear"): demote(F.linear)},
    "gelu(fp32)": {(F, "gelu"): demote(F.gelu)},
    "matmul+softmax(fp32)": {(torch, "matmul"): demote(torch.matmul), (torch, "softmax"): demote(torch.softmax)},
    "layer_norm(fp32)": {(F, "layer_norm"): demote(F.layer_norm)},
    "sigmoid/log/exp(fp32)": {(torch, "sigmoid"): demote(torch.sigmoid), (torch, "log"): demote(torch.log), (torch, "exp"): demote(torch.exp)},
}
for nm, p in variants.items():
    o, t = run(torch.float64, p)
    d = (o - truth).abs()
    print(f"{nm}: max {d.max():.3e} mean {d.mean():.3e}")
