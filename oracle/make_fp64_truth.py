"""TEST INFRASTRUCTURE ONLY -- fp64 "truth" for the full-depth fixtures.

    python -m oracle.make_fp64_truth            # writes tests/golden/<fixture>_fp64.pt

The reference's fp32 forward is itself 1.4e-3 .. 2e-3 nats (max over 1024 points) away from an exact evaluation
of the same model at 115 layers (scripts/precision_ablation.py: the fp32 GEMMs dominate), so "within 1e-3 of the
reference's fp32 output" cannot separate a wrong kernel from a differently-rounded one.  This script evaluates the
SAME model (same seeded weights, inputs and injected eps as the fixture) with oracle/port.py in float64 -- the
port is pinned against the live reference in tests/test_oracle.py -- and stores the per-point log-prob.  The discrete
decisions of the embedder (kNN / FPS indices) are taken from the fp32 evaluation, exactly as the reference takes
them, so the fp64 run follows the same graph.  Needs no reference tree (pure port), runs in ~10 s per fixture.
"""
import os
import sys

import torch

from flowcompare_b200 import configs
from oracle import port
from oracle.make_golden import FIXTURES, GOLDEN_DIR, fixture_inputs


def fp64_log_prob(name):
    cfg, fsd, esd, batch = fixture_inputs(name)
    dcfg = configs.derive(cfg)
    f64, e64 = port.to_dtype(fsd, torch.float64), port.to_dtype(esd, torch.float64)
    e0, e1 = batch["extract_0"][:, :, :6], batch["extract_1"][:, :, :6]
    N = e1.shape[1]
    kind = dcfg["input_embedder"]
    if kind == "PAConv":
        from oracle import port_paconv
        ctx = port_paconv.paconv_embed(e64, e0.double(), dcfg)      # index ops run on the (exact) fp32 coordinates
    else:
        fn = port.dgcnn_embed_global if kind == "DGCNNembedderGlobal" else port.dgcnn_embed
        _, idxs = fn(esd, e0, cfg["n_neighbors"])                   # fp32 neighbour choices, as the reference makes them
        ctx, _ = fn(e64, e0.double(), cfg["n_neighbors"], idx_list=idxs)
        if kind == "DGCNNembedderGlobal":
            ctx = ctx.unsqueeze(1).expand(-1, N, -1)
    ex = batch["extra_context"]
    ex = None if ex is None else ex.double().unsqueeze(1).expand(-1, N, -1)
    return port.flow_log_prob(f64, dcfg, e1.double(), ctx, ex, batch["eps"].double())


def main(argv):
    torch.set_grad_enabled(False)
    names = [n for n in FIXTURES if n.startswith("full") or n.startswith("mid")]
    for name in names:
        if len(argv) > 1 and argv[1] not in name:
            continue
        lp = fp64_log_prob(name)
        path = os.path.join(GOLDEN_DIR, f"{name}_fp64.pt")
        torch.save({"log_prob": lp.clone(), "meta": {"fixture": name, "generator": "oracle/make_fp64_truth.py",
                                                     "source": "oracle/port.py in float64"}}, path)
        gold = torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False)
        d = (gold["log_prob"].double() - lp).abs()
        print(f"{name}: |reference fp32 - fp64| max {d.max().item():.3e} mean {d.mean().item():.3e} -> {path}")


if __name__ == "__main__":
    main(sys.argv)
