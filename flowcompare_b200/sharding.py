"""Multi-GPU evaluation: cloud pairs are independent, so they are sharded across ranks with replicated
weights and NO collective on the data path; one gather of per-cloud scalars ends a step.

`evaluate_sharded` is what `test_flow.evaluate_on_test` (reference test_flow.py:135-237) does per batch --
score the pairs, reduce to the running "nats" (= bpd, test_flow.py:224-226) -- spread over
`torch.distributed` ranks.  The scoring callable is injected (the CUDA engine on a GPU box, a stand-in in the
CPU gloo test), so the host logic is testable without a device.
"""
import math

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous block partition: the first (n_items % world) ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def evaluate_sharded(score_fn, extract_0, extract_1, extra_context, eps, input_dim=6, group=None):
    """Scores this rank's block of pairs and gathers per-cloud mean log-prob on every rank.

    score_fn((e0, e1, extra), eps) -> log_prob [b, N] for the local pairs (e.g. `engine.inner_loop(...)[1]`).
    Returns (per_cloud_mean_log_prob [B_total] in global pair order, nats) where nats = bpd over all pairs
    = -mean(log_prob) * log2(e) / input_dim (reference model_initialization.py:225-227).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = extract_0.shape[0]
    lo, hi = shard_bounds(B, rank, world)
    sl = slice(lo, hi)
    if hi > lo:
        lp = score_fn((extract_0[sl], extract_1[sl], None if extra_context is None else extra_context[sl]), eps[sl])
        local = lp.mean(dim=1).to(torch.float32)
    else:
        local = torch.zeros(0, dtype=torch.float32, device=extract_0.device)
    if world == 1:
        per_cloud = local
    else:
        # ragged blocks: pad to the largest block, gather, trim
        width = shard_bounds(B, 0, world)[1]
        buf = torch.zeros(width, dtype=torch.float32, device=local.device)
        buf[: local.numel()] = local
        out = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(out, buf, group=group)
        parts = []
        for r in range(world):
            a, b = shard_bounds(B, r, world)
            parts.append(out[r][: b - a])
        per_cloud = torch.cat(parts)
    nats = -per_cloud.double().mean().item() * math.log2(math.e) / input_dim
    return per_cloud, nats
