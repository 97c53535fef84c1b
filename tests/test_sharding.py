"""CPU tests of the multi-GPU host logic: world-size-2 gloo run of `evaluate_sharded` with the oracle port as
the scoring function must reproduce the single-process result (pairs are independent; one gather of
per-cloud scalars is the only collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flowcompare_b200 import configs, sharding, spec
from oracle import port

torch.set_grad_enabled(False)


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 5, 8, 20, 33):
        for w in (1, 2, 3, 4, 8):
            seen = []
            for r in range(w):
                lo, hi = sharding.shard_bounds(n, r, w)
                assert 0 <= lo <= hi <= n
                seen.extend(range(lo, hi))
            assert seen == list(range(n))
            sizes = [sharding.shard_bounds(n, r, w)[1] - sharding.shard_bounds(n, r, w)[0] for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


def _setup(label="dgcnn_attn_extra", pairs=5):
    cfg = configs.get_config(label, n_flow_layers=2, sample_size=48, n_samples_context=64)
    fsd, esd = spec.random_state_dicts(cfg, seed=3)
    batch = spec.synthetic_batch(cfg, pairs, seed=4)   # 5 pairs over 2 ranks: ragged 3 + 2; 2 pairs over 3 ranks: an empty block
    dcfg = configs.derive(cfg)

    def score(args, eps):
        return port.inner_loop(args, fsd, esd, dcfg, eps)[1]

    return cfg, batch, score


def _worker(rank, world, port_no, q, label="dgcnn_attn_extra", pairs=5):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.set_grad_enabled(False)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg, batch, score = _setup(label, pairs)
        per_cloud, nats = sharding.evaluate_sharded(score, batch["extract_0"], batch["extract_1"], batch["extra_context"],
                                                    batch["eps"], input_dim=cfg["input_dim"])
        q.put((rank, per_cloud.tolist(), nats))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world,label,pairs", [(2, "dgcnn_attn_extra", 5), (3, "dgcnn_attn", 2)])
def test_multi_rank_gloo_matches_single_process(world, label, pairs):
    """world 2: ragged blocks (3 + 2 pairs) with extra context; world 3 with 2 pairs: the last rank's block is EMPTY (it still
    takes part in the gather) and the config has no extra context."""
    cfg, batch, score = _setup(label, pairs)
    want, want_nats = sharding.evaluate_sharded(score, batch["extract_0"], batch["extract_1"], batch["extra_context"],
                                                batch["eps"], input_dim=cfg["input_dim"])
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port_no = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, q, label, pairs)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, per_cloud, nats in results:
        # B=1-vs-batched CPU BLAS may differ in the last bits, hence a tolerance and not equality
        assert torch.allclose(torch.tensor(per_cloud), want, atol=1e-4)
        assert abs(nats - want_nats) < 1e-5
