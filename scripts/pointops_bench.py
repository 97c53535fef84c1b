"""Device time of the pointops entry points vs the reference's own kernels (oracle/_ref, compiled from the reference's
unmodified sources for sm_100a), on the GPU box:   python scripts/pointops_bench.py > profiles/r02_pointops_vs_reference.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def _time(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


def table(B=128):
    from flowcompare_b200 import pointops as fpo
    from oracle import pointops_refcuda as ref
    g = torch.Generator().manual_seed(0)
    rows = []

    def tref(fn):   # the reference launchers run on the legacy default stream = torch's default stream: no host syncs while timing
        ref.SYNC = False
        try:
            return round(_time(fn), 1)
        finally:
            ref.SYNC = True
    for (n, m) in ((1250, 312), (312, 78), (78, 19), (19, 4)):
        x = (torch.rand(B, n, 3, generator=g) * 2 - 1).cuda()
        a, b = fpo.furthestsampling(x, m), ref.furthestsampling(x, m)
        rows.append({"op": "fps", "B": B, "n": n, "m": m, "ours_us": round(_time(lambda: fpo.furthestsampling(x, m)), 1),
                     "reference_us": tref(lambda: ref.furthestsampling(x, m)), "match": bool(torch.equal(a, b))})
        q = torch.gather(x, 1, a.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        a2, b2 = fpo.knnquery_heap(32, x, q), ref.knnquery_heap(32, x, q)[0]
        rows.append({"op": "knn_heap k=32", "B": B, "n": n, "m": m, "ours_us": round(_time(lambda: fpo.knnquery_heap(32, x, q)), 1),
                     "reference_us": tref(lambda: ref.knnquery_heap(32, x, q)), "match": bool(torch.equal(a2, b2))})
        a3, b3 = fpo.nearestneighbor(x, q)[1], ref.nearestneighbor(x, q)[1]
        rows.append({"op": "three_nn", "B": B, "n": n, "m": m, "ours_us": round(_time(lambda: fpo.nearestneighbor(x, q)), 1),
                     "reference_us": tref(lambda: ref.nearestneighbor(x, q)), "match": bool(torch.equal(a3, b3))})
    return rows


if __name__ == "__main__":
    print("# device time per call (us, CUDA events around 5 calls incl. the wrappers' output allocation), B = 128 clouds")
    for r in table():
        print(r)
