"""TEST INFRASTRUCTURE ONLY -- tests/golden/dataops.pt: outputs of the UNMODIFIED reference's utils.co_unit_sphere
(utils.py:259-280) on seeded clouds (python -m oracle.make_dataops_golden; needs /root/reference)."""
import os

import torch

from oracle import refload

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dataops.pt")


def inputs(seed=0, n0=1250, n1=1024):
    g = torch.Generator().manual_seed(4242 + seed)
    p0 = torch.rand(n0, 6, generator=g) * torch.tensor([2.2, 2.2, 4.2, 1, 1, 1]) + torch.tensor([10.0, -3.0, 1.0, 0, 0, 0])
    p1 = torch.rand(n1, 6, generator=g) * torch.tensor([2.0, 2.0, 4.0, 1, 1, 1]) + torch.tensor([10.1, -2.9, 1.1, 0, 0, 0])
    return p0, p1


def main():
    refload.load()
    import utils
    out = {}
    for seed in range(3):
        p0, p1 = inputs(seed)
        a, b, inv = utils.co_unit_sphere(p0.clone(), p1.clone(), return_inverse=True)
        out[seed] = {"points_0": a.clone(), "points_1": b.clone(), "furthest_distance": inv["furthest_distance"].clone(),
                     "mean": inv["mean"].clone()}
    out["meta"] = {"generator": "oracle/make_dataops_golden.py", "source": "unmodified reference utils.co_unit_sphere, CPU fp32"}
    torch.save(out, OUT)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
