"""TEST INFRASTRUCTURE ONLY -- tests/golden/voxelize.pt: outputs of the UNMODIFIED reference's utils.voxelize (utils.py:446-454,
which calls knn.get_knn, knn.py:79-90) on seeded clouds, called the way dataloaders/ams_voxel_loader.py:200-204 calls it
(start / end = per-axis min / max of the cloud, size = the loader's final_voxel_size).
python -m oracle.make_voxelize_golden; needs /root/reference."""
import os

import torch

from oracle import refload

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "voxelize.pt")

CASES = {   # name: (seed, n points, box extent, voxel size)
    "loader_default": (0, 6000, [31.0, 27.5, 9.0], [3.0, 3.0, 4.0]),
    "fine": (1, 20000, [10.0, 12.0, 3.0], [0.7, 0.9, 1.1]),
    "single_layer": (2, 3000, [8.0, 5.0, 0.4], [1.0, 1.0, 4.0]),
    "planar": (3, 2500, [6.0, 9.0], [0.5, 1.5]),
}


def inputs(name):
    seed, n, extent, size = CASES[name]
    g = torch.Generator().manual_seed(7700 + seed)
    D = len(size)
    pos = torch.rand(n, D, generator=g) * torch.tensor(extent) + torch.tensor([120.0, -45.0, 2.0][:D])
    return pos, pos.min(dim=0)[0], pos.max(dim=0)[0], torch.tensor(size)


def main():
    refload.load()
    import utils
    out = {}
    for name in CASES:
        pos, start, end, size = inputs(name)
        labels, centers = utils.voxelize(pos, start=start, end=end, size=size)
        out[name] = {"labels": labels.clone(), "centers": centers.clone()}
        print(name, tuple(labels.shape), tuple(centers.shape), int(labels.max()))
    out["meta"] = {"generator": "oracle/make_voxelize_golden.py", "source": "unmodified reference utils.voxelize, CPU fp32"}
    torch.save(out, OUT)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
