"""TEST INFRASTRUCTURE ONLY -- golden vectors for the change score from the UNMODIFIED reference.

    python -m oracle.make_change_golden      # needs /root/reference; writes tests/golden/change_score.pt

Runs the reference's own `log_prob_to_change` (test_flow.py:249-275, with `clamp_infs` :241-247) on seeded log-prob
tensors that contain -inf entries, for the std-multiple and the hard-cutoff variants, and stores inputs + outputs
(a few 10 KB).  The reference mutates its arguments in place (clamp_infs, and the score is written into
log_prob_1_given_0); clones are passed so the stored inputs are the originals."""
import os

import torch

from oracle import refload
from oracle.make_golden import GOLDEN_DIR


def cases():
    g = torch.Generator().manual_seed(2024)
    out = []
    for B, N, ninf in ((4, 1024, 3), (1, 257, 0), (3, 64, 5)):
        lp10 = torch.randn(B, N, generator=g) * 6 - 25
        lp00 = torch.randn(B, N, generator=g) * 1.5 - 12
        for i in range(ninf):
            lp10[i % B, (7 * i + 3) % N] = float("-inf")
            lp00[(i + 1) % B, (11 * i + 5) % N] = float("-inf")
        out.append((lp10, lp00))
    return out


def main():
    tf = refload.load_test_flow()
    recs = []
    for lp10, lp00 in cases():
        for multiple, cut in ((5.4, None), (1.5, None), (3.0, -27.5)):
            want = tf.log_prob_to_change(lp10.clone(), lp00.clone(), multiple, cut)
            recs.append({"lp10": lp10, "lp00": lp00, "multiple": multiple, "hard_cutoff": cut, "change": want.clone()})
    path = os.path.join(GOLDEN_DIR, "change_score.pt")
    torch.save({"cases": recs, "meta": {"generator": "oracle/make_change_golden.py",
                                        "source": "unmodified reference test_flow.log_prob_to_change, CPU fp32"}}, path)
    print(f"{len(recs)} cases -> {path} ({os.path.getsize(path)} B)")


if __name__ == "__main__":
    main()
