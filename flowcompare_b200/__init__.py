"""flowcompare_b200 -- B200-native per-point conditional log-likelihood path of FlowCompare.

Host side (`engine`) mirrors the reference's `model_dict` / `inner_loop` interface and calls hand-written
sm_100a CUDA through the C ABI in `libflowcompare_b200.so` (`lib`).  No CPU fallback.
"""
from .engine import FlowCompareB200, accelerate, get_knn, inner_loop, knn, log_prob_to_change  # noqa: F401
