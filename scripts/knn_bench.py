"""Times fc_knn_self on the embedder's four layer shapes (B clouds of N points, C = 6 / 64 / 64 / 128, k = 40) through the C ABI.
FC_KNN=fused selects the one-kernel form."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flowcompare_b200 import lib as fclib
B, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 1250)
lib = fclib.load()
st = torch.cuda.current_stream().cuda_stream
tot_ms, tot_fl = 0.0, 0.0
for C in (6, 64, 64, 128):
    x = torch.randn(B, N, C, device="cuda")
    idx = torch.empty(B, N, 40, dtype=torch.int32, device="cuda")
    run = lambda: lib.fc_knn_self(x.data_ptr(), C, B, N, C, 40, idx.data_ptr(), 0, st)
    for _ in range(2): assert run() == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2.0 * B * N * N * C
    tot_ms += ms; tot_fl += fl
    print(f"C={C:3d}: {ms:8.3f} ms  {fl / ms / 1e9:6.1f} TFLOP/s", flush=True)
print(f"{os.environ.get('FC_KNN', 'two-kernel')} B={B} N={N}: total {tot_ms:.3f} ms, {tot_fl / tot_ms / 1e9:.1f} TFLOP/s")
