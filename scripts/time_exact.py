"""Time the float64 exact-scan path (fp32 inputs; BASELINE config 1 shape and a larger case)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_similarity_b200 import ops  # noqa: E402

dev = torch.device("cuda")
for (N, Q, D, k) in [(10_000, 100, 384, 10), (100_000, 100, 384, 10), (1_000_000, 64, 768, 10)]:
    c = torch.randn(N, D, device=dev)
    q = torch.randn(Q, D, device=dev)
    for _ in range(2):
        ops.search_topk(q, c, k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.search_topk(q, c, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"exact scan fp32 {N}x{D} Q={Q} k={k}: {ms:.3f} ms  {Q / ms * 1e3:.0f} q/s  "
          f"{2.0 * Q * N * D / ms / 1e9:.1f} GFLOP/s (f64)")
