"""Time fp32 inputs: the float64 exact scan (mode="exact") against the bf16-shadow tensor path with float64
re-score on the originals (BASELINE config 1 shape and larger cases)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_similarity_b200 import ops  # noqa: E402

dev = torch.device("cuda")
for (N, Q, D, k) in [(10_000, 100, 384, 10), (100_000, 100, 384, 10), (1_000_000, 64, 768, 10), (1_000_000, 1024, 768, 10)]:
    c = torch.randn(N, D, device=dev)
    q = torch.randn(Q, D, device=dev)
    for _ in range(2):
        ops.search_topk(q, c, k, mode="exact")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.search_topk(q, c, k, mode="exact")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"exact scan fp32 {N}x{D} Q={Q} k={k}: {ms:.3f} ms  {Q / ms * 1e3:.0f} q/s  "
          f"{2.0 * Q * N * D / ms / 1e9:.1f} GFLOP/s (f64)")
    shadow, sinv = ops.make_shadow(c)
    for _ in range(2):
        ops.search_topk(q, c, k, corpus_shadow=shadow, shadow_inv_norm=sinv)
    e0.record()
    for _ in range(5):
        s2, i2 = ops.search_topk(q, c, k, corpus_shadow=shadow, shadow_inv_norm=sinv)
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 5
    s1, i1 = ops.search_topk(q, c, k, mode="exact")
    print(f"   bf16 shadow + f64 re-score:       {ms2:.3f} ms  {Q / ms2 * 1e3:.0f} q/s  (x{ms / ms2:.1f}; same indices: {bool(torch.equal(i1, i2))})")
