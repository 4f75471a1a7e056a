"""K3 merge_topk: lists ordered best-first (rank merge) against unordered lists (bitonic sort), CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_similarity_b200 import ops  # noqa: E402

dev = torch.device("cuda")
for (Q, G, k) in [(4096, 8, 10), (4096, 8, 100), (4096, 2, 100), (32, 8, 10), (16384, 8, 5)]:
    s64 = torch.randn(Q, G, k, dtype=torch.float64, device=dev)
    ix = torch.randint(0, 10_000_000, (Q, G, k), dtype=torch.int64, device=dev)
    srt = torch.sort(s64, dim=-1, descending=True)[0]
    for name, s in (("sorted lists", srt), ("unsorted", s64)):
        a, b = s.reshape(Q, -1).contiguous(), ix.reshape(Q, -1).contiguous()
        for _ in range(3):
            ops.merge_topk(a, b, k, G)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.merge_topk(a, b, k, G)
        e1.record()
        torch.cuda.synchronize()
        print(f"merge_topk Q={Q} lists={G} k={k} {name:13s}: {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us")
