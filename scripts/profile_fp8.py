"""Small driver for ncu: the fp8 (config 4) shard search at one batch size.
    python scripts/profile_fp8.py <Q> [rows]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_similarity_b200 import ops  # noqa: E402

dev = torch.device("cuda")
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N, D = int(sys.argv[2]) if len(sys.argv) > 2 else 12_500_000, 384
c = torch.empty(N, D, dtype=torch.float8_e4m3fn, device=dev)
for s in range(0, N, 1 << 20):
    n = min(1 << 20, N - s)
    x = torch.randn(n, D, device=dev)
    c[s:s + n] = (x / x.norm(dim=-1, keepdim=True) * 64).to(torch.float8_e4m3fn)
inv = ops.row_inv_norm(c)
x = torch.randn(Q, D, device=dev)
q = (x / x.norm(dim=-1, keepdim=True) * 64).to(torch.float8_e4m3fn)
for i in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s_, i_ = ops.search_topk(q, c, 10, corpus_inv_norm=inv)
    e1.record()
    torch.cuda.synchronize()
    print(f"rep {i}: {e0.elapsed_time(e1):.3f} ms")
