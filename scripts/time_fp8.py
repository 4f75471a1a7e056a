"""Time the fp8 (e4m3) corpus-streaming regime of BASELINE config 4 on one GPU's shard."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_similarity_b200 import ops  # noqa: E402

dev = torch.device("cuda")
N, D = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000, 384
c = torch.empty(N, D, dtype=torch.float8_e4m3fn, device=dev)
for s in range(0, N, 1 << 20):
    n = min(1 << 20, N - s)
    x = torch.randn(n, D, device=dev)
    c[s:s + n] = (x / x.norm(dim=-1, keepdim=True) * 64).to(torch.float8_e4m3fn)
inv = ops.row_inv_norm(c)
for Q in [int(x) for x in os.environ.get("QS", "1,2,4,8,16,32").split(",")]:
    q = torch.randn(Q, D, device=dev)
    q = (q / q.norm(dim=-1, keepdim=True) * 64).to(torch.float8_e4m3fn)
    for _ in range(3):
        ops.search_topk(q, c, 10, corpus_inv_norm=inv)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.search_topk(q, c, 10, corpus_inv_norm=inv)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"fp8 {N}x{D} Q={Q}: {ms:.3f} ms/search  {(N * D + N * 4) / ms / 1e6:.0f} GB/s (whole call)  {Q / ms * 1e3:.0f} q/s")
