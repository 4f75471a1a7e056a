"""One float64 exact scan for ncu (1M x 768 fp32, Q = 1024, k = 10): ncu -k regex:search_exact -c 1 ..."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_similarity_b200 import ops  # noqa: E402

dev = torch.device("cuda")
N, Q, D, k = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (1_000_000, 1024, 768, 10)))
c = torch.randn(N, D, device=dev)
q = torch.randn(Q, D, device=dev)
s, i = ops.search_topk(q, c, k, mode="exact")
torch.cuda.synchronize()
print(i[0].tolist())
