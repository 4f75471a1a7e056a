"""Small driver for ncu --set full: two K1 launches on a 1.6 GB bf16 token tensor (B=16384, L=64, D=768)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_similarity_b200 import ops  # noqa: E402

dev = torch.device("cuda")
B, L, D = 16384, 64, 768
tok = torch.randn(B, L, D, device=dev).to(torch.bfloat16)
mask = torch.ones(B, L, dtype=torch.int64, device=dev)
out = torch.empty(B, D, dtype=torch.bfloat16, device=dev)
inv = torch.empty(B, dtype=torch.float32, device=dev)
for _ in range(2):
    ops.pool_norm(tok, mask, out=out, out_inv_norm=inv, normalize=True)
torch.cuda.synchronize()
print("ok")
