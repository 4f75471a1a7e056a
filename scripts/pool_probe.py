"""Small driver for ncu: K1 launches over a batch-size sweep (scripts/bench_kernels.py times them with CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_similarity_b200 import ops  # noqa: E402

dev = torch.device("cuda")
shapes = [(B, 64, 768, torch.bfloat16, torch.bfloat16) for B in (1024, 2048, 4096, 8192, 16384)]
shapes += [(256, 128, 768, torch.float32, torch.bfloat16), (8192, 64, 384, torch.float32, torch.bfloat16)]
for (B, L, D, in_dt, out_dt) in shapes:
    tok = torch.randn(B, L, D, device=dev).to(in_dt)
    mask = torch.ones(B, L, dtype=torch.int64, device=dev)
    out = torch.empty(B, D, dtype=out_dt, device=dev)
    inv = torch.empty(B, dtype=torch.float32, device=dev)
    ops.pool_norm(tok, mask, out=out, out_inv_norm=inv, normalize=True)
    torch.cuda.synchronize()
    print("ok", B, L, D, tok.numel() * tok.element_size() / 1e6, "MB")
    del tok
