"""K1 diagnosis: CUDA-event time of one large pooling call per TSIM_POOL_DEBUG / TSIM_POOL_MODE setting.
    python scripts/pool_dbg.py 0 1        # debug masks to try (1 = skip the accumulation)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_similarity_b200 import ops  # noqa: E402

dev = torch.device("cuda")
B, L, D = 16384, 64, 768
tok = torch.randn(B, L, D, device=dev).to(torch.bfloat16)
tok2 = tok.clone()
mask = torch.ones(B, L, dtype=torch.int64, device=dev)
out = torch.empty(B, D, dtype=torch.bfloat16, device=dev)
inv = torch.empty(B, dtype=torch.float32, device=dev)
for dbg in sys.argv[1:] or ["0"]:
    os.environ["TSIM_POOL_DEBUG"] = dbg
    ts = []
    for i in range(8):
        t = tok if i % 2 == 0 else tok2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        ops.pool_norm(t, mask, out=out, out_inv_norm=inv, normalize=True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts[2:])[len(ts[2:]) // 2]
    print(f"dbg={dbg} mode={os.environ.get('TSIM_POOL_MODE', 'auto')}: {ms * 1e3:.1f} us  {tok.numel() * 2 / ms / 1e6:.0f} GB/s", flush=True)
