"""fp32 rows with 24 < k <= 100: the split (hi + lo) bf16 shadow on the tensor cores against the float64 exact scan."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_similarity_b200 import ops  # noqa: E402

dev = torch.device("cuda")


def timed(fn, reps=3):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


for (N, Q, D, k) in [(1_000_000, 1024, 768, 100), (1_000_000, 1024, 768, 50), (1_000_000, 1024, 768, 10), (250_000, 4096, 384, 50)]:
    c = torch.randn(N, D, device=dev)
    c = c / c.norm(dim=-1, keepdim=True)
    q = torch.randn(Q, D, device=dev)
    sh, sinv = ops.make_shadow(c, split=True)
    ms_s, a = timed(lambda: ops.search_topk(q, c, k, corpus_shadow=sh, shadow_inv_norm=sinv, return_score64=True, return_flags=True))
    ms_e, b = timed(lambda: ops.search_topk(q, c, k, mode="exact", return_score64=True), reps=2)
    line = f"fp32 {N}x{D} Q={Q} k={k}: split shadow {ms_s:7.2f} ms  exact scan {ms_e:7.2f} ms  x{ms_e / ms_s:5.1f}  same idx {bool(torch.equal(a[1], b[1]))} " \
           f"same f64 {bool(torch.equal(a[2], b[2]))} flagged {int((a[3] != 0).sum())}"
    if k <= 24:
        rs, rinv = ops.make_shadow(c)
        ms_r, _ = timed(lambda: ops.search_topk(q, c, k, corpus_shadow=rs, shadow_inv_norm=rinv))
        line += f"  rounded shadow {ms_r:7.2f} ms"
    print(line, flush=True)
    del c, sh
