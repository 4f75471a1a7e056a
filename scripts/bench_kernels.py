"""CUDA-event timings of the small kernels around the search: K1 pool_norm, row_inv_norm, and the
K3 select / merge passes.  Prints achieved GB/s on each kernel's ALGORITHMIC bytes."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_similarity_b200 import ops  # noqa: E402

dev = torch.device("cuda")
peak = 6547.8
if os.path.exists("MEASURED_PEAKS.json"):
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]


def timeit(fn, reps=20, flush=None):
    for _ in range(3):
        fn()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    for i in range(reps):
        if flush is not None:
            flush.zero_()          # 256 MB write: evicts L2 between repetitions
        e0[i].record()
        fn()
        e1[i].record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in zip(e0, e1))
    return t[len(t) // 2]


def timeit_rotating(fns, rounds=8):
    """Back-to-back launches over DISTINCT input buffers (each larger than what L2 keeps of the
    previous ones; no flush kernel in between, so no dirty lines of a flush are written back
    during the measured kernel): one event pair around the whole train, time per launch."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        for f in fns:
            f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (rounds * len(fns))


flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print(f"HBM peak used for fractions: {peak} GB/s")
for (B, L, D, in_dt, out_dt) in [(16, 256, 384, torch.float32, torch.float32), (16, 256, 768, torch.float32, torch.bfloat16),
                                 (256, 128, 768, torch.float32, torch.bfloat16), (1024, 128, 768, torch.bfloat16, torch.bfloat16),
                                 (4096, 64, 768, torch.bfloat16, torch.float8_e4m3fn)]:
    tok = torch.randn(B, L, D, device=dev).to(in_dt)
    mask = torch.ones(B, L, dtype=torch.int64, device=dev)
    out = torch.empty(B, D, dtype=out_dt, device=dev)
    inv = torch.empty(B, dtype=torch.float32, device=dev)
    ms = timeit(lambda: ops.pool_norm(tok, mask, out=out, out_inv_norm=inv, normalize=True), flush=flush)
    nbytes = tok.numel() * tok.element_size() + mask.numel() * 8 + out.numel() * out.element_size() + B * 4
    print(f"K1 pool_norm B={B} L={L} D={D} {str(in_dt)[6:]}->{str(out_dt)[6:]}: {ms * 1e3:8.1f} us  "
          f"{nbytes / ms / 1e6:7.0f} GB/s  ({nbytes / ms / 1e6 / peak * 100:5.1f}% of HBM peak, {nbytes / 1e6:.1f} MB)")
    if nbytes > 64e6:
        # sustained: a train of launches rotating over input copies totalling >= 1 GB (>> the 126 MB L2)
        ncopy = max(3, int(1.0e9 // nbytes) + 1)
        toks = [tok] + [tok.clone() for _ in range(ncopy - 1)]
        fns = [(lambda t=t: ops.pool_norm(t, mask, out=out, out_inv_norm=inv, normalize=True)) for t in toks]
        ms_r = timeit_rotating(fns)
        print(f"   rotating {ncopy} inputs, launch train: {ms_r * 1e3:8.1f} us  {nbytes / ms_r / 1e6:7.0f} GB/s  "
              f"({nbytes / ms_r / 1e6 / peak * 100:5.1f}% of HBM peak)")
        del toks, fns
    # the six-kernel ATen sequence the reference issues (modules.py:160-170) on the same GPU, for scale
    def aten():
        m = mask.unsqueeze(-1).expand(tok.size()).float()
        return torch.sum(tok * m, 1) / torch.clamp(m.sum(1), min=1e-9)
    ms_ref = timeit(aten, flush=flush)
    print(f"   reference ATen sequence on this GPU: {ms_ref * 1e3:8.1f} us  (x{ms_ref / ms:.1f})")

for (N, D, dt) in [(10_000_000, 768, torch.bfloat16), (25_000_000, 384, torch.float8_e4m3fn)]:
    x = torch.empty(N, D, dtype=dt, device=dev)
    x.view(torch.uint8).random_(0, 100)
    ms = timeit(lambda: ops.row_inv_norm(x), reps=5)
    nbytes = x.numel() * x.element_size() + N * 4
    print(f"row_inv_norm {N}x{D} {str(dt)[6:]}: {ms:7.3f} ms  {nbytes / ms / 1e6:7.0f} GB/s ({nbytes / ms / 1e6 / peak * 100:5.1f}%)")
    del x

for (Q, G, k) in [(4096, 8, 10), (4096, 8, 100), (32, 8, 10)]:
    s64 = torch.randn(Q, G * k, dtype=torch.float64, device=dev)
    ix = torch.randint(0, 10_000_000, (Q, G * k), dtype=torch.int64, device=dev)
    ms = timeit(lambda: ops.merge_topk(s64, ix, k, G))
    nbytes = Q * G * k * 16 + Q * k * 20
    print(f"K3 merge_topk Q={Q} lists={G} k={k}: {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s")
