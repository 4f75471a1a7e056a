"""Time the per-GPU work of every BASELINE.json config shape on ONE B200 (CUDA events, whole call through
the public op: candidate passes + select + fallback), next to the roofline each is bound by.

    python scripts/time_configs.py [names...]     names: c2 c3 c4 c5 ridge
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_shard  # noqa: E402
from text_similarity_b200 import ops  # noqa: E402

dev = torch.device("cuda")
peaks = {"hbm_gbs": 6547.8, "bf16_tflops": 1583.0}
if os.path.exists("MEASURED_PEAKS.json"):
    peaks.update(json.load(open("MEASURED_PEAKS.json")))
want = set(sys.argv[1:]) or {"c2", "c3", "c4", "c5", "ridge"}


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, ms, Q, N, D, esz, k):
    flops = 2.0 * Q * N * D
    nbytes = N * D * esz + N * 4 + Q * D * esz + Q * k * 12
    tf, gbs = flops / ms / 1e9, nbytes / ms / 1e6
    peak_tf = peaks["bf16_tflops"] * (2 if esz == 1 else 1)
    print(f"{name}: {ms:9.3f} ms  {Q / ms * 1e3:10.0f} q/s  {tf:7.1f} TFLOP/s ({tf / peak_tf * 100:4.1f}% of "
          f"{'fp8 = 2x ' if esz == 1 else ''}bf16 peak)  {gbs:7.0f} GB/s ({gbs / peaks['hbm_gbs'] * 100:5.1f}% of HBM peak)",
          flush=True)


if "c2" in want:      # config 2: 1M x 768 bf16, 1024 queries, top-10, one GPU
    c = make_shard(1_000_000, 768, 1, dev); inv = ops.row_inv_norm(c); q = make_shard(1024, 768, 2, dev)
    report("config2 1Mx768 bf16 Q=1024 k=10", timed(lambda: ops.search_topk(q, c, 10, corpus_inv_norm=inv), 20),
           1024, 1_000_000, 768, 2, 10)
    del c, inv
if "c3" in want:      # config 3: 10M x 768 bf16 over 8 GPUs -> 1.25M-row shard, 4096 queries, top-100 (+ whole corpus on one GPU)
    for rows in (1_250_000, 10_000_000):
        c = make_shard(rows, 768, 3, dev); inv = ops.row_inv_norm(c); q = make_shard(4096, 768, 4, dev)
        report(f"config3 shard {rows}x768 bf16 Q=4096 k=100", timed(lambda: ops.search_topk(q, c, 100, corpus_inv_norm=inv), 6),
               4096, rows, 768, 2, 100)
        del c, inv
if "c4" in want:      # config 4: 100M x 384 e4m3 over 8 GPUs -> 12.5M-row shard, batch 1..32, top-10
    N, D = 12_500_000, 384
    c = torch.empty(N, D, dtype=torch.float8_e4m3fn, device=dev)
    for s in range(0, N, 1 << 20):
        n = min(1 << 20, N - s)
        x = torch.randn(n, D, device=dev)
        c[s:s + n] = (x / x.norm(dim=-1, keepdim=True) * 64).to(torch.float8_e4m3fn)
    inv = ops.row_inv_norm(c)
    for Q in (1, 2, 4, 8, 16, 32):
        x = torch.randn(Q, D, device=dev)
        q = (x / x.norm(dim=-1, keepdim=True) * 64).to(torch.float8_e4m3fn)
        report(f"config4 shard {N}x{D} e4m3 Q={Q} k=10", timed(lambda: ops.search_topk(q, c, 10, corpus_inv_norm=inv), 20),
               Q, N, D, 1, 10)
    del c, inv
if "c5" in want:      # config 5: all-pairs top-5 over 1M x 768 bf16, self excluded (whole job on one GPU)
    N, D, k = 1_000_000, 768, 5
    x = make_shard(N, D, 5, dev); inv = ops.row_inv_norm(x)
    idx = torch.empty(N, k, dtype=torch.int64, device=dev)
    sc = torch.empty(N, k, dtype=torch.float32, device=dev)
    for tile in (16_384, 65_536):
        def job():
            for b in range(0, N, tile):
                e = min(N, b + tile)
                ops.search_topk(x[b:e], x, k, corpus_inv_norm=inv, exclude_self_base=b, out_scores=sc[b:e], out_idx=idx[b:e])
        report(f"config5 all-pairs {N}x{D} bf16 k=5 (query tiles of {tile})", timed(job, 1), N, N, D, 2, k)
    del x, inv
if "ridge" in want:   # between the regimes: 10M x 768 bf16, Q = 64 .. 512
    c = make_shard(10_000_000, 768, 1, dev); inv = ops.row_inv_norm(c)
    for Q in (64, 128, 192, 256, 512):
        q = make_shard(Q, 768, 2, dev)
        report(f"ridge 10Mx768 bf16 Q={Q} k=10", timed(lambda: ops.search_topk(q, c, 10, corpus_inv_norm=inv), 10),
               Q, 10_000_000, 768, 2, 10)
