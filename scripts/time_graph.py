"""Eager call sequence vs CUDA-graph replay of one search, small batches (HBM-bound regime)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_shard  # noqa: E402
from text_similarity_b200.sharded import ShardedCorpus  # noqa: E402

dev = torch.device("cuda")


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def fp8_rows(n, d):
    c = torch.empty(n, d, dtype=torch.float8_e4m3fn, device=dev)
    for s in range(0, n, 1 << 20):
        m = min(1 << 20, n - s)
        x = torch.randn(m, d, device=dev)
        c[s:s + m] = (x / x.norm(dim=-1, keepdim=True) * 64).to(torch.float8_e4m3fn)
    return c


for name, corpus_rows, mk in (("fp8 12.5Mx384", fp8_rows(12_500_000, 384), fp8_rows),
                              ("bf16 10Mx768", make_shard(10_000_000, 768, 1, dev), lambda n, d: make_shard(n, d, 2, dev))):
    corpus = ShardedCorpus(corpus_rows)
    nbytes = corpus_rows.numel() * corpus_rows.element_size() + corpus_rows.shape[0] * 4
    for Q in (1, 8, 32):
        q = mk(Q, corpus_rows.shape[1])
        te = timed(lambda: corpus.search(q, 10))
        tg = timed(lambda: corpus.search_graphed(q, 10))
        print(f"{name} Q={Q}: eager {te:.3f} ms ({nbytes / te / 1e6:.0f} GB/s)   graph replay {tg:.3f} ms "
              f"({nbytes / tg / 1e6:.0f} GB/s)", flush=True)
    del corpus, corpus_rows
