"""Small driver for ncu: runs the search hot path a few times on synthetic data.

    python scripts/profile_search.py --rows 2000000 --queries 4096 --k 10 --reps 3
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_shard  # noqa: E402
from text_similarity_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=2_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--queries", type=int, default=4096)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--mode", default="auto")
a = ap.parse_args()
dev = torch.device("cuda", 0)
corpus = make_shard(a.rows, a.dim, 1234, dev)
inv = ops.row_inv_norm(corpus)
q = make_shard(a.queries, a.dim, 4321, dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.reps):
    ev0.record()
    s, ix = ops.search_topk(q, corpus, a.k, corpus_inv_norm=inv, mode=a.mode)
    ev1.record()
    torch.cuda.synchronize()
    print(f"rep {i}: {ev0.elapsed_time(ev1):.3f} ms  ({a.queries / ev0.elapsed_time(ev1) * 1e3:.0f} q/s)")
print("checksum", int(ix.sum()), float(s.sum()))
