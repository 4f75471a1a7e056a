"""select_rescore in isolation is hard to launch; time the whole k = 100 shard search per knob setting instead and
print the per-kernel durations seen by CUDA events around the call (whole call only).
    python scripts/time_select.py "X=1" "TSIM_PAIR_RANK_MAX=256" "TSIM_PAIR_RANK_MAX=0"
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_shard  # noqa: E402
from text_similarity_b200 import _lib, build, ops  # noqa: E402

build.build(experiment=True)
_lib.use_experiment_build()
dev = torch.device("cuda")
rows, Q, D, k = 1_250_000, 4096, 768, int(os.environ.get("K", "100"))
corpus = make_shard(rows, D, 1, dev)
inv = ops.row_inv_norm(corpus)
q = make_shard(Q, D, 2, dev)
cfgs = [dict(kv.split("=") for kv in c.split()) for c in sys.argv[1:]]
keys = sorted({k_ for c in cfgs for k_ in c})
for rnd in range(3):
    for c in cfgs:
        for k_ in keys:
            os.environ.pop(k_, None)
        os.environ.update(c)
        ops.search_topk(q, corpus, k, corpus_inv_norm=inv)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.search_topk(q, corpus, k, corpus_inv_norm=inv)
        e1.record()
        torch.cuda.synchronize()
        print(f"round {rnd} {str(c):40s} {e0.elapsed_time(e1) / 10:8.3f} ms/search")
