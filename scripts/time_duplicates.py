"""Cost of queries the first tensor pass cannot prove (heavy duplication in the corpus): the retry stage
(KP = 112 tensor pass over compact blocks of flagged queries) against the float64 scan it replaces.

    python scripts/time_duplicates.py            # 2M x 768 bf16, Q = 4096, 0 / 8 / 64 / 400 flagged queries
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_shard  # noqa: E402
from text_similarity_b200 import _lib, build, ops  # noqa: E402

build.build(experiment=True)      # the knobs below exist only in the -DTSIM_EXPERIMENT flavour (libtsim_exp.so)
_lib.use_experiment_build()

N, D, Q, k = int(os.environ.get("ROWS", "2000000")), 768, 4096, 10
dev = torch.device("cuda")
corpus = make_shard(N, D, 1, dev)
g = torch.Generator().manual_seed(5)
for src in range(400):                         # rows 0..399 occur 31 times each
    pos = torch.randperm(N - 1000, generator=g)[:30] + 1000
    corpus[pos.to(dev)] = corpus[src].clone()
inv = ops.row_inv_norm(corpus)
for nflag in (0, 8, 64, 400):
    q = make_shard(Q, D, 2, dev)
    q[:nflag] = corpus[:nflag]
    for knob in ("", "1"):
        os.environ.pop("TSIM_NO_RETRY", None)
        if knob:
            os.environ["TSIM_NO_RETRY"] = knob
        for _ in range(2):
            s, i, fl = ops.search_topk(q, corpus, k, corpus_inv_norm=inv, return_flags=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.search_topk(q, corpus, k, corpus_inv_norm=inv)
        e1.record()
        torch.cuda.synchronize()
        print(f"flagged {int((fl != 0).sum()):4d} (retry-answered {int((fl == 2).sum()):4d}, scan-answered {int((fl == 1).sum()):4d})  "
              f"{'float64 scan only' if knob else 'retry stage      '}: {e0.elapsed_time(e1) / 3:9.3f} ms / search")
