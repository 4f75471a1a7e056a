"""Interleaved A/B timing of planner knobs (environment variables read per call by libtsim) on the
bench workload: same process, same box, same thermal state.
    python scripts/ab_probe.py "X=1" "TSIM_CHUNK_ROWS=32768" "TSIM_FORCE_BOOT=1 TSIM_CHUNK_ROWS=16384"
    ROWS=12500000 D=384 DT=fp8 Q=32 python scripts/ab_probe.py "X=1" "TSIM_NO_FUSED=1"      # config 4's shard
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_shard  # noqa: E402
from text_similarity_b200 import _lib, build, ops  # noqa: E402

build.build(experiment=True)      # the knobs below exist only in the -DTSIM_EXPERIMENT flavour (libtsim_exp.so)
_lib.use_experiment_build()

rows = int(os.environ.get("ROWS", "10000000"))
Q, D, k = int(os.environ.get("Q", "4096")), int(os.environ.get("D", "768")), int(os.environ.get("K", "10"))
dtype = torch.float8_e4m3fn if os.environ.get("DT", "bf16") == "fp8" else torch.bfloat16
cfgs = [dict(kv.split("=") for kv in c.split()) for c in sys.argv[1:]]
dev = torch.device("cuda")
corpus = make_shard(rows, D, 1, dev, dtype)
inv = ops.row_inv_norm(corpus)
q = make_shard(Q, D, 2, dev, dtype)
keys = sorted({k_ for c in cfgs for k_ in c})
res = {i: [] for i in range(len(cfgs))}
for rnd in range(4):
    for i, c in enumerate(cfgs):
        for k_ in keys:
            os.environ.pop(k_, None)
        os.environ.update(c)
        ops._destroy_plans()      # plan-time knobs: make a fresh plan under this environment
        ops.search_topk(q, corpus, k, corpus_inv_norm=inv)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 12
        e0.record()
        for _ in range(n):
            ops.search_topk(q, corpus, k, corpus_inv_norm=inv)
        e1.record()
        torch.cuda.synchronize()
        if rnd > 0:
            res[i].append(e0.elapsed_time(e1) / n)
for i, c in enumerate(cfgs):
    ms = sorted(res[i])[len(res[i]) // 2]
    print(f"{str(c):70s} median {ms:8.4f} ms/search  {2.0 * Q * rows * D / (ms * 1e-3) / 1e12:7.0f} TFLOP/s  "
          f"{rows * (D * corpus.element_size() + 4) / (ms * 1e-3) / 1e9:6.0f} GB/s  all={['%.3f' % x for x in res[i]]}")
