"""A/B the float64 scan's experiment knobs: python scripts/ab_exact.py "X=1" "TSIM_MMA_VARIANT=12" "TSIM_SCAN_SLICES=74" ...
(every argument is a comma-separated set of VAR=value pairs applied for one timing)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_similarity_b200 import _lib, build, ops  # noqa: E402

build.build(experiment=True)      # the knobs below exist only in the -DTSIM_EXPERIMENT flavour (libtsim_exp.so)
_lib.use_experiment_build()

dev = torch.device("cuda")
shapes = [(1_000_000, 1024, 768, 10), (1_000_000, 1024, 768, 100), (250_000, 4096, 384, 50), (1_000_000, 64, 768, 10)]
data = {sh: (torch.randn(sh[0], sh[2], device=dev), torch.randn(sh[1], sh[2], device=dev)) for sh in shapes}
for arg in sys.argv[1:]:
    env = dict(kv.split("=") for kv in arg.split(","))
    os.environ.update(env)
    out = []
    for sh in shapes:
        c, q = data[sh]
        N, Q, D, k = sh
        for _ in range(2):
            ops.search_topk(q, c, k, mode="exact")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.search_topk(q, c, k, mode="exact")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out.append(f"Q={Q} k={k}: {ms:7.2f} ms {2.0 * Q * N * D / ms / 1e9:5.2f} TF")
    for key in env:
        os.environ.pop(key, None)
    print(f"{arg:40s} " + " | ".join(out), flush=True)
