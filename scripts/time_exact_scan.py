"""Float64 exact scan: FP64 tensor-core kernel (DMMA) against the one-query-per-warp DFMA kernel
(TSIM_NO_MMA_SCAN=1); same call, same results."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_similarity_b200 import _lib, build, ops  # noqa: E402

build.build(experiment=True)      # the knobs below exist only in the -DTSIM_EXPERIMENT flavour (libtsim_exp.so)
_lib.use_experiment_build()

dev = torch.device("cuda")


def timed(fn, reps):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


for (N, Q, D, k, dt) in [(100_000, 100, 384, 10, torch.float32), (1_000_000, 64, 768, 10, torch.float32),
                         (1_000_000, 1024, 768, 10, torch.float32), (1_000_000, 1024, 768, 100, torch.float32),
                         (1_000_000, 1024, 768, 128, torch.bfloat16), (250_000, 4096, 384, 50, torch.float32)]:
    c = torch.randn(N, D, device=dev).to(dt)
    q = torch.randn(Q, D, device=dev).to(dt)
    os.environ.pop("TSIM_NO_MMA_SCAN", None)
    ms1, (s1, i1) = timed(lambda: ops.search_topk(q, c, k, mode="exact"), 3)
    os.environ["TSIM_NO_MMA_SCAN"] = "1"
    ms0, (s0, i0) = timed(lambda: ops.search_topk(q, c, k, mode="exact"), 2)
    os.environ.pop("TSIM_NO_MMA_SCAN", None)
    tf = 2.0 * Q * N * D / 1e9
    print(f"exact scan {str(dt).split('.')[-1]} {N}x{D} Q={Q} k={k}: DMMA {ms1:.2f} ms ({tf / ms1:.2f} TFLOP/s f64)  "
          f"one-query-per-warp {ms0:.2f} ms ({tf / ms0:.2f})  x{ms0 / ms1:.2f}  "
          f"same: {bool(torch.equal(i0, i1) and torch.equal(s0, s1))}", flush=True)
