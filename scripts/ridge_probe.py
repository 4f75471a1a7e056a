"""Ridge-batch diagnosis (128 < Q <= 512 on 10M x 768 bf16): is a search bound by HBM, the tensor pipe, shared memory
or the power cap?  For every knob set (experiment library) report
  * burst: one search after 0.3 s of idle, CUDA events (boost clocks, no power history), median of 7;
  * sustained: back-to-back searches for ~2 s with nvidia-smi sampled from a thread (SM clock, power, throttle reasons).
    Q=256 python scripts/ridge_probe.py "X=1" "TSIM_NO_SPLIT=1" "TSIM_NO_SPLIT=1 TSIM_DEBUG=2"
"""
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_shard  # noqa: E402
from text_similarity_b200 import _lib, build, ops  # noqa: E402

build.build(experiment=True)
_lib.use_experiment_build()

rows = int(os.environ.get("ROWS", "10000000"))
Q, D, k = int(os.environ.get("Q", "256")), int(os.environ.get("D", "768")), int(os.environ.get("K", "10"))
dtype = torch.float8_e4m3fn if os.environ.get("DT", "bf16") == "fp8" else torch.bfloat16
cfgs = [dict(kv.split("=") for kv in c.split()) for c in sys.argv[1:]]
keys = sorted({k_ for c in cfgs for k_ in c})
dev = torch.device("cuda")
corpus = make_shard(rows, D, 1, dev, dtype)
inv = ops.row_inv_norm(corpus)
q = make_shard(Q, D, 2, dev, dtype)


def smi():
    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active",
                          "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
    f = [x.strip() for x in out.split(",")]
    return float(f[0]), float(f[1]), f[2]


for c in cfgs:
    for k_ in keys:
        os.environ.pop(k_, None)
    os.environ.update(c)
    ops._destroy_plans()      # plan-time knobs: make a fresh plan under this environment
    for _ in range(3):
        ops.search_topk(q, corpus, k, corpus_inv_norm=inv)
    torch.cuda.synchronize()
    burst = []
    for _ in range(7):
        time.sleep(0.3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.search_topk(q, corpus, k, corpus_inv_norm=inv)
        e1.record()
        torch.cuda.synchronize()
        burst.append(e0.elapsed_time(e1))
    samples, stop = [], threading.Event()

    def sampler():
        while not stop.is_set():
            samples.append(smi())

    th = threading.Thread(target=sampler)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.time()
    th.start()
    e0.record()
    while time.time() - t0 < 2.0:
        for _ in range(20):
            ops.search_topk(q, corpus, k, corpus_inv_norm=inv)
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    stop.set()
    th.join()
    sus = e0.elapsed_time(e1) / n
    tail = samples[len(samples) // 2:] or samples
    clk = statistics.median(s[0] for s in tail)
    pw = statistics.median(s[1] for s in tail)
    flops = 2.0 * Q * rows * D
    gb = rows * (D * corpus.element_size() + 4) / 1e9
    print(f"{str(c):60s} burst {statistics.median(burst):7.3f} ms ({flops / statistics.median(burst) / 1e9:6.0f} TF, {gb / statistics.median(burst) * 1e3:5.0f} GB/s)"
          f"  sustained {sus:7.3f} ms ({flops / sus / 1e9:6.0f} TF, {gb / sus * 1e3:5.0f} GB/s)  sm {clk:.0f} MHz  {pw:.0f} W  {tail[-1][2]}"
          f"  tensor-busy-if-full-rate {flops / sus / 1e9 / (148 * 8192 * clk * 1e6 / 1e12) * 100:.0f} %", flush=True)
