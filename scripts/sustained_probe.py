"""Sustained-throughput probe: repeat the Q=4096 search for a few seconds and report TFLOP/s and clocks.
Used to separate the power-cap effect of DRAM re-reads (small, L2-resident corpus vs a large one)."""
import os
import subprocess
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_shard  # noqa: E402
from text_similarity_b200 import ops  # noqa: E402

rows = int(sys.argv[1])
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
Q, D, k = 4096, 768, 10
dev = torch.device("cuda")
corpus = make_shard(rows, D, 1, dev)
inv = ops.row_inv_norm(corpus)
q = make_shard(Q, D, 2, dev)
for _ in range(3):
    ops.search_topk(q, corpus, k, corpus_inv_norm=inv)
torch.cuda.synchronize()
n = 0
t0 = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < secs:
    for _ in range(10):
        ops.search_topk(q, corpus, k, corpus_inv_norm=inv)
    n += 10
    torch.cuda.synchronize()
    if n % 50 == 0 or rows > 2_000_000:
        clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"],
                             capture_output=True, text=True).stdout.strip()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"rows={rows}: {n} searches in {ms / 1e3:.2f} s -> {2.0 * Q * rows * D * n / (ms * 1e-3) / 1e12:.0f} TFLOP/s "
      f"(whole call), last clocks/power: {clk}")
