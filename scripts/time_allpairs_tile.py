"""All-pairs top-5 over 1M x 768 bf16 (BASELINE config 5) on one GPU as a function of the query-tile size:
the sample passes and select_rescore are paid once per search call, so bigger tiles amortise them."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_shard  # noqa: E402
from text_similarity_b200 import ops  # noqa: E402
from text_similarity_b200.sharded import all_pairs_query_sharded  # noqa: E402

N, D, k = 1_000_000, 768, 5
dev = torch.device("cuda")
full = make_shard(N, D, 1, dev)
inv = ops.row_inv_norm(full)
ref = None
for tile in [int(x) for x in (sys.argv[1:] or ["16384", "32768", "65536", "131072"])]:
    all_pairs_query_sharded(full[:2 * tile], inv[:2 * tile], k, 1, 0, tile=tile)      # warm (plans, workspace)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s, i, _ = all_pairs_query_sharded(full, inv, k, 1, 0, tile=tile)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    same = "" if ref is None else f"  same rows as tile {ref[0]}: {torch.equal(i, ref[1])}"
    if ref is None:
        ref = (tile, i.clone())
    print(f"tile {tile:7d}: {ms:8.1f} ms  {2.0 * N * N * D / (ms * 1e-3) / 1e12:6.0f} TFLOP/s{same}", flush=True)
