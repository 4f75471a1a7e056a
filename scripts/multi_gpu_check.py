"""torchrun check of the sharded path over NCCL: every rank holds a contiguous shard, the merged
result must equal a single-GPU search of the whole corpus (computed locally on every rank).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/multi_gpu_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_shard  # noqa: E402
from text_similarity_b200 import ops  # noqa: E402
from text_similarity_b200.sharded import ShardedCorpus, shard_bounds  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for (N, Q, D, k) in [(300_001, 257, 768, 10), (200_000, 64, 384, 100), (50_000, 5, 768, 10)]:
    full = make_shard(N, D, seed=99, dev=dev)            # same seed on every rank -> same corpus
    full[N // 2] = full[7]
    full[N - 1] = full[7]
    q = make_shard(Q, D, seed=7, dev=dev)
    b, e = shard_bounds(N, world, rank)
    sc = ShardedCorpus(full[b:e].contiguous(), idx_base=b, group=dist.group.WORLD)
    s, i = sc.search(q, k)
    fs, fi = ops.search_topk(q, full, k)
    same = torch.equal(i, fi) and torch.equal(s, fs)
    # exclude-self across shards: queries are rows 1000.. of the corpus
    qs = full[1000:1000 + Q].contiguous()
    s2, i2 = sc.search(qs, 5, exclude_self_base=1000)
    fs2, fi2 = ops.search_topk(qs, full, 5, exclude_self_base=1000)
    same = same and torch.equal(i2, fi2) and torch.equal(s2, fs2)
    # local search replayed from a CUDA graph, then the same all-gather + merge
    for _ in range(2):
        s3, i3 = sc.search_graphed(q, k)
        same = same and torch.equal(i3, fi) and torch.equal(s3, fs)
    t = torch.tensor([int(same)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ok = ok and bool(t.item())
    if rank == 0:
        print(f"N={N} Q={Q} D={D} k={k} world={world}: {'OK' if t.item() else 'MISMATCH'}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
