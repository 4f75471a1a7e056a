"""BASELINE configs 4 and 5 on N GPUs (torchrun, one rank per GPU, NCCL), timed like bench.py: barrier +
synchronize on both sides, CUDA events, max over ranks.  One JSON line per measurement (rank 0).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
        scripts/bench_configs_multi.py [c4] [c5]

config 4: 100M x 384 e4m3 corpus sharded by rows, query batches of 1 / 8 / 32, top-10 (HBM-bound stream).
config 5: all-pairs top-5 over 1M x 768 bf16, self excluded: corpus rows sharded, every rank scores all rows
          (query tiles of 16384) against its shard, one all-gather + merge per tile.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_shard  # noqa: E402
from text_similarity_b200.sharded import ShardedCorpus, shard_bounds  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
group = dist.group.WORLD if world > 1 else None
want = set(sys.argv[1:]) or {"c4", "c5"}


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def emit(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


if "c4" in want:
    N, D, k = 100_000_000, 384, 10
    b, e = shard_bounds(N, world, rank)
    rows = e - b
    shard = torch.empty(rows, D, dtype=torch.float8_e4m3fn, device=dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for s in range(0, rows, 1 << 20):
        n = min(1 << 20, rows - s)
        x = torch.randn(n, D, device=dev, generator=g)
        shard[s:s + n] = (x / x.norm(dim=-1, keepdim=True) * 64).to(torch.float8_e4m3fn)
    corpus = ShardedCorpus(shard, idx_base=b, group=group)
    for Q in (1, 8, 32):
        x = torch.randn(Q, D, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
        q = (x / x.norm(dim=-1, keepdim=True) * 64).to(torch.float8_e4m3fn)
        for graphed in (False, True):
            fn = (lambda: corpus.search_graphed(q, k)) if graphed else (lambda: corpus.search(q, k))
            ms = timed(fn, 30)
            emit(config="4: 100Mx384 e4m3, top-10", n_gpus=world, rows_per_gpu=rows, queries_per_batch=Q, cuda_graph=graphed,
                 ms_per_batch=ms, queries_per_s=Q / ms * 1e3, stream_GBps_per_gpu=(rows * D + rows * 4) / ms / 1e6,
                 aggregate_TBps=(N * D + N * 4) / ms / 1e9)
    del corpus, shard

if "c5" in want:
    N, D, k, tile = 1_000_000, 768, 5, 16_384
    full = make_shard(N, D, seed=55, dev=dev)           # same seed on every rank: all rows are queries everywhere
    b, e = shard_bounds(N, world, rank)
    corpus = ShardedCorpus(full[b:e], idx_base=b, group=group)
    idx = torch.empty(N, k, dtype=torch.int64, device=dev)

    def job():
        for t0 in range(0, N, tile):
            t1 = min(N, t0 + tile)
            _, rows_ = corpus.search(full[t0:t1], k, exclude_self_base=t0)
            idx[t0:t1] = rows_

    ms = timed(job, 2, warm=1)
    emit(config="5: all-pairs top-5 over 1Mx768 bf16, self excluded", n_gpus=world, rows_per_gpu=e - b, query_tile=tile,
         seconds_per_job=ms / 1e3, rows_per_s=N / ms * 1e3, aggregate_TFLOPs=2.0 * N * N * D / ms / 1e9,
         self_hits=int((idx == torch.arange(N, device=dev)[:, None]).sum().item()))
if world > 1:
    dist.destroy_process_group()
