"""Corpus-row sharding across the GPUs of one box (BASELINE north_star (d)).

Each rank holds a contiguous block of corpus rows and answers every query against it; ONE
NCCL all-gather moves the per-shard (float64 score, global row) lists and every rank merges
them with the K3 merge kernel.  The reference has no multi-device code at all (SURVEY.md
section 2.3) and its single-device chunk loop overwrites earlier chunks
(src/pipeline/search_pipeline.py:60-61,83,88); this is the repaired intent at box scale.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block [begin, end) of `rank`: ceil(N / G) rows each, last one ragged."""
    per = (n_rows + world - 1) // world
    begin = min(n_rows, rank * per)
    return begin, min(n_rows, begin + per)


def gather_shard_results(score64: torch.Tensor, idx: torch.Tensor, group, bufs=None
                         ) -> Tuple[torch.Tensor, torch.Tensor]:
    """The ONE collective of the search: all-gather every rank's [Q, k] (float64 score, int64 global
    row) lists and lay them out as [Q, world * k] merge input (rank-major inside a query row).
    Works on any backend (NCCL over NVLink in production, gloo in the CPU tests).  The GPU path
    (ShardedCorpus.search) does the same all-gather but merges the receive buffer in place
    (ops.merge_gathered), without this function's permute + reshape copy."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    Q, k = idx.shape
    if bufs is None:
        send = torch.empty(2, Q, k, dtype=torch.int64, device=idx.device)
        recv = torch.empty(world * 2, Q, k, dtype=torch.int64, device=idx.device)
    else:
        send, recv = bufs
    if score64.data_ptr() != send[0].data_ptr():
        send[0].copy_(score64.contiguous().view(torch.int64))
    if idx.data_ptr() != send[1].data_ptr():
        send[1].copy_(idx)
    dist.all_gather_into_tensor(recv, send, group=group)  # output = inputs concatenated on dim 0
    recv = recv.view(world, 2, Q, k)
    s64 = recv[:, 0].view(torch.float64).permute(1, 0, 2).reshape(Q, world * k)
    rows = recv[:, 1].permute(1, 0, 2).reshape(Q, world * k)
    return s64, rows


class ShardedCorpus:
    """One rank's block of the corpus embedding matrix plus what the search needs with it."""

    def __init__(self, shard: torch.Tensor, idx_base: int = 0, group=None,
                 inv_norm: Optional[torch.Tensor] = None, bf16_shadow: bool = True, split_shadow: bool = False):
        if shard.dim() != 2:
            raise ValueError("shard must be [rows, D]")
        self.shard = shard
        self.idx_base = int(idx_base)
        self.group = group
        self.world = 1
        if group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(group)
        # inverse norms of the rows AS STORED, computed once (the reference recomputes the corpus
        # norms for every query: F.cosine_similarity at search_pipeline.py:77)
        self.inv_norm = inv_norm if inv_norm is not None else (
            ops.row_inv_norm(shard) if shard.shape[0] and shard.is_cuda else None)
        # fp32 / fp16 rows: a bf16 shadow (+50 % / +100 % memory) lets the tensor cores nominate candidates;
        # results are still defined on -- and re-scored in float64 from -- the original rows
        self.shadow = self.shadow_inv = None
        if (bf16_shadow and shard.is_cuda and shard.shape[0] and shard.dtype in (torch.float32, torch.float16)
                and shard.shape[1] % 8 == 0):
            self.shadow, self.shadow_inv = ops.make_shadow(shard)
        # ... and a split (hi + lo) shadow (+100 % / +200 %) carries 24 < k <= 100 on the tensor cores as
        # well (opt-in: without it those calls take the float64 scan)
        self.split = self.split_inv = None
        if split_shadow and shard.is_cuda and shard.shape[0] and shard.dtype in (torch.float32, torch.float16):
            self.split, self.split_inv = ops.make_shadow(shard, split=True)
        self._gather_buf = {}
        self._graphs = {}

    def _shadow_kw(self, k: int = 10):
        if self.split is not None and (k > 24 or self.shadow is None) and k <= 100:
            return {"corpus_shadow": self.split, "shadow_inv_norm": self.split_inv}
        return {} if self.shadow is None else {"corpus_shadow": self.shadow, "shadow_inv_norm": self.shadow_inv}

    # -- CUDA-graph replay of the local search -------------------------------------------------------
    # One search is ~8 dependent stream operations (threshold memset, sample pass, tighten, main pass,
    # select / re-score, two fallback kernels); replaying them as one captured graph removes the launch
    # gaps and the host-side planning, which matters for the ~1 ms small-batch (HBM-bound) searches.
    def _graph(self, Q: int, k: int, dtype: torch.dtype, exclude_self_base: int, mode: str):
        key = (Q, k, dtype, exclude_self_base, mode)
        g = self._graphs.get(key)
        if g is not None:
            return g
        dev = self.shard.device
        q_static = torch.zeros(Q, self.shard.shape[1], dtype=dtype, device=dev)
        # float64 scores and global rows land in the all-gather send buffer (one message: scores, then rows)
        send = torch.empty(2, Q, k, dtype=torch.int64, device=dev)
        recv = torch.empty(self.world * 2, Q, k, dtype=torch.int64, device=dev) if self.world > 1 else None
        s64, idx = send[0].view(torch.float64), send[1]
        scores = torch.empty(Q, k, dtype=torch.float32, device=dev)
        # The graph replays raw pointers: it owns its workspace (kept alive in self._graphs next to the graph),
        # never the growable per-stream cache of ops, whose buffers are replaced -- and freed -- when a later
        # call on a recycled stream handle needs more room.
        skw = self._shadow_kw(k) if mode == "auto" else {}
        is_split = bool(skw) and skw["corpus_shadow"] is self.split
        shadowed = bool(skw) and not is_split
        ws = torch.empty(ops.search_workspace_bytes(Q, self.shard.shape[0], self.shard.shape[1], k, dtype,
                                                    self.shard.dtype, mode, shadow=shadowed, device=dev, split=is_split),
                         dtype=torch.uint8, device=dev)

        def run():
            ops.search_topk(q_static, self.shard, k, corpus_inv_norm=self.inv_norm, idx_base=self.idx_base,
                            exclude_self_base=exclude_self_base, mode=mode, out_scores=scores,
                            out_score64=s64, out_idx=idx, workspace=ws, **skw)

        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            run()                      # warm-up on the capture stream (module load, function attributes)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            run()
        g = self._graphs[key] = (graph, q_static, scores, send, recv, ws)
        return g

    def search_graphed(self, queries: torch.Tensor, k: int, exclude_self_base: int = -1, mode: str = "auto"
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
        """`search` with the local search replayed from a CUDA graph captured on first use of this
        (batch size, k, dtype).  The returned tensors are the graph's static outputs: consume them
        before the next call with the same shape (or clone)."""
        Q = queries.shape[0]
        graph, q_static, scores, send, recv, _ws = self._graph(Q, k, queries.dtype, exclude_self_base, mode)
        q_static.copy_(queries, non_blocking=True)
        graph.replay()
        if self.world == 1:
            return scores, send[1]
        import torch.distributed as dist
        dist.all_gather_into_tensor(recv, send, group=self.group)
        merged, _, rows = ops.merge_gathered(recv, Q, k, self.world)    # reads the rank-major buffer in place
        return merged, rows

    def search_local(self, queries: torch.Tensor, k: int, **kw):
        return ops.search_topk(queries, self.shard, k, corpus_inv_norm=self.inv_norm,
                               idx_base=self.idx_base, **self._shadow_kw(k), **kw)

    def search(self, queries: torch.Tensor, k: int, exclude_self_base: int = -1, mode: str = "auto",
               return_score64: bool = False):
        """Global top-k on every rank: (scores float32 [Q, k], global rows int64 [Q, k]) and, with
        ``return_score64``, the float64 scores the ranking was made on."""
        if self.world == 1:
            return self.search_local(queries, k, exclude_self_base=exclude_self_base, mode=mode,
                                     return_score64=return_score64)
        Q = queries.shape[0]
        key = (Q, k)
        bufs = self._gather_buf.get(key)
        if bufs is None:
            send = torch.empty(2, Q, k, dtype=torch.int64, device=queries.device)
            recv = torch.empty(self.world * 2, Q, k, dtype=torch.int64, device=queries.device)
            bufs = self._gather_buf[key] = (send, recv)
        send, recv = bufs
        # the search kernels write their float64 scores and int64 rows straight into the send buffer
        ops.search_topk(queries, self.shard, k, corpus_inv_norm=self.inv_norm, idx_base=self.idx_base,
                        exclude_self_base=exclude_self_base, mode=mode,
                        out_score64=send[0].view(torch.float64), out_idx=send[1], **self._shadow_kw(k))
        import torch.distributed as dist
        dist.all_gather_into_tensor(recv, send, group=self.group)        # THE collective of the search
        scores, m64, rows = ops.merge_gathered(recv, Q, k, self.world)   # K3 reads the rank-major buffer in place
        return (scores, rows, m64) if return_score64 else (scores, rows)

    def search_host(self, host_queries: torch.Tensor, k: int, host_scores: torch.Tensor,
                    host_idx: torch.Tensor, graphed: bool = False) -> None:
        """End-to-end step: pinned host queries -> device, search, results -> pinned host."""
        if graphed:
            scores, rows = self.search_graphed(host_queries, k)       # H2D straight into the graph's input
        else:
            q = host_queries.to(self.shard.device, non_blocking=True)
            scores, rows = self.search(q, k)
        host_scores.copy_(scores, non_blocking=True)
        host_idx.copy_(rows, non_blocking=True)


# ---- all-pairs mining over one embedding matrix (BASELINE config 5; reference: cos_sim over the whole matrix,
# src/utils/metrics.py:469-507, and the paraphrase-mining consumers) ------------------------------------------
def all_pairs_corpus_sharded(corpus: ShardedCorpus, all_rows: torch.Tensor, k: int, tile: int = 16_384,
                             out_scores: Optional[torch.Tensor] = None, out_idx: Optional[torch.Tensor] = None
                             ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k neighbours of EVERY row (itself excluded) with the corpus side sharded: this rank scans its shard for
    every query tile of ``all_rows`` (the whole matrix, replicated), one all-gather + merge per tile.  Every rank
    ends with the full ([N, k] float32 scores, [N, k] int64 rows)."""
    n = all_rows.shape[0]
    dev = all_rows.device
    scores = out_scores if out_scores is not None else torch.empty(n, k, dtype=torch.float32, device=dev)
    idx = out_idx if out_idx is not None else torch.empty(n, k, dtype=torch.int64, device=dev)
    for b in range(0, n, tile):
        e = min(n, b + tile)
        s, i = corpus.search(all_rows[b:e], k, exclude_self_base=b)
        scores[b:e].copy_(s)
        idx[b:e].copy_(i)
    return scores, idx


def all_pairs_query_sharded(full: torch.Tensor, inv_norm: torch.Tensor, k: int, world: int, rank: int,
                            tile: int = 16_384) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """The same job with the QUERY side sharded and the corpus replicated ("replicas only", SURVEY.md 8e: a
    1M x 768 bf16 matrix is 1.5 GB): rank r answers rows [r0, r1) against the whole matrix, no collective at
    all; the result stays sharded by query.  Returns (scores [r1 - r0, k], rows [r1 - r0, k], r0)."""
    n = full.shape[0]
    r0, r1 = shard_bounds(n, world, rank)
    dev = full.device
    scores = torch.empty(r1 - r0, k, dtype=torch.float32, device=dev)
    idx = torch.empty(r1 - r0, k, dtype=torch.int64, device=dev)
    for b in range(r0, r1, tile):
        e = min(r1, b + tile)
        s, i = ops.search_topk(full[b:e], full, k, corpus_inv_norm=inv_norm, exclude_self_base=b)
        scores[b - r0:e - r0].copy_(s)
        idx[b - r0:e - r0].copy_(i)
    return scores, idx, r0
