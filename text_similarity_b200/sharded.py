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
    Works on any backend (NCCL over NVLink in production, gloo in the CPU tests)."""
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
                 inv_norm: Optional[torch.Tensor] = None):
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
        self._gather_buf = {}

    def search_local(self, queries: torch.Tensor, k: int, **kw):
        return ops.search_topk(queries, self.shard, k, corpus_inv_norm=self.inv_norm,
                               idx_base=self.idx_base, **kw)

    def search(self, queries: torch.Tensor, k: int, exclude_self_base: int = -1, mode: str = "auto"
               ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Global top-k on every rank: (scores float32 [Q, k], global rows int64 [Q, k])."""
        if self.world == 1:
            return self.search_local(queries, k, exclude_self_base=exclude_self_base, mode=mode)
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
                        out_score64=send[0].view(torch.float64), out_idx=send[1])
        s64, idx = gather_shard_results(send[0].view(torch.float64), send[1], self.group, bufs)
        scores, _, rows = ops.merge_topk(s64, idx, k, self.world)
        return scores, rows

    def search_host(self, host_queries: torch.Tensor, k: int, host_scores: torch.Tensor,
                    host_idx: torch.Tensor) -> None:
        """End-to-end step: pinned host queries -> device, search, results -> pinned host."""
        q = host_queries.to(self.shard.device, non_blocking=True)
        scores, rows = self.search(q, k)
        host_scores.copy_(scores, non_blocking=True)
        host_idx.copy_(rows, non_blocking=True)
