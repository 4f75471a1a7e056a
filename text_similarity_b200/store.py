"""Corpus embedding store: the rows the exact search scans, resident in HBM, with the
add / remove / count surface of the reference's index and on-disk persistence.

The reference persists only its HNSW graph (``index.bin``, src/pipeline/search_pipeline.py:106-109,122)
and edits it through ``add_to_index`` / ``remove_from_index`` / ``num_indexed`` (:154-175).  The exact
engine needs no graph: the "index" IS the embedding matrix, so this store keeps

* ``rows``      [capacity, D]  bf16 / e4m3 / fp32, unit-norm rows as written by the pooling kernel K1,
* ``inv_norm``  [capacity]     float32, 1 / ||row as stored|| (K1 emits it; the search consumes it),
* ``ids``       [capacity]     int64 caller-visible labels (default: insertion order),

dense in ``[0, n)``: removal moves the last live row into the hole (no tombstones, so the search kernels
never see a deleted row and k is never inflated), growth doubles the capacity.  Search returns labels.

On disk a store is a directory of raw little-endian files (``rows.bin``, ``inv_norm.bin``, ``ids.bin``)
plus ``meta.json``; ``load`` memory-maps them and streams them to the GPU through a pair of pinned
staging buffers, so a 15 GB shard never exists twice in host memory.  ``save(..., world, rank)`` /
``load(..., world, rank)`` write and read one contiguous row block per rank (the sharding of
text_similarity_b200.sharded).
"""
from __future__ import annotations

import json
import os
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import ops

_DTYPES = {"bfloat16": torch.bfloat16, "float32": torch.float32, "float16": torch.float16,
           "float8_e4m3fn": torch.float8_e4m3fn}
_STAGE_BYTES = 64 << 20


def _dtype_name(dt: torch.dtype) -> str:
    for name, t in _DTYPES.items():
        if t == dt:
            return name
    raise ValueError(f"unsupported store dtype {dt}")


class EmbeddingStore:
    def __init__(self, dim: int, dtype: torch.dtype = torch.bfloat16, device: Union[str, torch.device] = "cuda",
                 capacity: int = 1024):
        _dtype_name(dtype)
        self.dim = int(dim)
        self.dtype = dtype
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("EmbeddingStore lives in GPU memory: there is no CPU fallback")
        self.n = 0
        self._next_id = 0
        cap = max(int(capacity), 1)
        self.rows = torch.empty(cap, self.dim, dtype=dtype, device=self.device)
        self.inv_norm = torch.empty(cap, dtype=torch.float32, device=self.device)
        self.ids = torch.empty(cap, dtype=torch.int64, device=self.device)
        self._pos = {}          # label -> row position (host side; built lazily: only removal and explicit ids need it)
        self._pos_valid = True

    # -- size -------------------------------------------------------------------------------------
    def __len__(self) -> int:
        return self.n

    num_indexed = __len__      # the reference's name (search_pipeline.py:174-175)

    @property
    def capacity(self) -> int:
        return self.rows.shape[0]

    def _reserve(self, extra: int) -> None:
        need = self.n + extra
        if need <= self.capacity:
            return
        cap = max(need, 2 * self.capacity)
        for name in ("rows", "inv_norm", "ids"):
            old = getattr(self, name)
            new = torch.empty((cap,) + tuple(old.shape[1:]), dtype=old.dtype, device=self.device)
            new[:self.n] = old[:self.n]
            setattr(self, name, new)

    def _positions(self) -> dict:
        """label -> row position; rebuilt from the device-side ids after bulk appends (a Python dict entry per
        row is too slow to maintain eagerly for 10M-row stores that never remove anything)."""
        if not self._pos_valid:
            self._pos = {int(lab): i for i, lab in enumerate(self.ids[:self.n].tolist())}
            self._pos_valid = True
        return self._pos

    def _take_ids(self, count: int, ids: Optional[Sequence[int]]) -> torch.Tensor:
        if ids is None:
            out = torch.arange(self._next_id, self._next_id + count, dtype=torch.int64)   # fresh labels: no clash possible
        else:
            out = torch.as_tensor(list(ids) if not isinstance(ids, torch.Tensor) else ids.cpu(), dtype=torch.int64)
            if out.numel() != count:
                raise ValueError("one id per row")
            pos = self._positions()
            if len(set(out.tolist())) != count or any(int(i) in pos for i in out.tolist()):
                raise ValueError("ids must be unique and not already in the store")
        if count:
            self._next_id = max(self._next_id, int(out.max()) + 1)
        return out

    def _register(self, ids: torch.Tensor) -> None:
        self.ids[self.n:self.n + ids.numel()] = ids.to(self.device)
        self.n += ids.numel()
        self._pos_valid = False

    # -- add --------------------------------------------------------------------------------------
    def add(self, rows: torch.Tensor, inv_norm: Optional[torch.Tensor] = None,
            ids: Optional[Sequence[int]] = None) -> torch.Tensor:
        """Append already-pooled rows ([m, D], any float dtype; cast to the store dtype as they are --
        cosine is scale free, so rows need not be unit norm).  Returns their labels."""
        if rows.dim() != 2 or rows.shape[1] != self.dim:
            raise ValueError(f"rows must be [m, {self.dim}]")
        m = rows.shape[0]
        lab = self._take_ids(m, ids)
        self._reserve(m)
        dst = self.rows[self.n:self.n + m]
        dst.copy_(rows.to(self.device))
        self.inv_norm[self.n:self.n + m] = inv_norm.to(self.device) if inv_norm is not None and rows.dtype == self.dtype \
            else ops.row_inv_norm(dst)
        self._register(lab)
        return lab

    def add_tokens(self, token_embeddings: torch.Tensor, attention_mask: torch.Tensor,
                   ids: Optional[Sequence[int]] = None, order: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Encode-side fusion: K1 mean-pools, normalises, casts and writes the
        batch's sentence rows and inverse norms STRAIGHT into the store's tail -- no intermediate
        [B, D] tensor, no Python-side stacking (reference sentence_encoder.py:167-173).
        ``order[b]`` = position of batch row b among the appended rows (the un-sort of a
        length-sorted batch); default: batch order."""
        B = token_embeddings.shape[0]
        lab = self._take_ids(B, ids)
        self._reserve(B)
        pos = torch.arange(B, device=self.device) if order is None else order.to(self.device)
        ops.pool_norm(token_embeddings, attention_mask, normalize=True, out=self.rows,
                      out_rows=pos + self.n, out_inv_norm=self.inv_norm)
        self._register(lab)
        return lab

    # -- remove -----------------------------------------------------------------------------------
    def remove(self, ids: Sequence[int]) -> int:
        """Delete rows by label; unknown labels are skipped like the reference's try/except
        (search_pipeline.py:164-169).  The last live row moves into each hole.  Returns the count."""
        done = 0
        pos = self._positions()
        for lab in ids:
            lab = int(lab)
            p = pos.pop(lab, None)
            if p is None:
                continue
            last = self.n - 1
            if p != last:
                self.rows[p] = self.rows[last]
                self.inv_norm[p] = self.inv_norm[last]
                moved = int(self.ids[last])
                self.ids[p] = moved
                pos[moved] = p
            self.n -= 1
            done += 1
        return done

    # -- search -----------------------------------------------------------------------------------
    def search(self, queries: torch.Tensor, k: int, mode: str = "auto", return_positions: bool = False
               ) -> Tuple[torch.Tensor, torch.Tensor]:
        """(scores float32 [Q, k'], labels int64 [Q, k']), k' = min(k, len(store)), best first; ties by the
        lower ROW POSITION (insertion order until a removal moves a row)."""
        if self.n == 0:
            Q = queries.shape[0]
            return (torch.empty(Q, 0, dtype=torch.float32, device=self.device),
                    torch.empty(Q, 0, dtype=torch.int64, device=self.device))
        k = max(1, min(int(k), self.n))
        q = queries.to(self.device)
        if q.dtype != self.dtype and self.dtype != torch.float32:
            q = q.to(self.dtype)
        scores, pos = ops.search_topk(q, self.rows[:self.n], k, corpus_inv_norm=self.inv_norm[:self.n], mode=mode)
        if return_positions:
            return scores, pos
        # positions past the last row the kernels could return (-1: fewer than k finite-scoring rows) stay -1
        # instead of wrapping around to the last live row's label
        labels = torch.where(pos >= 0, self.ids[:self.n][pos.clamp_min(0)], torch.full_like(pos, -1))
        return scores, labels

    # -- persistence ------------------------------------------------------------------------------
    def save(self, path: str, world: int = 1, rank: int = 0) -> None:
        """Write this store (one rank's row block) under ``path``.  Device -> pinned staging buffer ->
        file, 64 MB at a time."""
        os.makedirs(path, exist_ok=True)
        sfx = "" if world == 1 else f".{rank:03d}-of-{world:03d}"
        for name, t in (("rows", self.rows[:self.n]), ("inv_norm", self.inv_norm[:self.n]), ("ids", self.ids[:self.n])):
            raw = t.contiguous().view(torch.uint8).reshape(-1)
            stage = torch.empty(min(_STAGE_BYTES, max(raw.numel(), 1)), dtype=torch.uint8).pin_memory()
            with open(os.path.join(path, f"{name}{sfx}.bin"), "wb") as f:
                for b in range(0, raw.numel(), stage.numel()):
                    e = min(raw.numel(), b + stage.numel())
                    stage[:e - b].copy_(raw[b:e])
                    torch.cuda.current_stream(self.device).synchronize()
                    f.write(stage[:e - b].numpy().tobytes())
        meta = {"format": 1, "dim": self.dim, "dtype": _dtype_name(self.dtype), "rows": self.n,
                "next_id": self._next_id, "world": world, "rank": rank}
        with open(os.path.join(path, f"meta{sfx}.json"), "w") as f:
            json.dump(meta, f)

    @classmethod
    def load(cls, path: str, device: Union[str, torch.device] = "cuda", world: int = 1, rank: int = 0,
             spare: int = 0) -> "EmbeddingStore":
        """mmap -> two pinned staging buffers -> HBM (the copy of chunk i overlaps the read of i+1)."""
        sfx = "" if world == 1 else f".{rank:03d}-of-{world:03d}"
        with open(os.path.join(path, f"meta{sfx}.json")) as f:
            meta = json.load(f)
        if meta.get("format") != 1:
            raise ValueError(f"unknown store format in {path}")
        st = cls(meta["dim"], _DTYPES[meta["dtype"]], device, capacity=max(1, meta["rows"] + spare))
        n = meta["rows"]
        copy_stream = torch.cuda.Stream(st.device)
        for name, t in (("rows", st.rows), ("inv_norm", st.inv_norm), ("ids", st.ids)):
            dst = t[:n].view(torch.uint8).reshape(-1) if n else None
            nbytes = 0 if dst is None else dst.numel()
            fn = os.path.join(path, f"{name}{sfx}.bin")
            if os.path.getsize(fn) != nbytes:
                raise ValueError(f"{fn}: {os.path.getsize(fn)} bytes on disk, {nbytes} expected")
            if nbytes == 0:
                continue
            src = np.memmap(fn, dtype=np.uint8, mode="r")
            stages = [torch.empty(min(_STAGE_BYTES, nbytes), dtype=torch.uint8).pin_memory() for _ in range(2)]
            events = [None, None]
            for i, b in enumerate(range(0, nbytes, stages[0].numel())):
                e = min(nbytes, b + stages[0].numel())
                s = i & 1
                if events[s] is not None:
                    events[s].synchronize()          # the previous copy out of this buffer has finished
                stages[s][:e - b].numpy()[:] = src[b:e]     # page cache -> pinned staging buffer
                with torch.cuda.stream(copy_stream):
                    dst[b:e].copy_(stages[s][:e - b], non_blocking=True)
                    events[s] = torch.cuda.Event()
                    events[s].record(copy_stream)
            copy_stream.synchronize()
            del src
        st.n = n
        st._next_id = meta["next_id"]
        st._pos_valid = False
        return st
