"""Search pipelines with the reference's surface (src/pipeline/search_pipeline.py:14-175), the
scoring and selection done by the CUDA search path (K2 + K3) through the C ABI.

Repairs of the reference's defects follow the evident intent: ``name=`` is
accepted (A9); ``__call__(queries, k)`` searches ``self.corpus`` (A2); chunks are sliced
``[i : i + chunk]`` (A4), searched with global row numbers and merged instead of overwritten (A7);
k is clamped by the corpus size (A6); results are best first, ties by lower index (A8).
"""
from __future__ import annotations

import os
import time
from typing import Dict, List, Optional, Tuple, Union

import torch
from torch import nn

from . import ops
from .config import SearchConfiguration

TextOrTensor = Union[List[str], torch.Tensor]


class Pipeline:
    def __init__(self, params: SearchConfiguration, model: nn.Module, name: Optional[str] = None):
        self.params = params
        self.model = model
        self.name = name

    def encode_corpus(self, documents: TextOrTensor, convert_to_numpy: bool = False):
        """list -> ``model.encode_text``; tensor -> passed through (reference :19-22)."""
        if isinstance(documents, list):
            return self.model.encode_text(documents, output_np=convert_to_numpy)
        return documents


class SearchPipeline(Pipeline):
    def __init__(self, *args, corpus: Optional[TextOrTensor] = None, **kwargs):
        super().__init__(*args, **kwargs)
        self.corpus = corpus

    def _index(self, corpus):
        raise NotImplementedError()

    def _search(self, queries: TextOrTensor, max_num_results: int):
        raise NotImplementedError()

    def __call__(self, queries, max_num_results):
        return self._search(queries, max_num_results)


class _EncodedCorpus:
    """Corpus rows as stored for search + their inverse norms, cached per corpus object."""

    def __init__(self, rows: torch.Tensor, inv_norm: torch.Tensor):
        self.rows, self.inv_norm = rows, inv_norm


class SentenceMiningPipeline(SearchPipeline):
    """Exact cosine top-k over the corpus, in chunks of ``corpus_chunk_size`` rows
    (reference :39-93).  Text is embedded by the model (unit-norm rows stored in
    ``params.corpus_dtype``, bf16 by default); tensors are searched as given (bf16 tensors take the
    tcgen05 path, fp32 tensors the float64 exact scan)."""

    def __init__(self, corpus_chunk_size: int, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.corpus_chunk_size = int(corpus_chunk_size)
        self._cache_key = None
        self._cache: List[Tuple[int, _EncodedCorpus]] = []
        self.last_corpus_encode_seconds = 0.0

    # -- encoding ---------------------------------------------------------------------------------
    def _dtype(self) -> torch.dtype:
        return getattr(self.params, "corpus_dtype", torch.bfloat16)

    def _encode_queries(self, queries: TextOrTensor) -> torch.Tensor:
        if isinstance(queries, list):
            if hasattr(self.model, "encode_text_normalized"):
                return self.model.encode_text_normalized(queries, self._dtype())[0]
            return self.encode_corpus(queries)
        q = queries if queries.dim() == 2 else queries.unsqueeze(0)
        return q.to(self.params.device) if not q.is_cuda else q

    def _corpus_chunks(self, corpus: TextOrTensor) -> List[Tuple[int, _EncodedCorpus]]:
        """[(first global row, encoded chunk)], cached while the same corpus object is searched."""
        key = (id(corpus), len(corpus), self.corpus_chunk_size)
        if key == self._cache_key:
            return self._cache
        chunks = []
        start_time = time.time()
        for begin in range(0, len(corpus), self.corpus_chunk_size):
            piece = corpus[begin:begin + self.corpus_chunk_size]  # reference :61, slice end repaired
            if isinstance(corpus, list):
                if hasattr(self.model, "encode_text_normalized"):
                    rows, inv = self.model.encode_text_normalized(piece, self._dtype())
                else:
                    rows = self.model.encode_text(piece)
                    inv = ops.row_inv_norm(rows)
            else:
                rows = piece if piece.is_cuda else piece.to(self.params.device)
                rows = rows if rows.stride(-1) == 1 else rows.contiguous()
                inv = ops.row_inv_norm(rows)
            chunks.append((begin, _EncodedCorpus(rows, inv)))
        self.last_corpus_encode_seconds = time.time() - start_time  # reference prints this (:65-71)
        self._cache_key, self._cache = key, chunks
        return chunks

    # -- search -----------------------------------------------------------------------------------
    def search_tensors(self, queries: TextOrTensor, max_num_results: int, corpus: Optional[TextOrTensor] = None
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Tensor-native result: (scores float32 [Q, k], corpus rows int64 [Q, k]) best first, ties by
        the lower row, k = min(max_num_results, len(corpus))."""
        if corpus is not None:
            self.corpus = corpus  # reference :58-59
        if self.corpus is None:
            raise ValueError("no corpus: pass corpus= to the pipeline or to _search")
        q = self._encode_queries(queries)
        chunks = self._corpus_chunks(self.corpus)
        k = max(1, min(int(max_num_results), len(self.corpus)))  # clamp by the corpus, not the queries (A6)
        mode = getattr(self.params, "search_mode", "auto")
        if len(chunks) == 1:
            begin, enc = chunks[0]
            qq = q if q.dtype == enc.rows.dtype or enc.rows.dtype == torch.float32 else q.to(enc.rows.dtype)
            return ops.search_topk(qq, enc.rows, k, corpus_inv_norm=enc.inv_norm, idx_base=begin, mode=mode)
        s64_parts, idx_parts = [], []
        for begin, enc in chunks:
            qq = q if q.dtype == enc.rows.dtype or enc.rows.dtype == torch.float32 else q.to(enc.rows.dtype)
            _, idx, s64 = ops.search_topk(qq, enc.rows, k, corpus_inv_norm=enc.inv_norm, idx_base=begin,
                                          mode=mode, return_score64=True)
            s64_parts.append(s64)
            idx_parts.append(idx)
        # cross-chunk merge, absent in the reference (:83,88 overwrite): K3 second pass
        group = max(1, 4096 // k)
        while len(s64_parts) > 1:
            ns, ni = [], []
            for g in range(0, len(s64_parts), group):
                ss, ii = s64_parts[g:g + group], idx_parts[g:g + group]
                _, m64, mi = ops.merge_topk(torch.cat(ss, 1), torch.cat(ii, 1), k, len(ss))
                ns.append(m64)
                ni.append(mi)
            s64_parts, idx_parts = ns, ni
        return s64_parts[0].float(), idx_parts[0]

    def _search(self, queries: TextOrTensor, corpus: Optional[TextOrTensor] = None,
                max_num_results: Optional[int] = None, return_embeddings: bool = False
                ) -> Dict[int, Union[torch.Tensor, List[Tuple[int, object]]]]:
        """Reference signature (:44-49).  Called as ``_search(queries, k)`` by ``__call__`` (:35-36),
        in which case the second positional is the result count and ``self.corpus`` is searched."""
        if max_num_results is None and isinstance(corpus, int):
            corpus, max_num_results = None, corpus
        if max_num_results is None:
            raise TypeError("_search() missing max_num_results")
        scores, idx = self.search_tensors(queries, max_num_results, corpus)
        idx_host = idx.cpu()
        top_candidates: Dict[int, Union[torch.Tensor, List[Tuple[int, object]]]] = {}
        for query_idx in range(idx_host.shape[0]):
            rows = idx_host[query_idx]
            rows = rows[rows >= 0]
            if return_embeddings:
                assert isinstance(self.corpus, torch.Tensor)  # reference :82
                top_candidates[query_idx] = self.corpus[rows.to(self.corpus.device)]
            else:
                top_candidates[query_idx] = [(int(c), self.corpus[int(c)]) for c in rows]
        return top_candidates

    def __call__(self, queries: TextOrTensor, max_num_results: int):
        return super().__call__(queries, max_num_results)


class SemanticSearchPipeline(SentenceMiningPipeline):
    """Same surface as the reference's HNSW pipeline (:96-175) -- ``index_path``, ``_index``,
    ``_search`` returning ``Dict[int, List[str]]`` best first, ``add_to_index`` /
    ``remove_from_index`` / ``num_indexed`` -- served by the EXACT engine: there is no ANN backend
    in this build (north_star), so ``ef`` / ``ef_construction`` / ``M`` are accepted and unused.
    The encoded corpus is persisted at ``index_path/index.pt`` (the reference persists index.bin)."""

    def __init__(self, index_path, *args, **kwargs):
        super().__init__(kwargs.pop("corpus_chunk_size", 1 << 30), *args, **kwargs)
        self.index_path = index_path
        self._removed = set()
        self._rows = self._inv = None
        saved = os.path.join(self.index_path, "index.pt")
        if os.path.exists(saved):
            blob = torch.load(saved, map_location=self.params.device)
            self._rows, self._inv = blob["rows"], blob["inv_norm"]
            self._removed = set(blob.get("removed", []))
        elif self.corpus is not None:
            self._index(self.corpus)

    def _index(self, corpus: List[str]):
        os.makedirs(self.index_path, exist_ok=True)
        self._rows, self._inv = self.model.encode_text_normalized(list(corpus), self._dtype())
        self._save()

    def _save(self):
        torch.save({"rows": self._rows, "inv_norm": self._inv, "removed": sorted(self._removed)},
                   os.path.join(self.index_path, "index.pt"))

    def _search(self, queries: TextOrTensor, max_num_results: int) -> Dict[int, List[str]]:
        q = self._encode_queries(queries)
        k = int(max_num_results)
        fetch = min(self._rows.shape[0], k + len(self._removed))  # over-fetch past tombstones
        _, idx = ops.search_topk(q.to(self._rows.dtype), self._rows, max(fetch, 1), corpus_inv_norm=self._inv)
        out: Dict[int, List[str]] = {}
        for qi, rows in enumerate(idx.cpu().tolist()):
            hits = [r for r in rows if r >= 0 and r not in self._removed][:k]
            out[qi] = [self.corpus[r] for r in hits]
        return out

    def __call__(self, queries: TextOrTensor, max_num_results: int):
        return self._search(queries, max_num_results)

    def add_to_index(self, text: Union[str, List[str]]):
        text = [text] if isinstance(text, str) else list(text)
        rows, inv = self.model.encode_text_normalized(text, self._dtype())
        self._rows = torch.cat([self._rows, rows]) if self._rows is not None else rows
        self._inv = torch.cat([self._inv, inv]) if self._inv is not None else inv
        self.corpus = list(self.corpus or []) + text

    def remove_from_index(self, ids):
        n = 0 if self._rows is None else self._rows.shape[0]
        for i in ids:
            if 0 <= int(i) < n:  # unknown ids are skipped, like the reference's try/except (:164-169)
                self._removed.add(int(i))

    def num_indexed(self):
        return (0 if self._rows is None else self._rows.shape[0]) - len(self._removed)
