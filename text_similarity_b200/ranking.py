"""Consumers of the search path ("next" rows of SURVEY.md section 8f), built on the tensor-native
search API: bi-encoder retrieval + cross-encoder re-rank, the `most_similar_vectors` helper, and
nearest-centroid assignment (the k = 1 case of the same kernels)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import ops
from .pipeline import SentenceMiningPipeline


class RankingPipeline(SentenceMiningPipeline):
    """Bi-encoder top-k followed by a cross-encoder re-rank (reference
    src/pipeline/ranking_pipeline.py:4-46, whose `_rank` does not parse: `[[query, el for el in
    to_rank]]` at :29).  Intent kept: per query, retrieve `top_k` passages with the exact
    cosine search, score (query, passage) pairs with `cross_encoder.predict`, return hits carrying both
    scores plus the mean cross score.  Unlike the reference, every query gets its own hits (the
    reference re-ranks query 0's hits for all queries, :22)."""

    def __init__(self, cross_encoder, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.cross_encoder = cross_encoder

    def _rank(self, queries: List[str], corpus: Sequence[str], top_k: int = 5, search_first: bool = True
              ) -> List[dict]:
        results = []
        if search_first:
            scores, rows = self.search_tensors(list(queries), top_k, corpus=list(corpus))
            scores, rows = scores.cpu().tolist(), rows.cpu().tolist()
        for qi, query in enumerate(queries):
            if search_first:
                hits = [{"corpus_id": r, "text": corpus[r], "score": s}
                        for r, s in zip(rows[qi], scores[qi]) if r >= 0]
            else:
                hits = [{"corpus_id": r, "text": t, "score": None} for r, t in enumerate(corpus)]
            cross_inp = [[query, h["text"]] for h in hits]
            cross_scores = [float(x) for x in self.cross_encoder.predict(cross_inp)] if cross_inp else []
            for h, cs in zip(hits, cross_scores):
                h["cross-score"] = cs
            hits.sort(key=lambda h: (-h["cross-score"], h["corpus_id"]))
            results.append({"results": hits, "cross_scores": cross_scores,
                            "avg_score": sum(cross_scores) / len(cross_scores) if cross_scores else 0.0})
        return results

    def __call__(self, queries: List[str], corpus: Sequence[str], top_k: int = 5, search_first: bool = True):
        return self._rank(queries, corpus, top_k, search_first)


def most_similar_vectors(vector: torch.Tensor, vectors: torch.Tensor, n: int = 1) -> List[torch.Tensor]:
    """The n rows of `vectors` most cosine-similar to `vector`, best first (reference
    src/utils/utils.py:96-106: cosine_similarity on an expanded view + a Python sort of N tuples)."""
    q = vector if vector.dim() == 2 else vector.unsqueeze(0)
    _, idx = ops.search_topk(q.contiguous(), vectors.contiguous(), min(n, vectors.shape[0]))
    return [vectors[i] for i in idx[0].tolist() if i >= 0]


def assign_to_centroids(embeddings: torch.Tensor, centroids: torch.Tensor) -> torch.Tensor:
    """Nearest centroid by cosine for every row (the assignment step of spherical k-means over the
    embeddings `ClusteringPipeline` consumes, reference src/pipeline/clustering.py:8-31): the search
    with the centroids as the corpus and k = 1.  Returns int64 [N]."""
    _, idx = ops.search_topk(embeddings.contiguous(), centroids.contiguous(), 1)
    return idx[:, 0]


def near_duplicates(embeddings: torch.Tensor, threshold: float = 0.95, k: int = 5,
                    inv_norm: Optional[torch.Tensor] = None, tile: int = 16_384) -> Dict[int, List[int]]:
    """Paraphrase / near-duplicate mining over one embedding matrix (BASELINE config 5): for every
    row the (up to k) other rows with cosine >= threshold, from the all-pairs search with the row
    itself excluded.  Returns {row: [neighbour rows, best first]} for rows that have any."""
    n = embeddings.shape[0]
    inv = inv_norm if inv_norm is not None else ops.row_inv_norm(embeddings)
    out: Dict[int, List[int]] = {}
    for b in range(0, n, tile):
        e = min(n, b + tile)
        s, idx = ops.search_topk(embeddings[b:e], embeddings, k, corpus_inv_norm=inv, exclude_self_base=b)
        keep = (s >= threshold) & (idx >= 0)
        rows = keep.any(dim=1).nonzero().flatten().tolist()
        idx_h, keep_h = idx.cpu(), keep.cpu()
        for r in rows:
            out[b + r] = idx_h[r][keep_h[r]].tolist()
    return out
