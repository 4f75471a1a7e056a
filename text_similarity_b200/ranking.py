"""Consumers of the search path ("next" rows of SURVEY.md section 8f), built on the tensor-native
search API: bi-encoder retrieval + cross-encoder re-rank, the `most_similar_vectors` helper, and
nearest-centroid assignment (the k = 1 case of the same kernels)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import ops
from .pipeline import Pipeline, SentenceMiningPipeline


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


class ClusteringPipeline(Pipeline):
    """k-means over sentence embeddings with the reference's surface (src/pipeline/clustering.py:8-31:
    ``ClusteringPipeline(n_clusters, *args, method="k-means", **kwargs)``, ``_cluster(corpus)``,
    ``set_n_clusters(n)``, ``__call__(embeddings, n_clusters)``).  The reference delegates to scikit-learn's
    Euclidean KMeans on the host and then reads ``self.model.labels_`` (an attribute the encoder does not have) and
    returns nothing; the evident intent -- {cluster id: [corpus items]} -- is what this returns.  Here the
    assignment step is the search kernel itself (nearest centroid by cosine = top-1 of K2 with the centroids as the
    corpus, ``assign_to_centroids``) and the update step one ``index_add_``: spherical k-means, all on the GPU."""

    def __init__(self, n_clusters: int, *args, method: str = "k-means", max_iter: int = 50, seed: int = 0, **kwargs):
        super().__init__(*args, **kwargs)
        if method != "k-means":
            raise ValueError(f"unsupported clustering method {method!r} (the reference implements k-means only)")
        self.method = method
        self.n_clusters = int(n_clusters)
        self.max_iter = int(max_iter)
        self.seed = int(seed)
        self.labels_: Optional[torch.Tensor] = None
        self.cluster_centers_: Optional[torch.Tensor] = None

    def set_n_clusters(self, n: int):
        self.n_clusters = int(n)

    def fit(self, embeddings: torch.Tensor) -> torch.Tensor:
        """Cluster labels int64 [N].  Deterministic: seeded choice of the initial centroids, fixed iteration cap,
        stops when no label changes."""
        x = torch.nn.functional.normalize(embeddings.float(), dim=-1).contiguous()
        n = x.shape[0]
        kc = max(1, min(self.n_clusters, n))
        g = torch.Generator(device="cpu").manual_seed(self.seed)
        centers = x[torch.randperm(n, generator=g)[:kc].to(x.device)].clone()
        labels = torch.full((n,), -1, dtype=torch.int64, device=x.device)
        for _ in range(self.max_iter):
            new = assign_to_centroids(x, centers)
            if torch.equal(new, labels):
                break
            labels = new
            sums = torch.zeros_like(centers).index_add_(0, labels, x)
            empty = sums.norm(dim=-1) == 0
            centers = torch.where(empty[:, None], centers, torch.nn.functional.normalize(sums, dim=-1))
        self.labels_, self.cluster_centers_ = labels, centers
        return labels

    def _cluster(self, corpus) -> Dict[int, list]:
        items = corpus
        if isinstance(corpus, list):
            emb = self.encode_corpus(corpus)
        else:
            emb = torch.as_tensor(corpus)
        emb = emb if emb.is_cuda else emb.to(self.params.device)
        labels = self.fit(emb).cpu().tolist()
        results: Dict[int, list] = {}
        for text_id, cluster_id in enumerate(labels):
            results.setdefault(int(cluster_id), []).append(items[text_id])
        return results

    def __call__(self, embeddings, n_clusters: int):
        self.set_n_clusters(n_clusters)
        return self._cluster(embeddings)


def compare_models(queries: Sequence[str], teacher_results: dict, student_results: dict, verbose: bool = False) -> float:
    """Set-overlap agreement of two pipelines' search results, in percent: the share of the student's hits that
    are among the teacher's hits for the same query (reference src/evaluation/eval_sentence_mining.py:11-34, which
    prints every hit and the final figure; here the figure is returned and printing is opt-in).  Works on the
    dict-of-lists either search pipeline returns -- ``{query: [(row, item), ...]}`` or ``{query: [item, ...]}``."""
    agree = total = 0
    for qidx in teacher_results:
        t_hits = list(teacher_results[qidx])
        s_hits = list(student_results[qidx])
        if verbose:
            print(f"Hits for query: {queries[qidx]}")
        for i, s_hit in enumerate(s_hits):
            same = s_hit in t_hits
            agree += int(same)
            total += 1
            if verbose:
                print(f"Sentence {i + 1}: {s_hit}" if same else
                      f"Results for the query are different for student hit number: {i + 1}: {s_hit}")
    accuracy = 100.0 * agree / total if total else 0.0
    if verbose:
        print(f"Accuracy for student model compared to teacher: {accuracy}")
    return accuracy
