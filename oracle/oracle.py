"""CPU oracle for the semantic-search hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the arithmetic of the reference's hot path
(cr1m5onk1ng/text_similarity).  It is the checker the CUDA path is compared with.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package
(``text_similarity_b200``) never imports anything from ``oracle/``.

Parity pinning: the reference ships no tests, fixtures or golden vectors
(SURVEY.md section 4 / 8c).  The oracle is therefore pinned against outputs of the
reference's own code executed in the build container:
``tests/golden/make_golden.py`` imports ``/root/reference`` (with stub modules for
its unavailable third-party imports), runs ``AvgPoolingStrategy.forward``,
``OnnxSentenceTransformerWrapper.forward`` and ``cos_sim``, plus the literal ATen
calls of ``SentenceMiningPipeline._search`` (``F.cosine_similarity`` + ``torch.topk``),
and commits the inputs/outputs as ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks every function below against them.

Two families of functions live here:

* ``*_literal`` -- the reference's lines restated one for one in fp32 torch
  (what the reference computes, including its unspecified top-k order).
* ``*_exact``   -- the same mathematical quantity computed in float64 from the
  *stored* inputs (bf16/fp8/fp32 values are exactly representable in float64,
  and so are their pairwise products), ranked by (score descending, index
  ascending).  This is the deterministic statement of "exact cosine top-k with
  ties broken by the lower index" (BASELINE.json north_star) and is what the
  CUDA path must match index for index.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

COS_EPS = 1e-8  # F.cosine_similarity default eps (reference: search_pipeline.py:77)
POOL_EPS = 1e-9  # clamp in AvgPoolingStrategy (reference: modules.py:168)


# --------------------------------------------------------------------------- pooling
def mean_pool_literal(embeddings: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """Masked mean pooling, restating reference src/modules/modules.py:158-171
    (identical arithmetic: src/models/sentence_encoder.py:35-38).

    embeddings [B, L, D] float, attention_mask [B, L] (int64 0/1 in the reference).
    Returns [B, D] in the promoted dtype (fp32 for fp32/fp16 inputs times a float mask).
    """
    assert embeddings.dim() == 3  # modules.py:159
    mask = attention_mask.unsqueeze(-1).expand(embeddings.size()).float()  # modules.py:162
    summed = torch.sum(embeddings * mask, 1)  # modules.py:165
    count = torch.clamp(mask.sum(1), min=POOL_EPS)  # modules.py:168
    return summed / count  # modules.py:170


def mean_pool_exact(embeddings: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """float64 evaluation of the same formula (cross-check of the literal version)."""
    e = embeddings.to(torch.float64)
    m = attention_mask.to(torch.float64).unsqueeze(-1)
    return (e * m).sum(1) / torch.clamp(m.sum(1), min=POOL_EPS)


def l2_normalize_exact(x: torch.Tensor, eps: float = COS_EPS) -> torch.Tensor:
    """Row-wise x / max(||x||, eps) in float64 -- the normalisation that
    ``F.cosine_similarity`` (search_pipeline.py:77) and ``cos_sim`` (metrics.py:99-100)
    apply inside the similarity; the build hoists it into the pooling kernel."""
    x = x.to(torch.float64)
    n = x.norm(dim=-1, keepdim=True).clamp_min(eps)
    return x / n


def pool_normalize_cast(embeddings: torch.Tensor, attention_mask: torch.Tensor,
                        out_dtype: torch.dtype = torch.float32, normalize: bool = True
                        ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Oracle of the fused pooling kernel: mean-pool (modules.py:160-170) ->
    optional L2 normalise -> cast to ``out_dtype`` (fp32 / bf16 / float8_e4m3fn).

    Returns (rows in out_dtype, inv_norm float32 [B]) where ``inv_norm`` is
    1 / max(||stored row||, eps) of the row *as stored* (after rounding), which is the
    factor the search needs to turn a dot product of stored rows into their cosine.
    """
    pooled = mean_pool_exact(embeddings, attention_mask)
    if normalize:
        pooled = l2_normalize_exact(pooled)
    stored = pooled.to(torch.float32).to(out_dtype)
    inv = 1.0 / stored.to(torch.float64).norm(dim=-1).clamp_min(COS_EPS)
    return stored, inv.to(torch.float32)


# --------------------------------------------------------------------------- similarity
def cos_sim_literal(a, b) -> torch.Tensor:
    """All-pairs cosine matrix, restating reference src/utils/metrics.py:81-101
    (no epsilon: a zero row yields NaN, as in the reference)."""
    if not isinstance(a, torch.Tensor):
        a = torch.tensor(a)  # metrics.py:87-88
    if not isinstance(b, torch.Tensor):
        b = torch.tensor(b)  # metrics.py:90-91
    if a.dim() == 1:
        a = a.unsqueeze(0)  # metrics.py:93-94
    if b.dim() == 1:
        b = b.unsqueeze(0)  # metrics.py:96-97
    a_n = a / a.norm(dim=-1)[:, None]  # metrics.py:99
    b_n = b / b.norm(dim=-1)[:, None]  # metrics.py:100
    return torch.mm(a_n, b_n.transpose(0, 1))  # metrics.py:101


def query_scores_literal(query: torch.Tensor, corpus: torch.Tensor) -> torch.Tensor:
    """One query against the corpus chunk, restating search_pipeline.py:76-77."""
    q = query.unsqueeze(0).expand_as(corpus)  # :76
    return F.cosine_similarity(q, corpus, dim=-1)  # :77


def search_literal(queries: torch.Tensor, corpus: torch.Tensor, k: int
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    """The reference's per-query loop (search_pipeline.py:73-79) with the defects of
    SURVEY.md Appendix A repaired (A5: dim of topk, A6: k clamp by corpus size).
    fp32 on CPU.  Order inside a row is ``torch.topk``'s (unspecified, sorted=False)."""
    queries = queries.float()
    corpus = corpus.float()
    k = min(k, corpus.shape[0])
    vals, idxs = [], []
    for q in queries:  # :73
        scores = query_scores_literal(q, corpus)
        top = torch.topk(scores, k, dim=0, sorted=False, largest=True)  # :78
        vals.append(top[0])
        idxs.append(top[1])  # :79
    return torch.stack(vals), torch.stack(idxs)


def search_cos_sim_literal(queries: torch.Tensor, corpus: torch.Tensor, k: int,
                           chunk: int = 1 << 18) -> Tuple[torch.Tensor, torch.Tensor]:
    """The reference's batched formulation: ``cos_sim`` (metrics.py:99-101) followed by
    ``torch.topk`` (search_pipeline.py:78), corpus processed in row chunks so the
    [Q, chunk] score block fits in RAM.  Used as the fair CPU baseline."""
    queries = queries.float()
    k = min(k, corpus.shape[0])
    qn = queries / queries.norm(dim=-1)[:, None]
    best_v: Optional[torch.Tensor] = None
    best_i: Optional[torch.Tensor] = None
    for s in range(0, corpus.shape[0], chunk):
        c = corpus[s:s + chunk].float()
        cn = c / c.norm(dim=-1)[:, None]
        sc = torch.mm(qn, cn.transpose(0, 1))
        v, i = torch.topk(sc, min(k, sc.shape[1]), dim=1, largest=True)
        i = i + s
        if best_v is None:
            best_v, best_i = v, i
        else:
            v = torch.cat([best_v, v], 1)
            i = torch.cat([best_i, i], 1)
            best_v, sel = torch.topk(v, k, dim=1, largest=True)
            best_i = torch.gather(i, 1, sel)
    return best_v, best_i


def cosine_scores_exact(queries: torch.Tensor, corpus: torch.Tensor) -> torch.Tensor:
    """[Q, N] float64 cosine of the stored values:
    dot / (max(||q||, eps) * max(||c||, eps))  (F.cosine_similarity's formula, eps=1e-8)."""
    q = queries.to(torch.float64)
    c = corpus.to(torch.float64)
    qn = q.norm(dim=-1).clamp_min(COS_EPS)
    cn = c.norm(dim=-1).clamp_min(COS_EPS)
    return (q @ c.T) / (qn[:, None] * cn[None, :])


def _rank_rows(scores: torch.Tensor, k: int, base: int = 0
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k of each row by (score descending, index ascending): a stable descending
    sort keeps equal scores in index order."""
    order = torch.sort(scores, dim=1, descending=True, stable=True)[1][:, :k]
    return torch.gather(scores, 1, order), order + base


def search_exact(queries: torch.Tensor, corpus: torch.Tensor, k: int,
                 idx_base: int = 0, exclude_self_base: int = -1,
                 chunk: int = 1 << 17) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exact cosine top-k in float64 with the north_star tie rule.

    Returns (scores float64 [Q, k'], idx int64 [Q, k']) best first, k' = min(k, N)
    (minus one when ``exclude_self_base`` removes a row).  ``idx_base`` is added to
    corpus row numbers (contiguous row sharding).  If ``exclude_self_base >= 0`` the
    corpus row whose global index equals ``exclude_self_base + query_number`` is
    skipped (all-pairs mining: a sentence is not its own neighbour)."""
    Q, N = queries.shape[0], corpus.shape[0]
    kk = min(k, N)
    best_v = torch.empty(Q, 0, dtype=torch.float64)
    best_i = torch.empty(Q, 0, dtype=torch.int64)
    q64 = queries.to(torch.float64)
    qn = q64.norm(dim=-1).clamp_min(COS_EPS)
    for s in range(0, N, chunk):
        c = corpus[s:s + chunk].to(torch.float64)
        cn = c.norm(dim=-1).clamp_min(COS_EPS)
        sc = (q64 @ c.T) / (qn[:, None] * cn[None, :])
        if exclude_self_base >= 0:
            rows = torch.arange(Q) + exclude_self_base - idx_base - s
            ok = (rows >= 0) & (rows < c.shape[0])
            sc[torch.arange(Q)[ok], rows[ok]] = -float("inf")
        v, i = _rank_rows(sc, min(kk, sc.shape[1]), base=s + idx_base)
        # chunks arrive in index order, so concatenating keeps ties index-ordered
        v = torch.cat([best_v, v], 1)
        i = torch.cat([best_i, i], 1)
        order = torch.sort(v, dim=1, descending=True, stable=True)[1][:, :kk]
        best_v = torch.gather(v, 1, order)
        best_i = torch.gather(i, 1, order)
    if exclude_self_base >= 0:
        keep = torch.isfinite(best_v).all(dim=0)
        best_v, best_i = best_v[:, keep], best_i[:, keep]
    return best_v, best_i


def merge_topk_exact(scores: torch.Tensor, idx: torch.Tensor, k: int
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge candidate lists [Q, n_lists * k_in] -> top-k by (score desc, index asc).
    Entries with index < 0 are padding.  (The reference has no merge -- its chunk loop
    overwrites results, search_pipeline.py:83,88 -- this is the repaired intent, A7.)"""
    s = scores.to(torch.float64).clone()
    s[idx < 0] = -float("inf")
    # sort by index first, then stable by score: equal scores stay index-ascending
    by_idx = torch.sort(idx, dim=1, stable=True)[1]
    s = torch.gather(s, 1, by_idx)
    ix = torch.gather(idx, 1, by_idx)
    order = torch.sort(s, dim=1, descending=True, stable=True)[1][:, :k]
    return torch.gather(s, 1, order), torch.gather(ix, 1, order)


def near_tie_mask(scores_exact_row: np.ndarray, tol: float) -> np.ndarray:
    """Positions of a best-first float64 score row whose neighbour is closer than ``tol``
    -- the only places where an fp32 evaluation may legitimately order rows differently."""
    d = np.abs(np.diff(scores_exact_row))
    m = np.zeros(scores_exact_row.shape[0], dtype=bool)
    m[:-1] |= d < tol
    m[1:] |= d < tol
    return m
