"""Sentence-embedding wrappers with the reference's surface (src/models/sentence_encoder.py,
src/models/modeling.py:11-87).  The HF encoder stack stays stock PyTorch (random-init weights are
fine); pooling, normalisation and the storage cast run in the fused CUDA kernel K1.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple, Union

import numpy as np
import torch
from torch import nn

from .config import Configuration
from .features import EmbeddingsFeatures
from .pooling import AvgPoolingStrategy, PoolingStrategy


class BaseEncoderModel(nn.Module):
    """Holds ``params`` and the HF ``context_embedder`` (reference modeling.py:11-26)."""

    def __init__(self, params: Configuration, context_embedder: nn.Module,
                 input_dict: bool = False, normalize: bool = False):
        super().__init__()
        self.params = params
        self.normalize = normalize
        self.input_dict = input_dict
        self.context_embedder = context_embedder

    @property
    def model_name(self):
        return self.params.model_parameters.model_name

    @property
    def config(self):
        return self.context_embedder.config

    @property
    def embedding_size(self):
        cfg = self.config
        return getattr(cfg, "dim", None) if "distilbert" in str(self.params.model) else cfg.hidden_size

    @property
    def params_num(self):
        return sum(p.numel() for p in self.context_embedder.parameters() if p.requires_grad)

    def save_pretrained(self, path):  # reference modeling.py:52-59
        assert path is not None
        os.makedirs(path, exist_ok=True)
        self.context_embedder.save_pretrained(path)
        if getattr(self.params, "tokenizer", None) is not None and hasattr(self.params.tokenizer, "save_pretrained"):
            self.params.tokenizer.save_pretrained(path)
        torch.save(self.params, os.path.join(path, "model_config.bin"))

    def forward(self, *args, **kwargs):
        raise NotImplementedError()

    def encode(self, *args, **kwargs):
        raise NotImplementedError()


class OnnxSentenceTransformerWrapper(BaseEncoderModel):
    """Inference-only bi-encoder with the pooler fixed to masked mean (reference
    sentence_encoder.py:17-39); the five pooling lines (:35-38) are one K1 launch."""

    def __init__(self, *args, projection: Optional[nn.Module] = None, **kwargs):
        super().__init__(*args, **kwargs)
        self.projection = projection if projection is not None else nn.Identity()

    def forward(self, input_ids, attention_mask, **kwargs):
        token_embeddings = self.context_embedder(input_ids=input_ids, attention_mask=attention_mask, **kwargs)[0]
        token_embeddings = self.projection(token_embeddings)  # :34 (projects tokens, then pools)
        from . import ops
        pooled, _ = ops.pool_norm(token_embeddings, attention_mask, out_dtype=torch.float32, normalize=False)
        return pooled


class SentenceTransformerWrapper(BaseEncoderModel):
    """SBERT-style bi-encoder wrapper (reference sentence_encoder.py:72-217): same constructor,
    ``encode`` / ``encode_text`` / ``get_sentence_embedding_dimension``.  Training (``forward`` with a
    loss, merge strategies) is outside the accelerated path and is not rebuilt here."""

    def __init__(self, pooler: Optional[PoolingStrategy] = None, merge_strategy=None, loss=None, *args,
                 parallel_mode: bool = True, projection: Optional[nn.Module] = None, **kwargs):
        super().__init__(*args, **kwargs)
        self.pooler = pooler if pooler is not None else AvgPoolingStrategy(self.params)
        self.merge_strategy = merge_strategy
        self.loss = loss
        self.parallel_mode = parallel_mode
        self.projection = projection if projection is not None else nn.Identity()

    # ---- the intent of reference :161 (`self.encode(features, parallel_mode=False)`, which raises
    # as written): the non-parallel branch of forward (:115-124) without the loss
    def embed(self, features: EmbeddingsFeatures) -> torch.Tensor:
        tokens = self._token_embeddings(features)
        return self.projection(self.pooler(tokens, features))

    def forward(self, features, return_output=False, head_mask=None):
        if self.parallel_mode or self.loss is not None:
            raise NotImplementedError(
                "training forward (pair merge + loss, reference sentence_encoder.py:100-131) is outside "
                "the search hot path; use embed()/encode_text() for inference")
        return self.embed(features)

    def encode(self, documents, output_np: bool = False):
        if isinstance(documents, EmbeddingsFeatures):
            with torch.no_grad():
                return self.embed(documents)
        return self.encode_text(documents, output_np)  # :133-134

    def _batches(self, documents: List[str]):
        """Mini-batches as (original positions int64 [b], features).  Default: the reference's loop -- sort by
        character length, fixed batches of ``params.batch_size``, tokenised per batch exactly like
        sentence_encoder.py:138-159.  With ``params.token_budget``: length-bucketed batching (below)."""
        if getattr(self.params, "token_budget", None):
            yield from self._token_budget_batches(documents)
            return
        order = np.argsort([len(sen) for sen in documents], kind="stable")
        bs = self.params.batch_size
        for start in range(0, len(documents), bs):
            rows = order[start:start + bs]
            enc = self.params.tokenizer(
                text=[documents[i] for i in rows],
                add_special_tokens=True,
                padding="longest",
                truncation=True,
                max_length=self.params.sequence_max_len,
                return_attention_mask=True,
                return_token_type_ids=False,
                return_tensors="pt",
            )
            feats = EmbeddingsFeatures(input_ids=enc["input_ids"].to(self.params.device),
                                       attention_mask=enc["attention_mask"].to(self.params.device))
            yield torch.as_tensor(rows, dtype=torch.int64), feats

    def _token_budget_batches(self, documents: List[str]):
        """Length-bucketed batching: the corpus is tokenised ONCE without padding, sorted by
        TOKEN count, and cut into batches whose padded size (sentences x longest sentence) stays within
        ``params.token_budget`` tokens -- hundreds of short sentences or a handful of long ones per encoder call,
        instead of the reference's fixed 16 (sentence_encoder.py:142) whatever their length.  Padding is at most the
        spread of lengths inside one bucket; the pooling kernel never reads it anyway."""
        budget = int(self.params.token_budget)
        enc = self.params.tokenizer(
            text=list(documents),
            add_special_tokens=True,
            padding=False,
            truncation=True,
            max_length=self.params.sequence_max_len,
            return_attention_mask=False,
            return_token_type_ids=False,
            return_tensors=None,
        )
        ids = enc["input_ids"]
        lens = np.asarray([len(x) for x in ids], dtype=np.int64)
        flat = np.concatenate([np.asarray(x, dtype=np.int64) for x in ids]) if len(ids) else np.zeros(0, np.int64)
        offs = np.concatenate([[0], np.cumsum(lens)])
        order = np.argsort(lens, kind="stable")
        pad_id = getattr(self.params.tokenizer, "pad_token_id", None) or 0
        start, n = 0, len(ids)
        while start < n:
            # sorted ascending: the longest sentence of a candidate batch [start, end) is its last one
            end = start + 1
            while end < n and (end + 1 - start) * int(lens[order[end]]) <= budget:
                end += 1
            rows = order[start:end]
            width = int(lens[rows[-1]])
            live = np.arange(width)[None, :] < lens[rows][:, None]           # [b, width] attention mask
            batch = np.full((len(rows), width), pad_id, dtype=np.int64)
            batch[live] = np.concatenate([flat[offs[i]:offs[i + 1]] for i in rows])   # row-major: sentence by sentence
            feats = EmbeddingsFeatures(
                input_ids=torch.from_numpy(batch).to(self.params.device, non_blocking=True),
                attention_mask=torch.from_numpy(live.astype(np.int64)).to(self.params.device, non_blocking=True))
            yield torch.as_tensor(rows, dtype=torch.int64), feats
            start = end

    def _token_embeddings(self, feats: EmbeddingsFeatures) -> torch.Tensor:
        """Encoder forward -> last hidden state.  ``params.encode_dtype = torch.bfloat16`` runs it under
        ``torch.autocast`` (tensor-core matmuls; whatever dtype comes out, K1 reads it as is)."""
        dt = getattr(self.params, "encode_dtype", None)
        if dt in (torch.bfloat16, torch.float16) and torch.device(self.params.device).type == "cuda":
            with torch.autocast("cuda", dtype=dt):
                return self.context_embedder(**feats.to_dict())[0]
        return self.context_embedder(**feats.to_dict())[0]

    def encode_text(self, documents: List[str], output_np: bool = False) -> Union[torch.Tensor, np.ndarray]:
        """Text -> [n, D] fp32 mean-pooled embeddings on ``params.device`` (reference :136-173).
        Rows are written by the pooling kernel straight to their un-sorted positions instead of
        being collected in a Python list and re-stacked (:167-173)."""
        self.to(self.params.device)
        self.eval()
        n = len(documents)
        out = None
        fused = isinstance(self.pooler, AvgPoolingStrategy) and isinstance(self.projection, nn.Identity)
        with torch.no_grad():
            for rows, feats in self._batches(documents):
                if fused:
                    tokens = self._token_embeddings(feats)
                    if out is None:
                        out = torch.empty(n, tokens.shape[-1], dtype=torch.float32, device=tokens.device)
                    from . import ops
                    ops.pool_norm(tokens, feats.attention_mask, normalize=False, out=out,
                                  out_rows=rows.to(tokens.device))
                else:
                    emb = self.embed(feats).detach()
                    if out is None:
                        out = torch.empty(n, emb.shape[-1], dtype=emb.dtype, device=emb.device)
                    out[rows.to(emb.device)] = emb
        if out is None:
            out = torch.empty(0, self.get_sentence_embedding_dimension(), device=self.params.device)
        return out.cpu().numpy() if output_np else out

    def encode_text_normalized(self, documents: List[str], out_dtype: torch.dtype = torch.bfloat16
                               ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Text -> (unit-norm rows [n, D] in ``out_dtype``, float32 inverse norms of the stored rows):
        the form the exact search stores its corpus in.  One fused K1 launch per batch."""
        if not (isinstance(self.pooler, AvgPoolingStrategy) and isinstance(self.projection, nn.Identity)):
            from . import ops
            emb = self.encode_text(documents)
            rows = torch.nn.functional.normalize(emb.float(), dim=-1).to(out_dtype)
            return rows, ops.row_inv_norm(rows)
        self.to(self.params.device)
        self.eval()
        n = len(documents)
        out = inv = None
        with torch.no_grad():
            for rows, feats in self._batches(documents):
                tokens = self._token_embeddings(feats)
                if out is None:
                    out = torch.empty(n, tokens.shape[-1], dtype=out_dtype, device=tokens.device)
                    inv = torch.empty(n, dtype=torch.float32, device=tokens.device)
                self.pooler.pool_normalized(tokens, feats, out=out, out_rows=rows.to(tokens.device), out_inv_norm=inv)
        if out is None:
            dev = self.params.device
            out = torch.empty(0, self.get_sentence_embedding_dimension(), dtype=out_dtype, device=dev)
            inv = torch.empty(0, dtype=torch.float32, device=dev)
        return out, inv

    def encode_text_into(self, documents: List[str], store) -> List[int]:
        """Text -> rows appended to an :class:`~text_similarity_b200.store.EmbeddingStore`, in document
        order: every length-sorted batch is pooled, normalised, cast and written by ONE K1 launch
        straight into the store's tail at its un-sorted positions (replaces the
        per-row ``extend`` + ``stack`` of reference sentence_encoder.py:167-173).  Returns the labels."""
        if not (isinstance(self.pooler, AvgPoolingStrategy) and isinstance(self.projection, nn.Identity)):
            rows, inv = self.encode_text_normalized(documents, store.dtype)
            return store.add(rows, inv).tolist()
        self.to(self.params.device)
        self.eval()
        n = len(documents)
        if n == 0:
            return []
        from . import ops
        base = len(store)
        labels = store._take_ids(n, None)
        store._reserve(n)
        with torch.no_grad():
            for rows, feats in self._batches(documents):
                tokens = self._token_embeddings(feats)
                ops.pool_norm(tokens, feats.attention_mask, normalize=True, out=store.rows,
                              out_rows=rows.to(tokens.device) + base, out_inv_norm=store.inv_norm)
        store._register(labels)
        return labels.tolist()

    def get_sentence_embedding_dimension(self):
        return self.context_embedder.config.hidden_size  # :175-176

    @classmethod
    def from_pretrained(cls, path, pooler=None, merge_strategy=None, loss=None, params=None, parallel_mode=True):
        """Reference :187-217 (needs a local checkpoint directory: no network here)."""
        import transformers
        cfg_path = os.path.join(path, "model_config.bin")
        if os.path.exists(cfg_path):
            params = torch.load(cfg_path, weights_only=False)
        assert params is not None, "Parameters not found, need to pass model parameters for the model to work"
        embedder_config = transformers.AutoConfig.from_pretrained(path)
        context_embedder = transformers.AutoModel.from_pretrained(path, config=embedder_config)
        return cls(pooler=pooler or AvgPoolingStrategy(params), merge_strategy=merge_strategy, loss=loss,
                   params=params, context_embedder=context_embedder, parallel_mode=parallel_mode)

    load_pretrained = from_pretrained  # name used by the reference's eval scripts
