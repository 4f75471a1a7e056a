"""Offline helpers for tests, smoke and bench: a deterministic tokenizer with the HuggingFace call
signature the reference uses (sentence_encoder.py:144-153) and the MiniLM-L6-shaped random-init
encoder of BASELINE config 1 (there is no network for checkpoints or vocabularies)."""
from __future__ import annotations

import zlib
from typing import List, Optional

import torch


class SyntheticTokenizer:
    """Whitespace words hashed into a BERT-sized vocabulary; [CLS]=101, [SEP]=102, [PAD]=0."""

    def __init__(self, vocab_size: int = 30522):
        self.vocab_size = vocab_size

    def _ids(self, sentence: str, max_length: Optional[int], special: bool) -> List[int]:
        ids = [1000 + zlib.crc32(w.encode()) % (self.vocab_size - 1000) for w in sentence.split()]
        if special:
            if max_length is not None:
                ids = ids[:max(0, max_length - 2)]
            return [101] + ids + [102]
        return ids[:max_length] if max_length is not None else ids

    def __call__(self, text, add_special_tokens=True, padding="longest", truncation=True, max_length=None,
                 return_attention_mask=True, return_token_type_ids=False, return_tensors="pt"):
        if isinstance(text, str):
            text = [text]
        rows = [self._ids(t, max_length if truncation else None, add_special_tokens) for t in text]
        if not padding and return_tensors is None:        # HF convention: ragged python lists
            out = {"input_ids": rows}
            if return_attention_mask:
                out["attention_mask"] = [[1] * len(r) for r in rows]
            return out
        width = max((len(r) for r in rows), default=0)
        if padding == "max_length" and max_length is not None:
            width = max_length
        ids = torch.zeros(len(rows), width, dtype=torch.int64)
        mask = torch.zeros(len(rows), width, dtype=torch.int64)
        for i, r in enumerate(rows):
            ids[i, :len(r)] = torch.tensor(r, dtype=torch.int64)
            mask[i, :len(r)] = 1
        out = {"input_ids": ids}
        if return_attention_mask:
            out["attention_mask"] = mask
        if return_token_type_ids:
            out["token_type_ids"] = torch.zeros_like(ids)
        return out

    def save_pretrained(self, path):
        pass


def synthetic_sentences(n: int, seed: int, min_words: int = 3, max_words: int = 40) -> List[str]:
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(min_words, max_words + 1, (n,), generator=g).tolist()
    words = torch.randint(0, 5000, (sum(lens),), generator=g).tolist()
    out, p = [], 0
    for ln in lens:
        out.append(" ".join(f"w{w}" for w in words[p:p + ln]))
        p += ln
    return out


def minilm_l6_encoder(seed: int = 0, layers: int = 6):
    """transformers.BertModel with the MiniLM-L6 shape (hidden 384, 12 heads, FFN 1536), random init."""
    from transformers import BertConfig, BertModel
    torch.manual_seed(seed)
    cfg = BertConfig(vocab_size=30522, hidden_size=384, num_hidden_layers=layers, num_attention_heads=12,
                     intermediate_size=1536, max_position_embeddings=512)
    return BertModel(cfg).eval()
