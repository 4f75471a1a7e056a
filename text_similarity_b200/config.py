"""Configuration carriers read on the search path.

Same class and field names as reference src/configurations/config.py:7-44, so scripts that build
``Configuration(model_parameters=..., model=..., save_path=..., tokenizer=..., device=...)`` keep
working.  Repaired: ``SearchConfiguration.ef / ef_construction / M`` are plain ints (the reference's
trailing commas make them 1-tuples, config.py:42-44).  Extra optional search knobs
(corpus storage dtype, top-k mode) are appended with defaults.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional, Tuple

import torch


@dataclass
class ModelParameters:
    model_name: str
    hidden_size: int = 768
    num_classes: int = 2
    use_pretrained_embeddings: bool = False
    freeze_weights: bool = True
    context_layers: Tuple[int, ...] = (-1,)
    output_attention = False  # un-annotated class attribute, as in the reference (config.py:14)


@dataclass
class Configuration:
    model_parameters: Optional[ModelParameters]
    model: str
    save_path: str
    tokenizer: Any = None
    sequence_max_len: int = 256
    dropout_prob: float = 0.1
    lr: float = 2e-5
    batch_size: int = 16
    epochs: int = 1
    device: torch.device = torch.device("cuda")
    warmup_steps: int = 0
    fp16: bool = True
    model_path: Optional[str] = None


@dataclass
class SearchConfiguration(Configuration):
    ef: int = 50
    ef_construction: int = 400
    M: int = 64
    # --- additions for the exact B200 engine (not in the reference) ---
    corpus_dtype: torch.dtype = torch.bfloat16   # storage type of the unit-norm corpus matrix
    search_mode: str = "auto"                    # "auto" | "tensor" | "exact" (include/tsim.h)
    # encode side.  Defaults reproduce the reference's loop: fp32 encoder, fixed batches of
    # `batch_size` sentences (sentence_encoder.py:142-167).
    encode_dtype: Optional[torch.dtype] = None   # torch.bfloat16: run the encoder under torch.autocast(bf16)
    token_budget: Optional[int] = None           # > 0: length-bucketed batches of at most this many (padded) tokens
