"""Pooling modules with the reference's interface (src/modules/modules.py:34-55,154-171), the
arithmetic done by the fused CUDA kernel K1 (csrc/pool_norm.cu) through the C ABI."""
from __future__ import annotations

import torch
from torch import nn

from . import ops


class LearningStrategy(nn.Module):
    """Base class of the strategy modules (reference modules.py:34-41)."""

    def forward(self, *args, **kwargs):
        raise NotImplementedError()


class PoolingStrategy(LearningStrategy):
    """Base class of poolers (reference modules.py:44-55): keeps ``params``."""

    def __init__(self, params=None, *args, **kwargs):
        super().__init__()
        self.params = params

    def forward(self, embeddings: torch.Tensor, features=None):
        raise NotImplementedError()


def _mask_of(features) -> torch.Tensor:
    if isinstance(features, torch.Tensor):
        return features
    if isinstance(features, dict):
        return features["attention_mask"]
    return features.to_dict()["attention_mask"]  # modules.py:160


class AvgPoolingStrategy(PoolingStrategy):
    """Masked mean over the token axis: ``sum_l e[b,l,:] m[b,l] / max(sum_l m[b,l], 1e-9)``.

    Drop-in for reference modules.py:154-171 (parameter-free, empty state_dict).  ``forward``
    returns the un-normalised fp32 mean exactly like the reference; ``pool_normalized`` is the
    fused form the search pipeline uses (mean -> L2 normalise -> bf16/fp8 cast + inverse norms in
    one pass over the token tensor).  Inference-only: the kernel has no backward, and pooling a tensor
    that requires grad outside ``torch.no_grad()`` raises instead of silently dropping the gradient (the
    reference's pooler is also used in training, modules.py:154-171; training is outside this build).
    """

    def forward(self, embeddings: torch.Tensor, features):
        assert len(embeddings.shape) == 3  # batch, seq_len, embed_size (modules.py:159)
        out, _ = ops.pool_norm(embeddings, _mask_of(features), out_dtype=torch.float32, normalize=False)
        return out

    def pool_normalized(self, embeddings: torch.Tensor, features, out_dtype=torch.bfloat16,
                        out=None, out_rows=None, out_inv_norm=None):
        assert len(embeddings.shape) == 3
        return ops.pool_norm(embeddings, _mask_of(features), out_dtype=out_dtype, normalize=True,
                             out=out, out_rows=out_rows, out_inv_norm=out_inv_norm)
