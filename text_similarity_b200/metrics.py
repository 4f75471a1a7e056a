"""``cos_sim`` with the reference's signature (src/utils/metrics.py:81-101) plus the fused
top-k form that never materialises the [M, N] matrix."""
from __future__ import annotations

import torch

from . import ops


def _as_2d(x) -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        x = torch.tensor(x)  # metrics.py:87-91
    if x.dim() == 1:
        x = x.unsqueeze(0)  # metrics.py:93-97
    return x


def cos_sim(a, b) -> torch.Tensor:
    """res[i][j] = cosine(a[i], b[j]) as a dense fp32 [M, N] matrix (reference metrics.py:81-101).

    The row norms come from the K1 norm kernel; the dense product itself is a plain library GEMM
    (the one place the build calls cuBLAS: this entry point exists for drop-in compatibility, the
    search path uses ``cos_sim_topk`` instead).  Zero rows give 0, not the reference's NaN (A11).
    """
    a, b = _as_2d(a), _as_2d(b)
    if not (a.is_cuda and b.is_cuda):
        raise RuntimeError("cos_sim needs CUDA tensors: there is no CPU fallback")
    a32, b32 = a.float().contiguous(), b.float().contiguous()
    an = a32 * ops.row_inv_norm(a32)[:, None]
    bn = b32 * ops.row_inv_norm(b32)[:, None]
    return torch.mm(an, bn.transpose(0, 1))


def cos_sim_topk(a, b, k: int, exclude_self: bool = False, mode: str = "auto"):
    """Exact top-k of every row of ``cos_sim(a, b)`` without building the matrix: (scores [M, k],
    idx [M, k]) best first, ties by lower index.  ``exclude_self`` skips b[i] for a[i] (all-pairs
    mining over one matrix).  This is what RetrievalAccuracyMeter's row-wise argmax
    (metrics.py:476-498) needs with k = 1."""
    a, b = _as_2d(a), _as_2d(b)
    return ops.search_topk(a, b, k, exclude_self_base=0 if exclude_self else -1, mode=mode)
