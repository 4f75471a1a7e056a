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


class AverageMeter:
    """Running average holder (reference metrics.py:125-161): the base of the meters."""

    def __init__(self, name, return_predictions=False):
        self.name = name
        self.return_predictions = return_predictions
        self.reset()

    def __str__(self):
        return f"average {self.name}: {self.avg}"

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0
        if self.return_predictions:
            self.all_predictions, self.all_labels = [], []

    def update(self, val, n=1, **kwargs):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class RetrievalAccuracyMeter(AverageMeter):
    """Bitext-retrieval accuracy in both directions (reference metrics.py:449-507): for every source
    sentence the most similar target (and vice versa) must be its own translation.  The reference
    builds the dense [M, M] cosine matrix, copies it to the host and runs M NumPy argmaxes per
    direction (:470-498); here each direction is ONE fused search with k = 5 (top-1 for the accuracy,
    top-5 for the wrong-match report of :487-490) and only [M, 5] results leave the GPU.  Ties resolve
    to the lower index, like ``np.argmax``."""

    def __init__(self, print_wrong_matches=True, **kwargs):
        super().__init__(name="accuracy", **kwargs)
        self.print_wrong_matches = print_wrong_matches
        self.src2tgt = 0
        self.tgt2src = 0
        self.lines = []
        self.precision = 0
        self.recall = 0
        self.f1 = 0

    def __str__(self):
        accuracy = "accuracy [src2tgt: {:.2f} tgt2src: {:.2f}]".format(self.src2tgt, self.tgt2src)
        f1 = "precision: {:.2f} recall: {:.2f} f1: {:.2f}".format(self.precision, self.recall, self.f1)
        return "\n\n".join(self.lines + [accuracy, f1])

    def update(self, src_embeddings, tgt_embeddings, source_sentences=None, target_sentences=None, **kwargs):
        src, tgt = _as_2d(src_embeddings), _as_2d(tgt_embeddings)
        if src.shape[0] != tgt.shape[0]:
            raise ValueError("source and target sides must be aligned sentence by sentence")
        m = src.shape[0]
        k = min(5, m)
        s_fwd, i_fwd = cos_sim_topk(src, tgt, k)
        _, i_bwd = cos_sim_topk(tgt, src, 1)
        own = torch.arange(m, device=i_fwd.device)
        hit_fwd = i_fwd[:, 0] == own
        self.src2tgt = float(hit_fwd.float().mean()) if m else 0.0
        self.tgt2src = float((i_bwd[:, 0] == own).float().mean()) if m else 0.0
        if self.print_wrong_matches and source_sentences is not None and target_sentences is not None:
            s_host, i_host = s_fwd.cpu(), i_fwd.cpu()
            for i in (~hit_fwd).nonzero().flatten().tolist():
                j = int(i_host[i, 0])
                self.lines.append(f"i: {i} j: {j}, INCORRECT\nsrc: {source_sentences[i]}\ntgt: {target_sentences[j]}\n"
                                  f"maximum score: {float(s_host[i, 0])}, top-{k}: "
                                  + ", ".join(f"{int(c)} ({float(v):.4f})" for c, v in zip(i_host[i], s_host[i])))
        self.avg = (self.src2tgt + self.tgt2src) / 2
        self.val = self.avg
