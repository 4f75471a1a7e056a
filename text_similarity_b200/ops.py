"""Tensor-level wrappers over the C ABI (libtsim.so).

torch is used for device memory and streams only; every kernel here is hand-written CUDA for
sm_100a reached through ctypes.  All functions require CUDA tensors and raise otherwise.
"""
from __future__ import annotations

import atexit
from collections import OrderedDict
from typing import Optional, Tuple

import torch

from . import _lib

_DT = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16,
       torch.float8_e4m3fn: _lib.E4M3}
_MASK_DT = {torch.int64: _lib.I64, torch.int32: _lib.I32, torch.uint8: _lib.U8, torch.bool: _lib.U8,
            torch.float32: _lib.F32}
_MODES = {"auto": _lib.MODE_AUTO, "exact": _lib.MODE_EXACT, "tensor": _lib.MODE_TENSOR}

_workspaces = {}

# Plan handles (tsim_plan_create): launch plan + cached TMA descriptors per (device, shape, k, dtypes, mode).
# Python owns them: least recently used ones are destroyed past _PLAN_CAP, the rest at interpreter exit.
_plans: "OrderedDict[tuple, int]" = OrderedDict()
_PLAN_CAP = 256


_SPLIT = -2      # shadow_dt value of _plan() for a split (hi + lo) bf16 shadow


def _plan(dev: torch.device, Q: int, N: int, D: int, k: int, q_dt: int, c_dt: int, mode: int, shadow_dt: int) -> int:
    key = (dev.index, Q, N, D, k, q_dt, c_dt, mode, shadow_dt)
    h = _plans.get(key)
    if h is not None:
        _plans.move_to_end(key)
        return h
    lib = _lib.load()
    with torch.cuda.device(dev):
        h = (lib.tsim_plan_create_split_shadow(Q, N, D, k, q_dt, c_dt) if shadow_dt == _SPLIT
             else lib.tsim_plan_create(Q, N, D, k, q_dt, c_dt, mode, shadow_dt))
    if not h:
        _lib.check(_lib.ERR_UNSUPPORTED if mode == _lib.MODE_TENSOR else _lib.ERR_INVALID_ARG, "tsim_plan_create")
    _plans[key] = h
    while len(_plans) > _PLAN_CAP:
        _, old = _plans.popitem(last=False)
        lib.tsim_plan_destroy(old)
    return h


def search_workspace_bytes(Q: int, N: int, D: int, k: int, q_dtype: torch.dtype, c_dtype: torch.dtype,
                           mode: str = "auto", shadow: bool = False, device: Optional[torch.device] = None,
                           split: bool = False) -> int:
    """Bytes of workspace ``search_topk`` needs for this call shape (for callers that own their workspace,
    e.g. one per captured CUDA graph).  ``shadow`` / ``split``: the call passes a rounded / split bf16 shadow."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    h = _plan(dev, Q, N, D, int(k), _DT[q_dtype], _DT[c_dtype], _lib.MODE_AUTO if (shadow or split) else _MODES[mode],
              _SPLIT if split else _lib.BF16 if shadow else -1)
    return int(_lib.load().tsim_plan_workspace_bytes(h))


@atexit.register
def _destroy_plans() -> None:
    if _lib._lib is None:
        return
    while _plans:
        _, h = _plans.popitem()
        _lib._lib.tsim_plan_destroy(h)


def _require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("text_similarity_b200 ops need CUDA tensors: there is no CPU fallback")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _workspace(dev: torch.device, nbytes: int, tag: str) -> torch.Tensor:
    """A per-(device, stream, purpose) scratch buffer that only ever grows."""
    key = (dev.index, _stream(dev), tag)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        _workspaces[key] = ws
    return ws


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise ValueError(f"unsupported dtype {t.dtype}") from None


def pool_norm(token_embeddings: torch.Tensor, attention_mask: torch.Tensor, *,
              out_dtype: torch.dtype = torch.float32, normalize: bool = True,
              out: Optional[torch.Tensor] = None, out_rows: Optional[torch.Tensor] = None,
              out_inv_norm: Optional[torch.Tensor] = None
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """K1 -- fused masked mean-pool (+ L2 normalise) + cast.

    token_embeddings [B, L, D] fp32/fp16/bf16, attention_mask [B, L] int64/int32/bool/uint8/fp32.
    Returns (rows, inv_norm): rows [B, D] in ``out_dtype`` (or written into ``out`` at rows
    ``out_rows``), inv_norm float32 = 1 / max(||stored row||, 1e-8).
    Mean pooling follows reference src/modules/modules.py:158-171.
    """
    lib = _lib.load()
    dev = _require_cuda(token_embeddings, attention_mask, out, out_rows, out_inv_norm)
    if token_embeddings.dim() != 3:
        raise ValueError("token_embeddings must be [batch, seq_len, embed_size]")  # modules.py:159
    if torch.is_grad_enabled() and token_embeddings.requires_grad:
        # the kernel has no backward: refusing beats returning a detached tensor that silently drops gradients
        raise RuntimeError("pool_norm (CUDA kernel K1) is inference-only: call it under torch.no_grad() or detach "
                           "the token embeddings; training through the pooler is outside this build")
    B, L, D = token_embeddings.shape
    if attention_mask.shape != (B, L):
        raise ValueError(f"attention_mask shape {tuple(attention_mask.shape)} != {(B, L)}")
    tok = token_embeddings if token_embeddings.stride(2) == 1 else token_embeddings.contiguous()
    mask = attention_mask if attention_mask.stride(1) == 1 else attention_mask.contiguous()
    if mask.dtype not in _MASK_DT:
        raise ValueError(f"unsupported mask dtype {mask.dtype}")
    if out is None:
        if out_rows is not None:
            raise ValueError("out_rows needs out")
        out = torch.empty(B, D, dtype=out_dtype, device=dev)
    else:
        if out.dim() != 2 or out.shape[1] != D or out.stride(1) != 1:
            raise ValueError("out must be [rows, D] with contiguous rows")
        out_dtype = out.dtype
    n_out = out.shape[0]
    if out_inv_norm is None:
        out_inv_norm = torch.empty(n_out, dtype=torch.float32, device=dev)
    if out_rows is not None:
        out_rows = out_rows.to(device=dev, dtype=torch.int64).contiguous()
        if out_rows.numel() != B:
            raise ValueError("out_rows must have one entry per batch row")
    elif n_out < B:
        raise ValueError("out has fewer rows than the batch")
    if B == 0:
        return out, out_inv_norm
    nbytes = lib.tsim_pool_workspace_bytes(B, L, D)
    ws = _workspace(dev, nbytes, "pool")
    with torch.cuda.device(dev):
        rc = lib.tsim_pool_norm(tok.data_ptr(), _dt(tok), mask.data_ptr(), _MASK_DT[mask.dtype],
                                B, L, D, tok.stride(0), tok.stride(1), mask.stride(0),
                                out.data_ptr(), _dt(out), out.stride(0), _ptr(out_rows),
                                out_inv_norm.data_ptr(), int(bool(normalize)),
                                ws.data_ptr(), ws.numel(), _stream(dev))
    _lib.check(rc, "tsim_pool_norm")
    return out, out_inv_norm


def row_inv_norm(x: torch.Tensor) -> torch.Tensor:
    """1 / max(||x[i]||, 1e-8) per row (float32) of a stored [N, D] matrix."""
    lib = _lib.load()
    dev = _require_cuda(x)
    if x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("x must be [N, D] with contiguous rows")
    out = torch.empty(x.shape[0], dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.tsim_row_inv_norm(x.data_ptr(), _dt(x), x.shape[0], x.shape[1], x.stride(0),
                                   out.data_ptr(), _stream(dev))
    _lib.check(rc, "tsim_row_inv_norm")
    return out


def _split_bf16(rows: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """x = hi + lo + r with hi = bf16(x), lo = bf16(x - hi): |r| <= 2^-18 |x| (two bf16 parts carry 16 mantissa bits)."""
    x = rows.float()
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


def _split_seg(D: int) -> int:
    """Segment width of a split shadow: D rounded up to whole 64-element k-blocks (tsim.h, split_shadow_seg)."""
    return (D + 63) // 64 * 64


def _split_query_shadow(queries: torch.Tensor) -> torch.Tensor:
    """[Q, 3 Dp] query shadow of a split pass: per 64-element block j the triple (hi_j, lo_j, hi_j), which the kernel
    multiplies with the corpus shadow's (hi_j, hi_j, lo_j)."""
    Q, D = queries.shape
    Dp = _split_seg(D)
    qh, ql = _split_bf16(queries)
    if Dp != D:
        qh = torch.nn.functional.pad(qh, (0, Dp - D))
        ql = torch.nn.functional.pad(ql, (0, Dp - D))
    qh, ql = qh.view(Q, Dp // 64, 64), ql.view(Q, Dp // 64, 64)
    return torch.stack([qh, ql, qh], dim=2).reshape(Q, 3 * Dp).contiguous()


def make_shadow(rows: torch.Tensor, split: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """bf16 shadow of fp32 / fp16 rows plus the inverse norms that go with it: what ``search_topk`` needs
    (``corpus_shadow=``, ``shadow_inv_norm=``) to search such rows at tensor-core speed, still exactly.

    Default: the rows rounded to bf16, [N, D] (+50 % memory on fp32 rows) -- good for k <= 24 (the proof has to span
    the 6e-3 the rounding can move a cosine, so 112 candidates are re-scored per query).
    ``split=True``: [N, 2 Dp] rows [hi | lo], Dp = D rounded up to 64 (+100 % on fp32 rows): one bf16 pass of three
    segments computes qh.ch + ql.ch + qh.cl, q.c to ~1e-5 -- the ordinary candidate lists then prove k up to 100
    (1M x 768 fp32, Q = 1024, k = 100: ~4 ms instead of the 70 ms float64 scan)."""
    _require_cuda(rows)
    if not split:
        shadow = rows.to(torch.bfloat16).contiguous()
        return shadow, row_inv_norm(shadow)
    N, D = rows.shape
    Dp = _split_seg(D)
    shadow = (torch.zeros if Dp != D else torch.empty)((N, 2 * Dp), dtype=torch.bfloat16, device=rows.device)
    for s in range(0, N, 1 << 20):          # in slabs: the fp32 temporaries of a 10M-row shard would not fit beside it
        hi, lo = _split_bf16(rows[s:s + (1 << 20)])
        shadow[s:s + (1 << 20), :D] = hi
        shadow[s:s + (1 << 20), Dp:Dp + D] = lo
    return shadow, row_inv_norm(rows.contiguous())


def search_topk(queries: torch.Tensor, corpus: torch.Tensor, k: int, *,
                corpus_inv_norm: Optional[torch.Tensor] = None, idx_base: int = 0,
                exclude_self_base: int = -1, mode: str = "auto",
                return_score64: bool = False, return_flags: bool = False,
                out_scores: Optional[torch.Tensor] = None, out_idx: Optional[torch.Tensor] = None,
                out_score64: Optional[torch.Tensor] = None,
                corpus_shadow: Optional[torch.Tensor] = None, shadow_inv_norm: Optional[torch.Tensor] = None,
                workspace: Optional[torch.Tensor] = None):
    """K2 + K3 -- exact cosine top-k of every query row against every corpus row.

    Returns (scores float32 [Q, k], idx int64 [Q, k]) best first, ties by lower index, idx -1 /
    score -inf past the last available row; optionally the float64 scores and the per-query
    stage flags (0: first tensor pass, 2: wide retry pass, 1: float64 scan).  Replaces the loop at reference src/pipeline/search_pipeline.py:73-79.
    ``workspace``: a caller-owned uint8 scratch tensor of at least ``search_workspace_bytes(...)`` bytes (a captured
    CUDA graph must own the buffer it replays into); default: a growable per-(device, stream) buffer.
    """
    lib = _lib.load()
    dev = _require_cuda(queries, corpus, corpus_inv_norm)
    if queries.dim() != 2 or corpus.dim() != 2 or queries.shape[1] != corpus.shape[1]:
        raise ValueError(f"queries {tuple(queries.shape)} and corpus {tuple(corpus.shape)} must be [*, D]")
    if queries.stride(1) != 1:
        queries = queries.contiguous()
    if corpus.stride(1) != 1:
        corpus = corpus.contiguous()
    Q, D = queries.shape
    N = corpus.shape[0]
    k = int(k)
    if mode not in _MODES:
        raise ValueError(f"mode must be one of {sorted(_MODES)}")
    if corpus_inv_norm is not None:
        if corpus_inv_norm.dtype != torch.float32 or corpus_inv_norm.numel() != N:
            raise ValueError("corpus_inv_norm must be float32 [N]")
        corpus_inv_norm = corpus_inv_norm.contiguous()
    def _out(t, dtype):
        if t is None:
            return torch.empty(Q, k, dtype=dtype, device=dev)
        if t.shape != (Q, k) or t.dtype != dtype or not t.is_contiguous() or t.device != dev:
            raise ValueError(f"output tensor must be contiguous {dtype} [{Q}, {k}] on {dev}")
        return t

    scores = _out(out_scores, torch.float32)
    idx = _out(out_idx, torch.int64)
    s64 = _out(out_score64, torch.float64) if (return_score64 or out_score64 is not None) else None
    flags = torch.empty(Q, dtype=torch.int32, device=dev) if return_flags else None
    shadow = corpus_shadow is not None and mode == "auto" and N > 0
    q_shadow = None
    split = False
    if shadow:
        # fp32 / fp16 rows: candidates from the bf16 shadows on the tensor cores, float64 re-score on the originals
        _require_cuda(corpus_shadow, shadow_inv_norm)
        split = corpus_shadow.dim() == 2 and corpus_shadow.shape == (N, 2 * _split_seg(D))
        if ((corpus_shadow.shape != corpus.shape and not split) or corpus_shadow.dtype != torch.bfloat16
                or corpus_shadow.stride(1) != 1):
            raise ValueError("corpus_shadow must be a bfloat16 [N, D] (rounded) or [N, 2 * ceil64(D)] (split) tensor with contiguous rows")
        if split:
            q_shadow = _split_query_shadow(queries)
        else:
            q_shadow = queries.to(torch.bfloat16).contiguous()
    plan = _plan(dev, Q, N, D, k, _dt(queries), _dt(corpus), _lib.MODE_AUTO if shadow else _MODES[mode],
                 _SPLIT if split else _lib.BF16 if shadow else -1)
    nbytes = lib.tsim_plan_workspace_bytes(plan)
    if workspace is None:
        ws = _workspace(dev, nbytes, "search")
    else:
        _require_cuda(workspace)
        if workspace.dtype != torch.uint8 or not workspace.is_contiguous() or workspace.numel() < nbytes:
            raise ValueError(f"workspace must be a contiguous uint8 tensor of >= {nbytes} bytes")
        ws = workspace
    inv = shadow_inv_norm if shadow else corpus_inv_norm
    with torch.cuda.device(dev):
        rc = lib.tsim_plan_search(plan, queries.data_ptr(), queries.stride(0),
                                  corpus.data_ptr() if N else None, corpus.stride(0) if N else D, _ptr(inv),
                                  _ptr(q_shadow), q_shadow.stride(0) if shadow else 0,
                                  _ptr(corpus_shadow) if shadow else None, corpus_shadow.stride(0) if shadow else 0,
                                  int(idx_base), int(exclude_self_base),
                                  scores.data_ptr(), _ptr(s64), idx.data_ptr(), _ptr(flags),
                                  ws.data_ptr(), ws.numel(), _stream(dev))
    _lib.check(rc, "tsim_plan_search")
    res = [scores, idx]
    if return_score64 or out_score64 is not None:
        res.append(s64)
    if return_flags:
        res.append(flags)
    return tuple(res)


def merge_topk(scores64: torch.Tensor, idx: torch.Tensor, k_out: int, n_lists: int = 1
               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """K3 second pass -- merge candidate lists [Q, n_lists * k_in] (float64 scores, int64 rows,
    -1 = padding) into the top ``k_out`` by (score desc, row asc).
    Returns (scores float32, scores float64, idx int64)."""
    lib = _lib.load()
    dev = _require_cuda(scores64, idx)
    if scores64.shape != idx.shape or scores64.dim() != 2:
        raise ValueError("scores64 and idx must both be [Q, n_lists * k_in]")
    if scores64.dtype != torch.float64 or idx.dtype != torch.int64:
        raise ValueError("merge_topk takes float64 scores and int64 indices")
    scores64 = scores64.contiguous()
    idx = idx.contiguous()
    Q, total = scores64.shape
    if total % n_lists:
        raise ValueError("row length must be n_lists * k_in")
    out_s = torch.empty(Q, k_out, dtype=torch.float32, device=dev)
    out_s64 = torch.empty(Q, k_out, dtype=torch.float64, device=dev)
    out_i = torch.empty(Q, k_out, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.tsim_merge_topk(scores64.data_ptr(), idx.data_ptr(), Q, n_lists, total // n_lists,
                                 int(k_out), out_s.data_ptr(), out_s64.data_ptr(), out_i.data_ptr(),
                                 _stream(dev))
    _lib.check(rc, "tsim_merge_topk")
    return out_s, out_s64, out_i


def merge_gathered(recv: torch.Tensor, Q: int, k: int, world: int
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """K3 second pass straight on the receive buffer of the all-gather of per-shard results: ``recv`` is int64
    [world * 2, Q, k] -- per rank a [Q, k] block of float64 score bits followed by a [Q, k] block of global rows --
    read in place (tsim_merge_topk_strided), no permute / reshape copy.  Returns (scores float32, scores float64,
    rows int64), each [Q, k]."""
    lib = _lib.load()
    dev = _require_cuda(recv)
    if recv.dtype != torch.int64 or not recv.is_contiguous() or recv.numel() != world * 2 * Q * k:
        raise ValueError(f"recv must be a contiguous int64 tensor of {world * 2 * Q * k} elements")
    out_s = torch.empty(Q, k, dtype=torch.float32, device=dev)
    out_s64 = torch.empty(Q, k, dtype=torch.float64, device=dev)
    out_i = torch.empty(Q, k, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.tsim_merge_topk_strided(recv.data_ptr(), recv.data_ptr() + Q * k * 8, Q, world, k, 2 * Q * k, k, k,
                                         out_s.data_ptr(), out_s64.data_ptr(), out_i.data_ptr(), _stream(dev))
    _lib.check(rc, "tsim_merge_topk_strided")
    return out_s, out_s64, out_i
