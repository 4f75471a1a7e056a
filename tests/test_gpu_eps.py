"""Pins the error bound the tensor path's completeness proof rests on (csrc/tsim_common.cuh: approx_eps,
shadow_eps; used at select_merge.cu "completeness proof").

Every "exact" answer of the tcgen05 path assumes |tensor-core dot - exact dot| <= eps * ||q|| * ||c||.  These tests
read the approximate scores themselves -- the packed candidate keys of ONE tensor pass, through the test hook
tsim_debug_tensor_pass, 112 rows at a time so that no row is dropped -- and compare them with float64 on the same
stored values, over widths 64 ... 16384, both element types and row families chosen to stress fp32 accumulation:
unit-norm Gaussian rows, rows with a 10^3 dynamic range, cancellation-heavy rows ((+x, -x) halves against (x, x)
queries: large partial sums, tiny result) and near-duplicates of the query (all products positive: one-sided
truncation adds up).  The bar: measured max error <= 1/4 of the bound the library uses.
"""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

SLICE = 112   # rows per call: one tile, one 112-entry list -> every row's score comes back


def _unpack(keys: torch.Tensor):
    """packed u64 keys (as int64) -> (float32 score, int64 row, bool valid)"""
    hi = (keys >> 32) & 0xFFFFFFFF
    lo = keys & 0xFFFFFFFF
    bits = torch.where((hi & 0x80000000) != 0, hi & 0x7FFFFFFF, (~hi) & 0xFFFFFFFF)
    score = (bits - ((bits >> 31) << 32)).to(torch.int32).view(torch.float32)   # two's complement wrap, then bit cast
    return score, 0xFFFFFFFF - lo, keys != 0


def tensor_pass_scores(q: torch.Tensor, corpus: torch.Tensor) -> torch.Tensor:
    """[Q, N] float32: what the tcgen05 pass computes for every (query, row): dot * (1 / ||row||)."""
    from text_similarity_b200 import _lib, ops
    lib = _lib.load()
    dev = q.device
    Q, D = q.shape
    N = corpus.shape[0]
    dt = ops._dt(q)
    qpad = torch.zeros(128, D, dtype=q.dtype, device=dev)
    qpad[:Q] = q
    inv = ops.row_inv_norm(corpus)
    out = torch.full((Q, N), float("nan"), dtype=torch.float32, device=dev)
    keys = torch.zeros(Q, 1, SLICE, dtype=torch.int64, device=dev)
    thr = torch.zeros(Q, dtype=torch.int32, device=dev)
    lists = ctypes.c_int64(0)
    for s in range(0, N, SLICE):
        n = min(SLICE, N - s)
        piece, pinv = corpus[s:s + n], inv[s:s + n]
        keys.zero_()
        rc = lib.tsim_debug_tensor_pass(qpad.data_ptr(), qpad.stride(0), piece.data_ptr(), piece.stride(0), dt,
                                        pinv.data_ptr(), Q, n, D, keys.data_ptr(), ctypes.byref(lists),
                                        thr.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "tsim_debug_tensor_pass")
        assert lists.value == 1
        score, row, valid = _unpack(keys[:, 0, :])
        assert int(valid.sum()) == Q * n, "a 112-row slice must come back whole"
        qi = torch.arange(Q, device=dev)[:, None].expand(Q, SLICE)
        out[qi[valid], s + row[valid]] = score[valid]
    assert not torch.isnan(out).any()
    return out


def _families(Q, N, D, g, dev):
    """name -> (queries fp32 [Q, D], rows fp32 [N, D])"""
    fam = {}
    x = torch.randn(N, D, generator=g, device=dev)
    y = torch.randn(Q, D, generator=g, device=dev)
    fam["unit_norm"] = (y / y.norm(dim=-1, keepdim=True), x / x.norm(dim=-1, keepdim=True))
    scale_x = 10 ** (3 * torch.rand(N, D, generator=g, device=dev) - 1.5)
    scale_y = 10 ** (3 * torch.rand(Q, D, generator=g, device=dev) - 1.5)
    fam["dynamic_range_1e3"] = (y * scale_y, x * scale_x)
    half = D // 2
    base = torch.randn(N, half, generator=g, device=dev)
    rows = torch.cat([base, -base], dim=1)                       # (+x, -x)
    qs = base[torch.arange(Q, device=dev) % N] + 0.05 * torch.randn(Q, half, generator=g, device=dev)
    fam["cancellation"] = (torch.cat([qs, qs], dim=1), rows)     # (x', x'): partial sums ~ |x|^2, result ~ noise
    dup = y[torch.arange(N, device=dev) % Q] * (1 + 0.02 * torch.randn(N, D, generator=g, device=dev))
    fam["near_duplicates"] = (y, dup)                            # cosine ~ 0.9998, every product positive on the diagonal
    pos = torch.rand(N, D, generator=g, device=dev) + 0.5
    fam["all_positive"] = (torch.rand(Q, D, generator=g, device=dev) + 0.5, pos)
    return fam


def _cast(t: torch.Tensor, dtype):
    if dtype == torch.float8_e4m3fn:
        t = t / t.abs().amax(dim=-1, keepdim=True) * 200.0       # spread over e4m3's range (max 448)
    return t.to(dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float8_e4m3fn])
@pytest.mark.parametrize("D", [64, 768, 4096, 16384])
def test_tensor_core_dot_error_is_within_a_quarter_of_eps(D, dtype):
    from text_similarity_b200 import _lib, ops
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1000 + D)
    Q, N = 64, 4 * SLICE
    eps = _lib.load().tsim_debug_eps(D, ops._DT[dtype], 0)
    assert eps >= 4.99e-5
    worst = {}
    for name, (qf, cf) in _families(Q, N, D, g, dev).items():
        q, c = _cast(qf, dtype), _cast(cf, dtype)
        approx = tensor_pass_scores(q, c).double()               # dot_tc * fl32(1 / ||c||)
        qd, cd = q.double(), c.double()
        qn, cn = qd.norm(dim=-1, keepdim=True), cd.norm(dim=-1, keepdim=True)
        exact = (qd @ cd.T) / cn.T                               # exact dot / ||c||, float64 on the stored values
        err = ((approx - exact).abs() / qn).max().item()         # in units of ||q|| ||c||
        worst[name] = err
        assert err <= 0.25 * eps, (name, D, dtype, err, eps)
    print(f"D={D} {dtype}: eps {eps:.3e}, measured max error / (|q||c|): " +
          ", ".join(f"{k} {v:.2e}" for k, v in worst.items()))


@pytest.mark.parametrize("D", [384, 768, 4096])
def test_shadow_error_is_within_shadow_eps(D):
    """fp32 rows searched through bf16 shadows: |dot(qs, cs) / ||cs|| - dot(q, c) / ||c||| <= shadow_eps * ||q||.
    Random rows sit far inside the bound (<= 1/4); rows whose every element rounds the same way by almost half a
    bf16 ulp are the bound's worst case and must still be inside it."""
    from text_similarity_b200 import _lib
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(77 + D)
    Q, N = 64, 2 * SLICE
    eps = _lib.load().tsim_debug_eps(D, _lib.BF16, 1)
    assert eps >= 5.99e-3
    fam = _families(Q, N, D, g, dev)
    # adversarial rounding: bf16 grid values pushed up by 0.49 ulp -> the shadow rounds every element down
    yq, xr = fam["near_duplicates"]
    up = lambda t: t.to(torch.bfloat16).float() * (1 + 0.49 * 2 ** -8)   # noqa: E731
    fam["one_sided_rounding"] = (up(yq), xr.to(torch.bfloat16).float())
    for name, (q32, c32) in fam.items():
        qs, cs = q32.to(torch.bfloat16), c32.to(torch.bfloat16)
        approx = tensor_pass_scores(qs, cs).double()
        qd, cd = q32.double(), c32.double()
        exact = (qd @ cd.T) / cd.norm(dim=-1, keepdim=True).T
        err = ((approx - exact).abs() / qd.norm(dim=-1, keepdim=True)).max().item()
        bar = eps if name == "one_sided_rounding" else 0.25 * eps
        assert err <= bar, (name, D, err, eps)
        print(f"shadow D={D} {name}: max error {err:.2e} (bound {eps:.2e})")


def test_eps_grows_with_width_not_below_floor():
    from text_similarity_b200 import _lib
    lib = _lib.load()
    assert lib.tsim_debug_eps(768, _lib.BF16, 0) == pytest.approx(5e-5)
    assert lib.tsim_debug_eps(16384, _lib.BF16, 0) == pytest.approx(1024 * 2 ** -21)
    assert lib.tsim_debug_eps(16384, _lib.E4M3, 0) == pytest.approx(512 * 2 ** -21)
    assert lib.tsim_debug_eps(768, _lib.BF16, 1) == pytest.approx(6e-3)
    assert lib.tsim_debug_eps(16384, _lib.BF16, 1) > 6e-3
