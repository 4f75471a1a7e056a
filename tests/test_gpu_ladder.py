"""GPU parity tests of the bootstrap + threshold-ladder schedules of the tcgen05 search (the launch
plans large corpora take: sample pass -> tighten -> main pass whose thresholds follow the per-query
rank counters).  The checker is the float64 exact scan (mode="exact"), itself pinned to the CPU
oracle by test_gpu_parity.py, plus the oracle directly on a subset of the queries: indices and
float64 score bits must be identical."""
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from text_similarity_b200 import ops as _ops
    return _ops


def _rows(n, d, seed, dtype, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, d, generator=g, device="cuda")
    x = x / x.norm(dim=-1, keepdim=True) * scale
    return x.to(dtype)


def _agree(ops, q, c, k, oracle_queries=4, max_flag_frac=0.05, **kw):
    a = ops.search_topk(q, c, k, mode="tensor", return_score64=True, return_flags=True, **kw)
    b = ops.search_topk(q, c, k, mode="exact", return_score64=True, **kw)
    torch.cuda.synchronize()
    assert torch.equal(a[1], b[1]), f"{(a[1] != b[1]).sum().item()} index mismatches vs the exact scan"
    assert torch.equal(a[2], b[2])
    if oracle_queries:
        sel = torch.linspace(0, q.shape[0] - 1, oracle_queries).long()
        ev, ei = O.search_exact(q[sel.cuda()].cpu(), c.cpu(), k, idx_base=kw.get("idx_base", 0))
        assert torch.equal(a[1][sel.cuda()].cpu(), ei)
    assert a[3].float().mean().item() <= max_flag_frac
    return a


@pytest.mark.parametrize("Q,k", [(16, 10), (48, 10), (128, 5), (200, 24), (130, 100)])
def test_sticky_bootstrap_ladder_bf16(ops, Q, k):
    # few query blocks -> sticky schedule; T = 5079 tiles >= 32 tiles per worker -> bootstrap + ladder
    N, D = 1_300_000, 64
    c = _rows(N, D, 101, torch.bfloat16)
    c[1_200_000:1_200_040] = c[5:45]             # exact duplicates (score ties)
    q = _rows(Q, D, 202 + Q, torch.bfloat16)
    q[1] = c[7]                                   # a query whose best rows are an exact tie pair
    _agree(ops, q, c, k)


@pytest.mark.parametrize("Q,k", [(1280, 10), (1280, 100), (1400, 24), (1300, 50)])
def test_round_robin_bootstrap_ladder_bf16(ops, Q, k):
    # 5-6 query blocks of 256 on 74 CTA pairs -> round-robin units; T = 1954 tiles -> sample + ladder
    N, D = 500_000, 64
    c = _rows(N, D, 303, torch.bfloat16)
    c[400_000:400_100] = c[1000:1100]
    q = _rows(Q, D, 404 + k, torch.bfloat16)
    _agree(ops, q, c, k, idx_base=12_345)


def test_ladder_fp8_small_batch(ops):
    N, D, k = 1_300_000, 64, 10
    c = _rows(N, D, 505, torch.float8_e4m3fn, scale=64.0)
    q = _rows(32, D, 606, torch.float8_e4m3fn, scale=64.0)
    _agree(ops, q, c, k)


def test_ladder_rising_scores_and_massive_ties(ops):
    # (a) rows ordered by increasing similarity to query 0: its threshold must keep rising through the
    # main pass; (b) 3000 copies of one row: every ladder level collapses onto one value for query 1,
    # the completeness proof fails and the exact scan answers -- results stay exact either way.
    N, D, k = 1_300_000, 64, 10
    g = torch.Generator(device="cuda").manual_seed(7)
    c = _rows(N, D, 707, torch.float32)
    q = _rows(40, D, 808, torch.float32)
    qv = q[0] / q[0].norm()
    alpha = torch.linspace(-0.5, 0.95, N, device="cuda")[:, None]
    noise = c - (c @ qv)[:, None] * qv
    noise = noise / noise.norm(dim=-1, keepdim=True)
    c = alpha * qv + (1 - alpha ** 2).sqrt() * noise
    dup = torch.randperm(N, generator=g, device="cuda")[:3000]
    c[dup] = q[1]
    _agree(ops, q.to(torch.bfloat16), c.to(torch.bfloat16), k, max_flag_frac=0.2)


def test_ladder_can_be_switched_off_and_results_do_not_change(ops, monkeypatch):
    N, D, Q, k = 600_000, 64, 1280, 10
    c = _rows(N, D, 909, torch.bfloat16)
    q = _rows(Q, D, 910, torch.bfloat16)
    a = ops.search_topk(q, c, k, mode="tensor", return_score64=True)
    monkeypatch.setenv("TSIM_NO_LADDER", "1")
    b = ops.search_topk(q, c, k, mode="tensor", return_score64=True)
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])


@pytest.mark.parametrize("Q,N,D,k", [(1280, 500_000, 64, 10), (4096, 150_000, 128, 100), (300, 40_000, 768, 10)])
def test_claimed_units_equal_static_dealing(ops, monkeypatch, Q, N, D, k):
    # Round-robin plans hand units to workers through a global counter (search_tc.cu get_unit); which
    # worker scans a unit must not change the result: identical indices and float64 score bits with the
    # static dealing (TSIM_STATIC_UNITS=1, read per call), and both equal to the exact scan.
    c = _rows(N, D, 808, torch.bfloat16)
    c[N // 2:N // 2 + 50] = c[100:150]
    q = _rows(Q, D, 909 + k, torch.bfloat16)
    a = _agree(ops, q, c, k, oracle_queries=2)
    monkeypatch.setenv("TSIM_STATIC_UNITS", "1")
    b = ops.search_topk(q, c, k, mode="tensor", return_score64=True, return_flags=True)
    torch.cuda.synchronize()
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    # back-to-back calls reuse the workspace: the claim areas are re-zeroed by every call
    monkeypatch.delenv("TSIM_STATIC_UNITS")
    for _ in range(3):
        d = ops.search_topk(q, c, k, mode="tensor", return_score64=True)
    torch.cuda.synchronize()
    assert torch.equal(a[1], d[1]) and torch.equal(a[2], d[2])
