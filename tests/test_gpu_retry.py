"""GPU tests of the retry stage of the tensor path: queries whose first-pass candidate set cannot be
proven complete (ties straddling ranks k..KP -- a sentence duplicated many times in the corpus) are
re-run as one compact block with 112-entry lists before anything falls back to the float64 scan
(tsim_api.cu search_impl).  The checker is the float64 exact scan (mode="exact"), itself pinned to the
CPU oracle by test_gpu_parity.py, plus the oracle directly on a few queries: indices and float64 score
bits must be identical whichever stage answered.  out_flags: 0 first pass, 2 retry pass, 1 float64 scan."""
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


def _rows(n, d, seed, dtype=torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, d, generator=g, device="cuda")
    return (x / x.norm(dim=-1, keepdim=True)).to(dtype)


def _dup_corpus(N, D, groups, seed):
    """Corpus in which, for (src, copies) in groups, row src re-appears `copies` times at scattered rows."""
    c = _rows(N, D, seed)
    g = torch.Generator().manual_seed(seed + 1)
    for src, copies in groups:
        pos = torch.randperm(N - 1000, generator=g)[:copies] + 1000
        c[pos.cuda()] = c[src].clone()
    return c


def _check(ops, q, c, k, **kw):
    a = ops.search_topk(q, c, k, mode="tensor", return_score64=True, return_flags=True, **kw)
    b = ops.search_topk(q, c, k, mode="exact", return_score64=True, **kw)
    torch.cuda.synchronize()
    assert torch.equal(a[1], b[1]), f"{(a[1] != b[1]).sum().item()} index mismatches vs the exact scan"
    assert torch.equal(a[2], b[2])
    return a


@pytest.mark.parametrize("Q,N,D,k", [(8, 200_000, 128, 10), (300, 200_000, 128, 10), (1400, 400_000, 64, 10),
                                     (64, 150_000, 768, 24)])
def test_duplicates_answered_by_retry_pass(ops, Q, N, D, k):
    # rows 0..3 duplicated 40-60x: a query equal to such a row has that many exact ties at the top, more
    # than the first pass's lists hold (KP = 16 / 32), fewer than the retry pass's 112
    c = _dup_corpus(N, D, [(0, 40), (1, 45), (2, 60), (3, 50)], 11)
    q = _rows(Q, D, 22)
    q[0], q[1], q[2], q[Q - 1] = c[0], c[1], c[2], c[3]
    a = _check(ops, q, c, k)
    fl = a[3].cpu()
    assert fl[0] == 2 and fl[1] == 2 and fl[2] == 2 and fl[Q - 1] == 2, fl[[0, 1, 2, Q - 1]]
    assert int((fl == 1).sum()) == 0
    ev, ei = O.search_exact(q[:3].cpu(), c.cpu(), k)
    assert torch.equal(a[1][:3].cpu(), ei)
    # ties come back lowest row first: row 0 itself, then its copies in ascending order
    assert a[1][0, 0].item() == 0 and bool((a[1][0, 1:] > a[1][0, :-1]).all())


def test_retry_then_float64_scan(ops):
    # 150 copies defeat the 112-entry lists too: that query is answered by the float64 scan (flag 1),
    # its 30-copy neighbour by the retry pass (flag 2), ordinary queries by the first pass (flag 0)
    N, D, k = 300_000, 128, 10
    c = _dup_corpus(N, D, [(0, 150), (1, 30)], 33)
    q = _rows(40, D, 44)
    q[5], q[6] = c[0], c[1]
    a = _check(ops, q, c, k)
    fl = a[3].cpu()
    assert fl[5] == 1 and fl[6] == 2 and int((fl == 0).sum()) == 38


@pytest.mark.parametrize("Q,n_retry,n_scan", [(100, 100, 0), (200, 200, 0), (700, 512, 188)])
def test_retry_rounds_and_overflow(ops, Q, n_retry, n_scan):
    # every query is flagged; the retry stage runs one round of 128 per 128 queries of the call, at most 4:
    # what overflows is answered by the float64 scan
    N, D, k = 120_000, 64, 10
    c = _dup_corpus(N, D, [(i, 20) for i in range(8)], 55)
    q = c[torch.arange(Q, device="cuda") % 8].clone()
    a = _check(ops, q, c, k)
    fl = a[3].cpu()
    assert int((fl == 2).sum()) == n_retry and int((fl == 1).sum()) == n_scan


def test_retry_with_self_exclusion_and_idx_base(ops):
    # all-pairs use: query i IS corpus row base + i and must not return itself, also in the retry pass
    N, D, k = 100_000, 128, 5
    c = _dup_corpus(N, D, [(2000, 30), (2001, 30)], 66)
    q = c[2000:2100].contiguous()
    a = _check(ops, q, c, k, idx_base=7_000_000, exclude_self_base=7_002_000)
    fl = a[3].cpu()
    assert fl[0] == 2 and fl[1] == 2
    assert not bool((a[1] == (torch.arange(100, device="cuda") + 7_002_000)[:, None]).any())


def test_retry_fp8(ops):
    N, D, k = 200_000, 128, 10
    c = _dup_corpus(N, D, [(0, 30)], 77)
    c8 = (c.float() * 64).to(torch.float8_e4m3fn)
    q = _rows(16, D, 88)
    q8 = (q.float() * 64).to(torch.float8_e4m3fn)
    q8.view(torch.uint8)[3] = c8.view(torch.uint8)[0]
    a = _check(ops, q8, c8, k)
    assert a[3][3].item() == 2


def test_release_library_ignores_the_no_retry_knob(ops, monkeypatch):
    """TSIM_NO_RETRY exists only in the -DTSIM_EXPERIMENT flavour: the release library answers the duplicated
    query with the retry pass whatever the environment says (no silent change of path), and the float64 scan
    (mode="exact", inside _check) gives the same bits."""
    N, D, k = 200_000, 128, 10
    c = _dup_corpus(N, D, [(0, 30)], 99)
    q = _rows(32, D, 100)
    q[7] = c[0]
    a = _check(ops, q, c, k)
    monkeypatch.setenv("TSIM_NO_RETRY", "1")
    monkeypatch.setenv("TSIM_DEBUG", "6")      # would skip the MMAs and the epilogue in the experiment build
    b = _check(ops, q, c, k)
    assert a[3][7].item() == 2 and b[3][7].item() == 2
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    from text_similarity_b200 import _lib
    assert _lib.counters()["env_reads"] == 0


def test_repeated_planned_search_encodes_no_descriptor(ops):
    """Plan handles (tsim_plan_create): a repeated search over the same arrays makes no launch plan and encodes no
    TMA descriptor (SURVEY.md 8b ownership row)."""
    from text_similarity_b200 import _lib
    c = _rows(300_000, 128, 5)
    inv = ops.row_inv_norm(c)
    q = _rows(256, 128, 6)                      # a multiple of the 256-query pair block: no padded copy
    out = ops.search_topk(q, c, 10, corpus_inv_norm=inv, mode="tensor")
    torch.cuda.synchronize()
    before = _lib.counters()
    for _ in range(3):
        again = ops.search_topk(q, c, 10, corpus_inv_norm=inv, mode="tensor")
    torch.cuda.synchronize()
    after = _lib.counters()
    assert torch.equal(out[1], again[1])
    assert after["plans"] == before["plans"] and after["map_encodes"] == before["map_encodes"]
    assert after["launches"] > before["launches"]
