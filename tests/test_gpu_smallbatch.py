"""GPU parity of the small-batch kernel (search_sw.cu: Q <= 32 on a shard of >= 16 * SMs * 128 rows -- BASELINE
config 4's regime): corpus rows on the MMA's M side, queries resident on the N side, sample lists + append lists.
Compared bit for bit (indices and float64 scores) with the float64 exact scan of the same library, which
tests/test_gpu_parity.py pins to the CPU oracle; the smaller cases are also checked against the oracle directly."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

N_SW = 320_000          # > 16 * 148 * 128 = 303,104 rows: the planner picks the small-batch kernel (rows of at most 512
                        # bytes at any Q <= 32, wider rows at 8 <= Q <= 32; the other cases below run search_tc.cu)


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from text_similarity_b200 import ops as _ops
    return _ops


def _rows(n, d, seed, dtype, unit=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, d, generator=g, device="cuda")
    if unit:
        x = x / x.norm(dim=-1, keepdim=True)
    return x.to(dtype)


def _launches(ops, fn):
    from text_similarity_b200 import _lib
    lib = _lib.load()
    before = lib.tsim_launch_count()
    out = fn()
    torch.cuda.synchronize()
    return out, lib.tsim_launch_count() - before


def _same_as_exact(ops, q, c, k, **kw):
    a = ops.search_topk(q, c, k, mode="tensor", return_score64=True, return_flags=True, **kw)
    b = ops.search_topk(q, c, k, mode="exact", return_score64=True, **kw)
    torch.cuda.synchronize()
    assert torch.equal(a[1], b[1]), f"{(a[1] != b[1]).sum().item()} index mismatches"
    assert torch.equal(a[2], b[2])
    return a


@pytest.mark.parametrize("Q", [1, 7, 32])
@pytest.mark.parametrize("dtype,D", [(torch.bfloat16, 768), (torch.float8_e4m3fn, 384), (torch.bfloat16, 136)])
def test_small_batch_matches_exact_scan(ops, Q, dtype, D):
    if dtype == torch.float8_e4m3fn:
        c = _rows(N_SW + 77, D, 1, torch.float32) * 16.0     # e4m3 rows keep a power-of-two scale, as K1 stores them
        c, q = c.to(dtype), (_rows(Q, D, 2, torch.float32) * 16.0).to(dtype)
    else:
        c, q = _rows(N_SW + 77, D, 1, dtype), _rows(Q, D, 2, dtype)      # ragged last tile (77 rows)
    # exact duplicates (real ties), a planted best row in the last tile and in a sample tile (row 0)
    c[200_000:200_020] = c[5:25]
    c[N_SW + 70] = q[0]
    c[0] = q[Q - 1]
    s, i, s64, fl = _same_as_exact(ops, q, c, 10)
    # (Q = 1: both planted rows equal the one query -- the tie goes to the lower row)
    assert i[0, 0].item() == (N_SW + 70 if Q > 1 else 0) and i[Q - 1, 0].item() == 0
    assert (N_SW + 70) in i[0, :2].tolist()
    assert (fl == 0).all()                                   # nothing needed the retry pass or the scan
    d = s64[:, 1:] - s64[:, :-1]
    assert (d <= 0).all() and ((i[:, 1:] > i[:, :-1]) | (d < 0)).all()


@pytest.mark.parametrize("k", [1, 24])
def test_small_batch_other_k(ops, k):
    """k <= 24 (candidate proofs over 16 or 32 rows) takes the same kernel; k = 25 and beyond the list kernel."""
    c, q = _rows(N_SW, 256, 21, torch.bfloat16), _rows(12, 256, 22, torch.bfloat16)
    c[250_000:250_030] = c[40:70]
    s, i, s64, fl = _same_as_exact(ops, q, c, k)
    assert (fl == 0).all()


def test_small_batch_uses_the_swapped_kernel(ops):
    """The plan for Q <= 32 on a large shard is the 3-launch chain (sample, thresholds, main): fewer MMAs, and
    select_rescore reads 192 sample lists + the append list instead of 4 lists per SM."""
    c, q = _rows(N_SW, 384, 3, torch.bfloat16), _rows(16, 384, 4, torch.bfloat16)
    inv = ops.row_inv_norm(c)
    ops.search_topk(q, c, 10, corpus_inv_norm=inv)
    _, n_small = _launches(ops, lambda: ops.search_topk(q, c, 10, corpus_inv_norm=inv))
    # prep, sample, tighten, main, select, retry (search + select), scan, merge of the scan's lists
    assert n_small == 9, n_small


def test_small_batch_vs_oracle_with_self_exclusion_and_shard_base(ops):
    N, D = N_SW, 64
    g = torch.Generator().manual_seed(5)
    c = torch.randn(N, D, generator=g)
    c = (c / c.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
    c[123_456] = c[17]                                       # a duplicate of a query row
    q = c[10:30].clone()                                     # queries are corpus rows 10..29 of this shard
    base = 1_000_000
    s, i, s64 = ops.search_topk(q.cuda(), c.cuda(), 5, mode="tensor", idx_base=base, exclude_self_base=base + 10,
                                return_score64=True)
    ev, ei = O.search_exact(q, c, 5, idx_base=base, exclude_self_base=base + 10)
    assert torch.equal(i.cpu(), ei)
    np.testing.assert_allclose(s64.cpu().numpy(), ev.numpy(), rtol=0, atol=1e-12)
    assert i[7, 0].item() == base + 123_456                  # query 7 = row 17: its twin wins, itself is excluded


def test_small_batch_zero_nan_and_unnormalised_rows(ops):
    N, D = N_SW, 128
    c = _rows(N, D, 6, torch.float32, unit=False)
    c[1000] = 0.0                                            # zero row: cosine 0 by the eps clamp
    c[2000, 3] = float("nan")                                # NaN rows are never returned
    c[128 * 5 + 1, 0] = float("nan")
    c = c.to(torch.bfloat16)
    q = _rows(5, D, 7, torch.float32, unit=False).to(torch.bfloat16)
    s, i, s64, fl = _same_as_exact(ops, q, c, 10)
    assert not torch.isnan(s64).any() and (i != 2000).all() and (i != 128 * 5 + 1).all()


def test_small_batch_heavy_duplicates_go_through_retry(ops):
    """40 copies of one row tie at the top: the 16-candidate proof fails, the wide retry pass answers."""
    c, q = _rows(N_SW, 256, 8, torch.bfloat16), _rows(4, 256, 9, torch.bfloat16)
    c[100_000:100_040] = q[2]
    s, i, s64, fl = _same_as_exact(ops, q, c, 10)
    assert i[2].tolist() == list(range(100_000, 100_010))
    assert fl[2].item() != 0


def test_small_batch_shards_plus_merge_equal_single_search(ops):
    N, D, Q, k, G = 8 * N_SW, 384, 32, 10, 8
    c = (_rows(N, D, 10, torch.float32) * 16.0).to(torch.float8_e4m3fn)
    q = (_rows(Q, D, 11, torch.float32) * 16.0).to(torch.float8_e4m3fn)
    full = ops.search_topk(q, c, k, return_score64=True)
    per = N // G
    parts = [ops.search_topk(q, c[r * per:(r + 1) * per], k, idx_base=r * per, return_score64=True) for r in range(G)]
    ms, ms64, mi = ops.merge_topk(torch.cat([p[2] for p in parts], 1), torch.cat([p[1] for p in parts], 1), k, G)
    assert torch.equal(mi, full[1]) and torch.equal(ms64, full[2])


def test_small_batch_graph_replay_and_two_streams(ops):
    """The 3-launch chain replays from a CUDA graph (ShardedCorpus.search_graphed) with fresh queries each time, and
    two streams can run small-batch searches side by side (no grid barrier in these kernels, one workspace per stream)."""
    from text_similarity_b200.sharded import ShardedCorpus
    c = (_rows(N_SW, 384, 31, torch.float32) * 16.0).to(torch.float8_e4m3fn)
    corp = ShardedCorpus(c)
    for seed in (32, 33, 34):
        q = (_rows(8, 384, seed, torch.float32) * 16.0).to(torch.float8_e4m3fn)
        gs, gi = corp.search_graphed(q, 10)
        gs, gi = gs.clone(), gi.clone()
        es, ei = ops.search_topk(q, c, 10, mode="exact")
        assert torch.equal(gi, ei) and torch.equal(gs, es)
    qa = (_rows(16, 384, 41, torch.float32) * 16.0).to(torch.float8_e4m3fn)
    qb = (_rows(3, 384, 42, torch.float32) * 16.0).to(torch.float8_e4m3fn)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    outs = {}
    for _ in range(3):
        with torch.cuda.stream(sa):
            outs["a"] = ops.search_topk(qa, c, 10, corpus_inv_norm=corp.inv_norm)
        with torch.cuda.stream(sb):
            outs["b"] = ops.search_topk(qb, c, 10, corpus_inv_norm=corp.inv_norm)
    torch.cuda.synchronize()
    for key, q in (("a", qa), ("b", qb)):
        es, ei = ops.search_topk(q, c, 10, mode="exact")
        assert torch.equal(outs[key][1], ei) and torch.equal(outs[key][0], es)


def test_small_batch_randomized_shapes(ops):
    """Seeded random shapes around the planner's boundaries (shard size, row width, Q, k, dtype, un-normalised rows,
    strided views): the tensor path -- whichever kernel the planner picks -- must equal the float64 exact scan."""
    rng = np.random.default_rng(1234)
    for case in range(8):
        fp8 = bool(rng.integers(0, 2))
        D = int(rng.choice([64, 128, 272, 384, 512, 768, 1024])) if not fp8 else int(rng.choice([64, 128, 384, 512, 1040]))
        N = int(rng.integers(300_000, 420_000))
        Q = int(rng.integers(1, 33))
        k = int(rng.choice([1, 5, 10, 24]))
        unit = bool(rng.integers(0, 2))
        scale = 16.0 if fp8 else 1.0
        dtype = torch.float8_e4m3fn if fp8 else torch.bfloat16
        base = (_rows(N, D + 16, 100 + case, torch.float32, unit=unit) * scale).to(dtype)
        c = base[:, :D]                                        # a strided view: row pitch D + 16 elements
        q = (_rows(Q, D, 200 + case, torch.float32, unit=unit) * scale).to(dtype)
        src = torch.randint(0, N, (50,), device="cuda")
        c[torch.randint(0, N, (50,), device="cuda")] = c[src]  # exact duplicates
        s, i, s64, fl = _same_as_exact(ops, q, c, k)
        assert (i >= 0).all() and (i < N).all(), (case, fp8, D, N, Q, k)


def test_small_batch_dense_cluster_is_retightened_not_overflowed(ops):
    """6000 rows crowd around one query, far above anything the 3072-row sample suggests.  Its append list would
    overflow (4096 entries -> flagged -> a second scan by the retry pass); instead the kernel re-makes the query's
    threshold from the list every 512 entries, and the first pass answers -- exactly.  (A shard big enough for the
    scan to outlast that reaction: 4M rows.)"""
    N, D = 4_000_000, 64
    c = _rows(N, D, 71, torch.float32)
    q = _rows(3, D, 72, torch.float32)
    g = torch.Generator(device="cuda").manual_seed(73)
    where = torch.randperm(N, generator=g, device="cuda")[:6000]
    noise = torch.randn(6000, D, generator=g, device="cuda")
    noise = noise / noise.norm(dim=-1, keepdim=True)
    # (square-root spacing: the best rows are ~2e-4 apart in cosine, so the 16-candidate proof holds; linear spacing
    # would put them 3e-6 apart and flag the query for a reason that has nothing to do with the list)
    eps_ = (0.05 + 0.4 * torch.linspace(0, 1, 6000, device="cuda") ** 0.5)[:, None]
    v = q[1][None, :] + eps_ * noise
    c[where] = v / v.norm(dim=-1, keepdim=True)
    cb, qb = c.to(torch.bfloat16), q.to(torch.bfloat16)
    s, i, s64, fl = _same_as_exact(ops, qb, cb, 10)
    assert (fl == 0).all()                                   # answered by the first pass, nothing overflowed
    assert set(i[1].tolist()) <= set(where.tolist())
