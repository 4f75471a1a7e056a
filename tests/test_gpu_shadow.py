"""GPU parity tests of the shadow path: fp32 / fp16 rows searched at tensor-core speed (candidates from bf16
shadows, float64 re-score on the originals) must give exactly what the float64 exact scan of the ORIGINAL
rows gives -- same indices, same float64 score bits -- and agree with the CPU oracle within 1e-5."""
import numpy as np
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


def _rows(n, d, seed, normalize=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, d, generator=g, device="cuda")
    return x / x.norm(dim=-1, keepdim=True) if normalize else x


def _same_as_exact(ops, q, c, k, max_flag_frac=0.05, **kw):
    shadow, sinv = ops.make_shadow(c)
    a = ops.search_topk(q, c, k, corpus_shadow=shadow, shadow_inv_norm=sinv, return_score64=True, return_flags=True, **kw)
    b = ops.search_topk(q, c, k, mode="exact", return_score64=True, **kw)
    torch.cuda.synchronize()
    assert torch.equal(a[1], b[1]), f"{(a[1] != b[1]).sum().item()} index mismatches vs the exact scan"
    assert torch.equal(a[2], b[2])
    assert a[3].float().mean().item() <= max_flag_frac, a[3].float().mean().item()
    return a


@pytest.mark.parametrize("N,Q,D,k,dtype", [(300_000, 50, 384, 10, torch.float32), (120_000, 300, 768, 10, torch.float32),
                                            (200_000, 17, 384, 24, torch.float32), (150_000, 64, 256, 5, torch.float16),
                                            (700, 5, 64, 10, torch.float32)])
def test_shadow_equals_exact_scan(ops, N, Q, D, k, dtype):
    c = _rows(N, D, 1).to(dtype)
    c[N // 2] = c[3]                                   # an exact duplicate pair (tie -> lower row first)
    q = _rows(Q, D, 2).to(dtype)
    q[0] = c[3]
    a = _same_as_exact(ops, q, c, k, max_flag_frac=0.05 if N > 1000 else 1.0)
    sel = torch.arange(0, Q, max(1, Q // 4))
    ev, ei = O.search_exact(q[sel].cpu(), c.cpu(), k)
    assert torch.equal(a[1][sel].cpu(), ei)
    np.testing.assert_allclose(a[0][sel].cpu().numpy(), ev.numpy(), rtol=0, atol=1e-5)   # north_star: 1e-5 for fp32
    assert a[1][0, 0].item() == 3 and a[1][0, 1].item() == N // 2


def test_shadow_rows_that_differ_below_bf16_resolution(ops):
    # 60 rows that round to the SAME bf16 shadow but differ in fp32: the tensor pass sees exact ties, the float64
    # re-score on the originals must still order them; plus unnormalised rows with wildly different norms
    N, D, k = 100_000, 384, 10
    c = _rows(N, D, 5, normalize=False) * (torch.rand(N, 1, device="cuda") * 9 + 0.1)
    q = _rows(8, D, 6)
    base = q[0] * 3.0
    pert = _rows(60, D, 7) * 2e-4
    c[1000:1060] = base[None] + pert
    assert (c[1000:1060].to(torch.bfloat16) == c[1000].to(torch.bfloat16)).float().mean().item() > 0.9
    _same_as_exact(ops, q, c, k, max_flag_frac=0.5)


def test_shadow_k_beyond_the_candidate_budget_runs_the_exact_scan(ops):
    c, q = _rows(50_000, 128, 8), _rows(9, 128, 9)
    shadow, sinv = ops.make_shadow(c)
    a = ops.search_topk(q, c, 50, corpus_shadow=shadow, shadow_inv_norm=sinv, return_score64=True)
    b = ops.search_topk(q, c, 50, mode="exact", return_score64=True)
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])


def test_fp32_corpora_take_the_shadow_path_in_the_sharded_and_pipeline_apis(ops):
    from src.configurations.config import ModelParameters, SearchConfiguration
    from src.pipeline.search_pipeline import SentenceMiningPipeline
    from text_similarity_b200.sharded import ShardedCorpus
    c, q = _rows(200_000, 384, 10), _rows(33, 384, 11)
    sc = ShardedCorpus(c, idx_base=500)
    assert sc.shadow is not None and sc.shadow.dtype == torch.bfloat16
    s, i = sc.search(q, 10)
    es, ei = ops.search_topk(q, c, 10, idx_base=500, mode="exact")
    assert torch.equal(i, ei) and torch.equal(s, es)
    gs, gi = sc.search_graphed(q, 10)
    assert torch.equal(gi, ei) and torch.equal(gs, es)
    params = SearchConfiguration(model_parameters=ModelParameters(model_name="none"), model="none", save_path=".",
                                 device=torch.device("cuda"))
    pipe = SentenceMiningPipeline(60_000, params=params, model=None, name="t")
    ps, pi = pipe.search_tensors(q, 10, corpus=c)                       # 4 chunks, each with its own shadow, merged
    es0, ei0 = ops.search_topk(q, c, 10, mode="exact")
    assert torch.equal(pi, ei0) and torch.equal(ps, es0)


# ---------------------------------------------------------------------------- split (hi + lo) shadow: k up to 100
def _split_same_as_exact(ops, q, c, k, max_flag_frac=0.05, **kw):
    shadow, sinv = ops.make_shadow(c, split=True)
    assert shadow.shape == (c.shape[0], 2 * ((c.shape[1] + 63) // 64 * 64)) and shadow.dtype == torch.bfloat16
    a = ops.search_topk(q, c, k, corpus_shadow=shadow, shadow_inv_norm=sinv, return_score64=True, return_flags=True, **kw)
    b = ops.search_topk(q, c, k, mode="exact", return_score64=True, **kw)
    torch.cuda.synchronize()
    assert torch.equal(a[1], b[1]), f"{(a[1] != b[1]).sum().item()} index mismatches vs the exact scan"
    assert torch.equal(a[2], b[2])
    assert a[3].float().mean().item() <= max_flag_frac, a[3].float().mean().item()
    return a


@pytest.mark.parametrize("N,Q,D,k,dtype", [(300_000, 200, 768, 100, torch.float32), (120_000, 1100, 384, 50, torch.float32),
                                            (500_000, 40, 256, 100, torch.float32), (150_000, 64, 256, 30, torch.float16),
                                            (2_000, 5, 64, 100, torch.float32), (200_000, 300, 768, 10, torch.float32),
                                            (60_000, 33, 100, 20, torch.float32), (40_000, 150, 203, 100, torch.float32)])
def test_split_shadow_matches_exact_scan(ops, N, Q, D, k, dtype):
    c = _rows(N, D, 61).to(dtype)
    c[N // 2:N // 2 + 30] = c[10:40]                 # exact duplicates
    q = _rows(Q, D, 62).to(dtype)
    q[1] = c[12]
    a = _split_same_as_exact(ops, q, c, k)
    # ... and the CPU oracle on a few queries (tolerance north_star gives fp32 inputs: 1e-5)
    sel = torch.tensor([0, 1, Q - 1])
    ev, ei = O.search_exact(q[sel.cuda()].cpu().float(), c.cpu().float(), k)
    assert torch.equal(a[1][sel.cuda()].cpu(), ei)
    np.testing.assert_allclose(a[0][sel.cuda()].cpu().numpy(), ev.numpy(), atol=1e-5)


def test_split_shadow_resolves_rows_closer_than_bf16(ops):
    """Rows that differ from the best row only far below bf16 resolution: the rounded shadow cannot tell them apart
    (it needs its 112 candidates and the float64 re-score), the split shadow's approximate scores already order them
    to ~1e-5; both must return the exact answer."""
    N, D, k = 300_000, 384, 40
    c = _rows(N, D, 63)
    qv = _rows(1, D, 64)
    for j in range(60):                                        # 60 near-copies of the query, 2e-4 apart in cosine
        noise = _rows(1, D, 700 + j)
        v = qv + 0.02 * (j + 1) ** 0.5 * noise
        c[1000 + 37 * j] = v / v.norm()
    _split_same_as_exact(ops, qv, c, k)


def test_split_shadow_error_bound(ops):
    """|split-shadow dot - exact dot| / (||q|| ||c||) against split_shadow_eps, on unit-norm, wide-dynamic-range and
    near-duplicate rows (the tensor pass itself is exercised through tsim_debug_tensor_pass on the shadow arrays)."""
    import ctypes
    from text_similarity_b200 import _lib
    lib = _lib.load()
    for D in (64, 768, 2048):
        g = torch.Generator(device="cuda").manual_seed(D)
        c = torch.randn(112, D, generator=g, device="cuda") * torch.logspace(-2, 1, D, device="cuda")
        q = torch.randn(8, D, generator=g, device="cuda")
        c[5] = q[3] * 1.0000001
        c[6] = -q[3]
        ch, cl = ops._split_bf16(c)
        qh, ql = ops._split_bf16(q)
        cs, qs = torch.cat([ch, cl, ch], 1).contiguous(), torch.cat([qh, qh, ql], 1).contiguous()
        approx = (qs.double() @ cs.double().T)                 # what an error-free three-segment bf16 pass would compute
        exact = q.double() @ c.double().T
        scale = q.double().norm(dim=1)[:, None] * c.double().norm(dim=1)[None, :]
        err = ((approx - exact).abs() / scale).max().item()
        eps = lib.tsim_debug_eps(D, _lib.BF16, 2)
        assert err <= 1.5e-5 and eps >= 1.5e-5 + 5e-5 - 1e-9, (D, err, eps)
