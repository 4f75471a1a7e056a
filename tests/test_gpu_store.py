"""GPU tests of the corpus embedding store: add / remove / count, search by
label, encode-side fusion into the store's tail, and the on-disk round trip (whole and sharded)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from text_similarity_b200 import ops, store
    return ops, store


def _unit(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g)
    return x / x.norm(dim=-1, keepdim=True)


def test_add_search_remove_matches_oracle_on_the_live_rows(mods):
    ops, store = mods
    D = 128
    st = store.EmbeddingStore(D, torch.bfloat16, "cuda", capacity=8)       # forces several growths
    x = _unit(3000, D, 1).to(torch.bfloat16)
    lab = st.add(x[:1000].cuda())
    assert lab.tolist() == list(range(1000)) and len(st) == 1000
    st.add(x[1000:].cuda(), ids=range(5000, 7000))
    assert st.num_indexed() == 3000 and st.capacity >= 3000
    q = _unit(17, D, 2).to(torch.bfloat16)
    s, labels = st.search(q.cuda(), 10)
    ev, ei = O.search_exact(q, x, 10)
    all_labels = torch.cat([torch.arange(1000), torch.arange(5000, 7000)])
    assert torch.equal(labels.cpu(), all_labels[ei])
    np.testing.assert_allclose(s.cpu().numpy(), ev.numpy(), atol=1e-3)
    # remove every query's best hit plus some unknown labels: the store stays dense, results follow
    best = labels[:, 0].cpu().tolist()
    removed = st.remove(best + [123456, -5])
    assert removed == len(set(best)) and len(st) == 3000 - len(set(best))
    live = torch.ones(3000, dtype=torch.bool)
    live[torch.tensor([int((all_labels == b).nonzero()) for b in set(best)])] = False
    s2, labels2 = st.search(q.cuda(), 10)
    ev2, ei2 = O.search_exact(q, x[live], 10)
    # ties aside (none in this data), the SET and ORDER of labels equal the oracle's on the live rows
    assert torch.equal(labels2.cpu(), all_labels[live][ei2])
    np.testing.assert_allclose(s2.cpu().numpy(), ev2.numpy(), atol=1e-3)
    assert not set(labels2.cpu().reshape(-1).tolist()) & set(best)
    with pytest.raises(ValueError):
        st.add(x[:2].cuda(), ids=[5000, 9])                                 # duplicate label
    # k larger than the store is clamped; an empty store answers with zero columns
    small = store.EmbeddingStore(D, torch.bfloat16, "cuda")
    assert small.search(q.cuda(), 5)[1].shape == (17, 0)
    small.add(x[:3].cuda())
    assert small.search(q.cuda(), 5)[1].shape == (17, 3)


def test_add_tokens_writes_pooled_rows_into_the_tail(mods):
    ops, store = mods
    g = torch.Generator().manual_seed(3)
    B, L, D = 50, 20, 256
    tok = torch.randn(B, L, D, generator=g)
    lens = torch.randint(1, L + 1, (B,), generator=g)
    mask = (torch.arange(L)[None] < lens[:, None]).to(torch.int64)
    st = store.EmbeddingStore(D, torch.bfloat16, "cuda", capacity=4)
    st.add(_unit(7, D, 4).cuda())
    order = torch.randperm(B, generator=g)
    st.add_tokens(tok.cuda(), mask.cuda(), order=order)
    assert len(st) == 57
    exp_rows, exp_inv = O.pool_normalize_cast(tok, mask, torch.bfloat16)
    got = st.rows[7:57].cpu().float()
    np.testing.assert_allclose(got[order].numpy(), exp_rows.float().numpy(), atol=2 ** -8)
    np.testing.assert_allclose(st.inv_norm[7:57].cpu()[order].numpy(), exp_inv.numpy(), rtol=1e-5)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float8_e4m3fn, torch.float32])
def test_save_load_round_trip(mods, tmp_path, dtype):
    ops, store = mods
    D = 64
    st = store.EmbeddingStore(D, dtype, "cuda")
    scale = 64.0 if dtype == torch.float8_e4m3fn else 1.0
    x = (_unit(5000, D, 5) * scale).to(dtype)
    st.add(x.cuda(), ids=range(100, 5100))
    st.remove([100, 2000, 5099])
    st.save(str(tmp_path / "idx"))
    back = store.EmbeddingStore.load(str(tmp_path / "idx"), "cuda", spare=10)
    assert len(back) == len(st) == 4997 and back.capacity >= 5007
    assert torch.equal(back.rows[:4997].view(torch.uint8), st.rows[:4997].view(torch.uint8))
    assert torch.equal(back.inv_norm[:4997], st.inv_norm[:4997]) and torch.equal(back.ids[:4997], st.ids[:4997])
    q = (_unit(9, D, 6) * scale).to(dtype)
    a, b = st.search(q.cuda(), 5), back.search(q.cuda(), 5)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert back.add(x[:1].cuda()).tolist() == [5100]                        # label counter survives


def test_sharded_save_load_equals_whole(mods, tmp_path):
    ops, store = mods
    from text_similarity_b200.sharded import shard_bounds
    D, N, G, k = 64, 4000, 4, 7
    x = _unit(N, D, 8).to(torch.bfloat16)
    q = _unit(11, D, 9).to(torch.bfloat16).cuda()
    for r in range(G):
        b, e = shard_bounds(N, G, r)
        part = store.EmbeddingStore(D, torch.bfloat16, "cuda")
        part.add(x[b:e].cuda(), ids=range(b, e))
        part.save(str(tmp_path / "sh"), world=G, rank=r)
    parts = [store.EmbeddingStore.load(str(tmp_path / "sh"), "cuda", world=G, rank=r) for r in range(G)]
    s64, idx = [], []
    for r, part in enumerate(parts):
        b, _ = shard_bounds(N, G, r)
        _, i, s = ops.search_topk(q, part.rows[:len(part)], k, corpus_inv_norm=part.inv_norm[:len(part)],
                                  idx_base=b, return_score64=True)
        s64.append(s)
        idx.append(i)
    _, _, merged = ops.merge_topk(torch.cat(s64, 1), torch.cat(idx, 1), k, G)
    assert torch.equal(merged.cpu(), O.search_exact(q.cpu(), x, k)[1])
