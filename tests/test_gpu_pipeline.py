"""GPU tests of the reference-facing surface: the same imports and calls a script written against
cr1m5onk1ng/text_similarity makes, checked against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def stack():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from src.configurations.config import ModelParameters, SearchConfiguration
    from src.models.sentence_encoder import SentenceTransformerWrapper
    from src.modules.modules import AvgPoolingStrategy
    from text_similarity_b200.utils import SyntheticTokenizer, minilm_l6_encoder, synthetic_sentences
    params = SearchConfiguration(model_parameters=ModelParameters(model_name="minilm-l6-shaped", hidden_size=384),
                                 model="synthetic-minilm", save_path="./results", tokenizer=SyntheticTokenizer(),
                                 sequence_max_len=64, batch_size=16, device=torch.device("cuda"))
    bert = minilm_l6_encoder(seed=0, layers=2)
    model = SentenceTransformerWrapper(pooler=AvgPoolingStrategy(params), merge_strategy=None, loss=None,
                                       params=params, context_embedder=bert, parallel_mode=False)
    corpus = synthetic_sentences(300, seed=1)
    corpus[200] = corpus[13]            # duplicate sentences -> exact ties
    queries = synthetic_sentences(9, seed=2) + [corpus[50]]
    return params, model, corpus, queries


def test_encode_text_matches_per_sentence_torch(stack):
    params, model, corpus, _ = stack
    docs = corpus[:40]
    emb = model.encode_text(docs)
    assert emb.shape == (40, 384) and emb.dtype == torch.float32 and emb.is_cuda
    assert model.get_sentence_embedding_dimension() == 384
    tok = params.tokenizer
    with torch.no_grad():
        for i in (0, 7, 39):
            enc = tok(text=[docs[i]], max_length=params.sequence_max_len)
            t = model.context_embedder(input_ids=enc["input_ids"].cuda(), attention_mask=enc["attention_mask"].cuda())[0]
            exp = O.mean_pool_literal(t.cpu(), enc["attention_mask"])
            np.testing.assert_allclose(emb[i].cpu().numpy(), exp[0].numpy(), atol=2e-5)
    as_np = model.encode(docs[:5], output_np=True)
    assert isinstance(as_np, np.ndarray) and as_np.shape == (5, 384)
    np.testing.assert_allclose(as_np, emb[:5].cpu().numpy(), atol=2e-5)


def test_pooler_and_onnx_wrapper_follow_reference(stack, golden):
    from src.dataset.dataset import EmbeddingsFeatures
    from src.models.sentence_encoder import OnnxSentenceTransformerWrapper
    from src.modules.modules import AvgPoolingStrategy
    params, model, _, _ = stack
    pooler = AvgPoolingStrategy(params)
    assert len(pooler.state_dict()) == 0
    emb = torch.from_numpy(golden["pool_minilm_emb"]).cuda()
    mask = torch.from_numpy(golden["pool_minilm_mask"]).cuda()
    feats = EmbeddingsFeatures(input_ids=torch.zeros_like(mask), attention_mask=mask)
    np.testing.assert_allclose(pooler(emb, feats).cpu().numpy(), golden["pool_minilm_out"], atol=1e-5)
    with pytest.raises(AssertionError):
        pooler(emb[0], feats)
    onnx = OnnxSentenceTransformerWrapper(params=params, context_embedder=model.context_embedder).cuda().eval()
    enc = params.tokenizer(text=["w1 w2 w3", "w4"], max_length=32)
    with torch.no_grad():
        got = onnx(enc["input_ids"].cuda(), enc["attention_mask"].cuda())
        tok = model.context_embedder(input_ids=enc["input_ids"].cuda(), attention_mask=enc["attention_mask"].cuda())[0]
    np.testing.assert_allclose(got.cpu().numpy(), O.mean_pool_literal(tok.cpu(), enc["attention_mask"]).numpy(), atol=2e-5)


def test_sentence_mining_pipeline_text(stack):
    from src.pipeline.search_pipeline import SentenceMiningPipeline
    params, model, corpus, queries = stack
    pipe = SentenceMiningPipeline(10_000, params=params, model=model, corpus=corpus, name="teacher")
    res = pipe(queries, 5)                       # the call the reference's eval scripts make
    assert sorted(res) == list(range(len(queries)))
    rows, inv = model.encode_text_normalized(corpus, torch.bfloat16)
    qrows, _ = model.encode_text_normalized(queries, torch.bfloat16)
    ev, ei = O.search_exact(qrows.cpu(), rows.cpu(), 5)
    for qi in range(len(queries)):
        assert [c for c, _ in res[qi]] == ei[qi].tolist()
        assert all(text == corpus[c] for c, text in res[qi])
    assert res[len(queries) - 1][0][0] == 50      # the query copied from corpus[50] finds it first
    # stored rows are unit norm up to bf16 rounding and inv_norm describes the stored rows
    n = rows.float().norm(dim=-1).cpu()
    assert (n - 1).abs().max() < 1e-2
    np.testing.assert_allclose((inv.cpu() * n).numpy(), 1.0, atol=1e-5)
    # chunked search (several chunks + merge) returns the same thing
    pipe_small = SentenceMiningPipeline(37, params=params, model=model, corpus=corpus, name="chunked")
    res2 = pipe_small._search(queries, None, 5)
    assert {q: [c for c, _ in v] for q, v in res2.items()} == {q: [c for c, _ in v] for q, v in res.items()}


def test_sentence_mining_pipeline_tensors(stack):
    from src.pipeline.search_pipeline import SentenceMiningPipeline
    params, model, _, _ = stack
    g = torch.Generator().manual_seed(3)
    corpus = torch.randn(5000, 384, generator=g)
    queries = torch.randn(12, 384, generator=g)
    pipe = SentenceMiningPipeline(2048, params=params, model=model)
    # fp32 tensors: answered by the float64 exact scan, 1e-5 tolerance
    out = pipe._search(queries.cuda(), corpus.cuda(), 7, return_embeddings=True)
    ev, ei = O.search_exact(queries, corpus, 7)
    for qi in range(12):
        assert torch.equal(out[qi].cpu(), corpus[ei[qi]])
    s, i = pipe.search_tensors(queries.cuda(), 7)
    assert torch.equal(i.cpu(), ei)
    np.testing.assert_allclose(s.cpu().numpy(), ev.numpy(), atol=1e-5)
    # fp32 tensors, 24 < k <= 100: the split (hi + lo) bf16 shadow, made on first use, chunks merged -- same answer
    s50, i50 = pipe.search_tensors(queries.cuda(), 50)
    ev50, ei50 = O.search_exact(queries, corpus, 50)
    assert torch.equal(i50.cpu(), ei50)
    np.testing.assert_allclose(s50.cpu().numpy(), ev50.numpy(), atol=1e-5)
    # fp32 rows whose width is no multiple of 8 (no rounded shadow): the split shadow pads its segments, any k <= 100
    odd = torch.randn(6000, 100, generator=g)
    qo = torch.randn(9, 100, generator=g)
    so, io = pipe.search_tensors(qo.cuda(), 10, corpus=odd.cuda())
    evo, eio = O.search_exact(qo, odd, 10)
    assert torch.equal(io.cpu(), eio)
    np.testing.assert_allclose(so.cpu().numpy(), evo.numpy(), atol=1e-5)
    # bf16 tensors: tcgen05 path, k larger than the corpus is clamped (reference clamps by #queries, A6)
    cb = corpus[:9].to(torch.bfloat16)
    s, i = pipe.search_tensors(queries.to(torch.bfloat16).cuda(), 100, corpus=cb.cuda())
    assert i.shape == (12, 9)
    assert torch.equal(i.cpu(), O.search_exact(queries.to(torch.bfloat16), cb, 9)[1])


def test_semantic_search_pipeline_surface(stack, tmp_path):
    from src.pipeline.search_pipeline import SemanticSearchPipeline
    params, model, corpus, queries = stack
    pipe = SemanticSearchPipeline(str(tmp_path / "index"), params=params, model=model, corpus=list(corpus[:120]))
    hits = pipe(queries, 3)
    assert all(len(v) == 3 and all(isinstance(t, str) for t in v) for v in hits.values())
    assert hits[len(queries) - 1][0] == corpus[50]
    assert pipe.num_indexed() == 120
    pipe.add_to_index(["w1 w2 w3 w4", queries[0]])
    assert pipe.num_indexed() == 122 and pipe(queries[:1], 1)[0] == [queries[0]]
    pipe.remove_from_index([121, 9999])
    assert pipe.num_indexed() == 121 and pipe(queries[:1], 1)[0] != [queries[0]]
    again = SemanticSearchPipeline(str(tmp_path / "index"), params=params, model=model, corpus=list(corpus[:120]))
    assert again(queries, 3) == hits                     # reloaded from index_path


def test_cos_sim_surface(golden):
    from src.utils.metrics import cos_sim, cos_sim_topk
    a = torch.from_numpy(golden["cossim_a"])
    b = torch.from_numpy(golden["cossim_b"])
    np.testing.assert_allclose(cos_sim(a.cuda(), b.cuda()).cpu().numpy(), golden["cossim_out"], atol=1e-5)
    np.testing.assert_allclose(cos_sim(a[0].cuda(), b.cuda()).cpu().numpy(), golden["cossim_1d_out"], atol=1e-5)
    s, i = cos_sim_topk(a.cuda(), b.cuda(), 3)
    assert torch.equal(i.cpu(), O.search_exact(a, b, 3)[1])
    with pytest.raises(RuntimeError):
        cos_sim(a, b)


def test_ranking_pipeline_and_helpers(stack):
    from src.pipeline.ranking_pipeline import RankingPipeline
    from src.utils.utils import most_similar_vectors
    from text_similarity_b200.ranking import assign_to_centroids, near_duplicates
    params, model, corpus, queries = stack

    class LengthCrossEncoder:                       # stand-in cross-encoder: prefers short passages
        def predict(self, pairs):
            return [1.0 / (1 + len(p[1])) for p in pairs]

    pipe = RankingPipeline(LengthCrossEncoder(), 10_000, params=params, model=model, name="rank")
    out = pipe(queries[:3], corpus[:100], top_k=5)
    assert len(out) == 3
    for res in out:
        hits = res["results"]
        assert len(hits) == 5 and all("cross-score" in h and h["text"] == corpus[h["corpus_id"]] for h in hits)
        assert [h["cross-score"] for h in hits] == sorted((h["cross-score"] for h in hits), reverse=True)
        assert abs(res["avg_score"] - sum(res["cross_scores"]) / 5) < 1e-12
    # retrieval part == the exact search
    rows, _ = model.encode_text_normalized(corpus[:100], torch.bfloat16)
    qrows, _ = model.encode_text_normalized(queries[:3], torch.bfloat16)
    _, ei = O.search_exact(qrows.cpu(), rows.cpu(), 5)
    assert [sorted(h["corpus_id"] for h in r["results"]) for r in out] == [sorted(x) for x in ei.tolist()]

    g = torch.Generator().manual_seed(9)
    vecs = torch.randn(500, 64, generator=g)
    best = most_similar_vectors(vecs[17].cuda(), vecs.cuda(), n=3)
    _, ei = O.search_exact(vecs[17:18], vecs, 3)
    assert [torch.equal(b.cpu(), vecs[i]) for b, i in zip(best, ei[0].tolist())] == [True] * 3

    cent = torch.randn(7, 64, generator=g)
    lab = assign_to_centroids(vecs.cuda(), cent.cuda())
    assert torch.equal(lab.cpu(), O.cosine_scores_exact(vecs, cent).argmax(1))

    x = torch.nn.functional.normalize(torch.randn(3000, 128, generator=g), dim=-1).to(torch.bfloat16)
    x[2000] = x[5]
    x[2500] = x[5]
    dup = near_duplicates(x.cuda(), threshold=0.99, k=3)
    assert dup == {5: [2000, 2500], 2000: [5, 2500], 2500: [5, 2000]}


def test_cuda_graph_replay_equals_eager_search():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from text_similarity_b200.sharded import ShardedCorpus
    g = torch.Generator(device="cuda").manual_seed(11)
    c = torch.randn(1_300_000, 64, generator=g, device="cuda")
    c = (c / c.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
    corpus = ShardedCorpus(c, idx_base=1000)
    for Q in (1, 24):                                   # no bootstrap / bootstrap + ladder plans
        for rep in range(3):                            # first call captures, the others replay
            q = torch.randn(Q, 64, generator=g, device="cuda")
            q = (q / q.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
            es, ei = corpus.search(q, 10)
            gs, gi = corpus.search_graphed(q, 10)
            assert torch.equal(gi, ei) and torch.equal(gs, es)
    hs = torch.empty(24, 10, dtype=torch.float32).pin_memory()
    hi = torch.empty(24, 10, dtype=torch.int64).pin_memory()
    corpus.search_host(q.cpu().pin_memory(), 10, hs, hi, graphed=True)
    torch.cuda.synchronize()
    assert torch.equal(hi, ei.cpu()) and torch.equal(hs, es.cpu())


def test_retrieval_accuracy_meter_matches_dense_argmax():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from src.utils.metrics import RetrievalAccuracyMeter
    g = torch.Generator().manual_seed(5)
    src = torch.randn(700, 96, generator=g)
    tgt = src + 2.5 * torch.randn(700, 96, generator=g)        # noisy "translations": ~1 in 4 retrievals fails
    meter = RetrievalAccuracyMeter(print_wrong_matches=True)
    meter.update(src.cuda(), tgt.cuda(), [f"s{i}" for i in range(700)], [f"t{i}" for i in range(700)])
    dense = O.cosine_scores_exact(src, tgt)                      # what the reference's cos_sim + np.argmax computes
    exp_fwd = (dense.argmax(1) == torch.arange(700)).double().mean().item()
    exp_bwd = (dense.argmax(0) == torch.arange(700)).double().mean().item()
    assert 0.2 < exp_fwd < 1.0
    assert abs(meter.src2tgt - exp_fwd) < 1e-6 and abs(meter.tgt2src - exp_bwd) < 1e-6
    assert abs(meter.avg - (exp_fwd + exp_bwd) / 2) < 1e-6
    assert len(meter.lines) == int(round((1 - exp_fwd) * 700)) and "INCORRECT" in str(meter)


def test_clustering_pipeline_separates_planted_clusters(stack):
    """ClusteringPipeline (reference src/pipeline/clustering.py:8-31) on the search kernel: three well separated
    directions must come back as three pure clusters; labels equal a float64 nearest-centroid assignment."""
    from src.pipeline.clustering import ClusteringPipeline
    params, model = stack[0], stack[1]
    g = torch.Generator().manual_seed(4)
    dirs = torch.nn.functional.normalize(torch.randn(3, 64, generator=g), dim=-1)
    which = torch.randint(0, 3, (600,), generator=g)
    x = torch.nn.functional.normalize(dirs[which] + 0.05 * torch.randn(600, 64, generator=g), dim=-1)
    clu = ClusteringPipeline(3, params, model, seed=1)
    out = clu(x.cuda(), 3)
    assert sorted(len(v) for v in out.values()) == sorted(torch.bincount(which, minlength=3).tolist())
    lab = clu.labels_.cpu()
    for c in range(3):
        assert len(set(which[lab == c].tolist())) == 1          # pure clusters
    want = (x.double() @ clu.cluster_centers_.cpu().double().T).argmax(dim=1)
    assert torch.equal(lab, want)


def test_bf16_autocast_bucketed_encode_matches_fp32_loop(stack):
    """f2: encoder under torch.autocast(bf16) + token-budget batches.  Pooled unit rows stay
    within 2^-7 of the reference-shaped fp32 loop, and the search over the rows it stored is exact (== oracle on
    those rows)."""
    from text_similarity_b200 import ops
    params, model, corpus, queries = stack
    rows_ref, _ = model.encode_text_normalized(corpus, torch.bfloat16)
    params.encode_dtype, params.token_budget = torch.bfloat16, 2048
    try:
        rows, inv = model.encode_text_normalized(corpus, torch.bfloat16)
        q, _ = model.encode_text_normalized(queries, torch.bfloat16)
    finally:
        params.encode_dtype, params.token_budget = None, None
    assert (rows.float() - rows_ref.float()).abs().max().item() <= 2 ** -7      # tolerance stated by the task
    s, i, s64 = ops.search_topk(q, rows, 10, corpus_inv_norm=inv, return_score64=True)
    ev, ei = O.search_exact(q.cpu(), rows.cpu(), 10)
    assert torch.equal(i.cpu(), ei)
    np.testing.assert_allclose(s64.cpu().numpy(), ev.numpy(), atol=1e-12)
