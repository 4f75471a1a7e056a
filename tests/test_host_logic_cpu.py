"""CPU tests of the host-side logic: the reference-facing classes import and behave without a GPU
up to the point where they would launch a kernel (where they must raise, not fall back), the
tokenizer helper, sharding bounds, and the N > 1 gather/merge orchestration under gloo."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from oracle import oracle as O


def test_reference_module_paths_import():
    from src.configurations.config import Configuration, ModelParameters, SearchConfiguration
    from src.dataset.dataset import EmbeddingsFeatures
    from src.models.sentence_encoder import OnnxSentenceTransformerWrapper, SentenceTransformerWrapper  # noqa: F401
    from src.modules.modules import AvgPoolingStrategy, PoolingStrategy
    from src.pipeline.search_pipeline import Pipeline, SearchPipeline, SentenceMiningPipeline  # noqa: F401
    from src.utils.metrics import cos_sim  # noqa: F401
    cfg = SearchConfiguration(model_parameters=ModelParameters(model_name="m"), model="m", save_path="p")
    assert (cfg.ef, cfg.ef_construction, cfg.M) == (50, 400, 64)       # ints, not 1-tuples (A10)
    assert cfg.batch_size == 16 and cfg.sequence_max_len == 256 and isinstance(cfg, Configuration)
    f = EmbeddingsFeatures(torch.zeros(2, 3, dtype=torch.int64), torch.ones(2, 3, dtype=torch.int64))
    assert set(f.to_dict()) == {"input_ids", "attention_mask"}
    assert set(EmbeddingsFeatures.from_dict({**f.to_dict(), "token_type_ids": f.input_ids}).to_dict()) == \
        {"input_ids", "attention_mask", "token_type_ids"}
    assert issubclass(AvgPoolingStrategy, PoolingStrategy) and len(AvgPoolingStrategy(cfg).state_dict()) == 0


def test_no_cpu_fallback_anywhere_on_the_surface():
    from src.configurations.config import ModelParameters, SearchConfiguration
    from src.modules.modules import AvgPoolingStrategy
    from src.pipeline.search_pipeline import SentenceMiningPipeline
    from src.utils.metrics import cos_sim
    cfg = SearchConfiguration(model_parameters=ModelParameters(model_name="m"), model="m", save_path="p",
                              device=torch.device("cpu"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        AvgPoolingStrategy(cfg)(torch.randn(2, 3, 8), {"attention_mask": torch.ones(2, 3, dtype=torch.int64)})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cos_sim(torch.randn(2, 8), torch.randn(3, 8))
    pipe = SentenceMiningPipeline(100, params=cfg, model=None, name="x")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pipe._search(torch.randn(2, 8), torch.randn(5, 8), 2)
    with pytest.raises(ValueError):
        SentenceMiningPipeline(100, params=cfg, model=None)(torch.randn(2, 8), 2)   # no corpus anywhere


def test_product_package_does_not_import_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for base in ("text_similarity_b200", "src"):
        for dirpath, _, files in os.walk(os.path.join(root, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, fn)).read()
                    assert "import oracle" not in text and "from oracle" not in text, \
                        f"{dirpath}/{fn} imports the oracle"


def test_scripts_compile_and_stay_off_the_oracle():
    """scripts/ are measurement drivers and fuzzers, not tests: they compile, and none of them may use oracle/ (only
    tests/, smoke() and bench.py's CPU legs may) -- the fuzzers check against the library's float64 scan and plain
    torch formulas."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n = 0
    for fn in sorted(os.listdir(os.path.join(root, "scripts"))):
        if fn.endswith(".py"):
            path = os.path.join(root, "scripts", fn)
            text = open(path).read()
            compile(text, path, "exec")
            assert "import oracle" not in text and "from oracle" not in text, f"scripts/{fn} imports the oracle"
            n += 1
    assert n >= 3


def test_synthetic_tokenizer_follows_hf_call_shape():
    from text_similarity_b200.utils import SyntheticTokenizer, synthetic_sentences
    tok = SyntheticTokenizer()
    docs = synthetic_sentences(5, seed=0)
    enc = tok(text=docs, add_special_tokens=True, padding="longest", truncation=True, max_length=16,
              return_attention_mask=True, return_token_type_ids=False, return_tensors="pt")
    ids, mask = enc["input_ids"], enc["attention_mask"]
    assert ids.shape == mask.shape and ids.shape[1] <= 16 and ids.dtype == torch.int64
    assert (ids[:, 0] == 101).all() and ((ids == 0) == (mask == 0)).all()
    assert torch.equal(tok(text=docs, max_length=16)["input_ids"], ids)      # deterministic


def test_shard_bounds_cover_rows_exactly():
    from text_similarity_b200.sharded import shard_bounds
    for n, g in [(10, 3), (10_000_000, 8), (7, 8), (0, 2), (1000, 1)]:
        spans = [shard_bounds(n, g, r) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(e - b for b, e in spans) == (n + g - 1) // g if n else True


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, ret):
    import torch.distributed as dist
    from text_similarity_b200.sharded import gather_shard_results, shard_bounds
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)                       # identical data on every rank
    corpus = torch.randn(1001, 32, generator=g)
    corpus[900] = corpus[5]
    queries = torch.randn(6, 32, generator=g)
    k = 8
    b, e = shard_bounds(corpus.shape[0], world, rank)
    # the per-shard search is the CUDA kernel in production; here the oracle stands in for it so that
    # the exchange + layout logic can run on CPU ranks
    s64, idx = O.search_exact(queries, corpus[b:e], k, idx_base=b)
    all_s, all_i = gather_shard_results(s64, idx, dist.group.WORLD)
    assert all_s.shape == (6, world * k) and all_i.shape == (6, world * k)
    ms, mi = O.merge_topk_exact(all_s, all_i, k)
    fs, fi = O.search_exact(queries, corpus, k)
    ok = torch.equal(mi, fi) and torch.equal(ms, fs)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_two_rank_gather_and_merge_equals_single_search():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gloo_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_store_meter_and_shadow_refuse_the_cpu():
    from src.utils.metrics import RetrievalAccuracyMeter
    from text_similarity_b200 import ops
    from text_similarity_b200.sharded import ShardedCorpus
    from text_similarity_b200.store import EmbeddingStore
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        EmbeddingStore(64, torch.bfloat16, "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        RetrievalAccuracyMeter().update(torch.randn(4, 8), torch.randn(4, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.make_shadow(torch.randn(4, 8))
    sc = ShardedCorpus(torch.randn(16, 8), idx_base=3)       # a CPU shard is only a description (gloo tests)
    assert sc.shadow is None and sc.inv_norm is None and sc.idx_base == 3
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sc.search(torch.randn(2, 8), 3)


def test_abi_shadow_entry_validates_arguments_before_touching_the_gpu():
    from text_similarity_b200 import _lib
    lib = _lib.load()
    assert lib.tsim_search_shadow_workspace_bytes(100, 10_000, 384, 10, _lib.BF16) > 0
    assert lib.tsim_search_shadow_workspace_bytes(100, 10_000, 384, 0, _lib.BF16) == 0
    rc = lib.tsim_search_topk_shadow(None, _lib.F32, 8, None, _lib.F32, 8, None, 8, None, 8, _lib.BF16, None,
                                     4, 10, 8, 3, 0, -1, None, None, None, None, None, 0, None)
    assert rc == _lib.ERR_INVALID_ARG and b"shadow" in lib.tsim_last_error()
    rc = lib.tsim_search_topk_shadow(1, _lib.F32, 8, 1, _lib.F32, 8, 1, 8, 1, 8, _lib.E4M3, None,
                                     4, 10, 8, 3, 0, -1, None, None, None, None, None, 0, None)
    assert rc == _lib.ERR_INVALID_ARG and b"bf16" in lib.tsim_last_error()


def test_bench_reference_arm_contract(monkeypatch):
    """bench.py --impl reference: rank 0 prints ONE JSON line carrying the base contract's keys plus impl /
    cpu_baseline / e2e; other ranks exit 0 without work (the CPU sample itself is shrunk here)."""
    import json
    import subprocess
    import sys as _sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([_sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""

    _sys.path.insert(0, root)
    import bench
    monkeypatch.setattr(bench, "_cpu_sample", lambda N, D, n_rows, n_q: (torch.randn(8, D), torch.randn(2000, D)))
    qps, cores, sample, per_step, done = bench.cpu_reference_steps(20_000, 64, 4096, 10, steps=3, warmup=1, budget_s=5.0)
    assert qps > 0 and cores >= 1 and "scaled linearly" in sample and done == 3 and per_step > 0
    loop = bench.cpu_loop_variant_i(20_000, 64, 10, budget_s=0.2)
    assert loop["value"] > 0 and loop["kind"] == "port" and "per-query loop" in loop["what"]

    monkeypatch.setattr(bench, "cpu_reference_steps",
                        lambda N, D, Q, k, steps, warmup, budget_s: (12.5, 4, "stub sample", 0.1, steps))
    lines = []
    monkeypatch.setattr(bench, "_emit", lambda fd, line: lines.append(line))
    args = type("A", (), {"workload": bench.DEFAULT_WORKLOAD, "steps": 2, "warmup": 1, "gpus": 1})()
    monkeypatch.setenv("RANK", "0")
    bench.run_reference(args, 1)
    (line,) = lines
    json.dumps(line)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"):
        assert key in line, key
    assert line["impl"] == "reference" and line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["config"]["workload"] == bench.DEFAULT_WORKLOAD and line["vs_baseline"] is None
    # both arms describe the workload with the SAME config object, and a step of the CPU arm is the bounded sample
    assert line["config"] == bench.shared_config(bench.DEFAULT_WORKLOAD, 1)
    assert line["steps"] == 2 and line["ms_per_step"] == pytest.approx(100.0)


def test_bench_torch_merge_reference_matches_oracle():
    """bench.py checks the merge kernel + all-gather layout against an independent torch merge: that merge itself
    must agree with the oracle's (score desc, index asc, padding last)."""
    import sys as _sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    _sys.path.insert(0, root)
    import bench
    from oracle import oracle as O
    g = torch.Generator().manual_seed(3)
    world, n, k = 4, 50, 7
    s = torch.randint(0, 6, (world, n, k), generator=g).double() / 4        # many ties
    i = torch.stack([torch.stack([torch.randperm(100, generator=g)[:k] + 100 * w for _ in range(n)]) for w in range(world)])
    i[1, :, 5:] = -1                                                          # padding
    # per-shard lists arrive best first
    for w in range(world):
        key = torch.where(i[w] < 0, torch.full_like(s[w], -1e9), s[w])
        o = torch.argsort(key, dim=1, descending=True, stable=True)
        s[w], i[w] = torch.gather(s[w], 1, o), torch.gather(i[w], 1, o)
    ms, mi = bench.torch_merge_reference(s, i, k)
    es, ei = O.merge_topk_exact(s.permute(1, 0, 2).reshape(n, -1), i.permute(1, 0, 2).reshape(n, -1), k)
    assert torch.equal(mi, ei) and torch.equal(ms[mi >= 0], es[ei >= 0])


def test_raw_maximum_bound_of_the_epilogue_filter_is_conservative():
    """search_tc.cu's hot path skips a 32-column chunk when  max_j(raw_j) * (max >= 0 ? inv_hi : inv_lo) <= thr.
    Property restated in float32 on the host: that product is never below any scaled score fl(raw_j * inv_j),
    whatever the signs (float rounding is monotone), so a skipped chunk cannot hold a candidate."""
    import numpy as np
    rng = np.random.default_rng(7)
    for trial in range(2000):
        scale = np.float32(10.0 ** rng.uniform(-3, 3))
        raw = (rng.standard_normal(32) * scale).astype(np.float32)
        if trial % 3 == 0:
            raw = -np.abs(raw)                      # an all-negative chunk takes the inv_lo branch
        inv = (1.0 + rng.uniform(-0.3, 0.3, 32)).astype(np.float32) * np.float32(10.0 ** rng.uniform(-2, 2))
        if trial % 7 == 0:
            inv[rng.integers(0, 32)] = np.float32(0.0)   # out-of-range column of a ragged last tile
        scaled = raw * inv                           # float32 products, as the exact path computes them
        mx = raw.max()
        bound = mx * (inv.max() if mx >= 0 else inv.min())
        assert bound >= scaled.max(), (trial, bound, scaled.max())


def test_excluded_reference_classes_are_importable(tmp_path):
    """the reference's serving / clustering / evaluation entry points import from their
    own module paths and keep their signatures; what needs a missing dependency says so."""
    from src.configurations.config import ModelParameters, SearchConfiguration
    from src.evaluation.eval_sentence_mining import compare_models
    from src.pipeline.clustering import ClusteringPipeline
    from src.pipeline.search_pipeline import APISearchPipeline, SemanticSearchPipeline
    assert issubclass(APISearchPipeline, SemanticSearchPipeline)
    params = SearchConfiguration(model_parameters=ModelParameters(model_name="m", hidden_size=8), model="m",
                                 save_path=str(tmp_path), tokenizer=None, device=torch.device("cpu"))
    pipe = APISearchPipeline(params, 5, str(tmp_path / "index"), model=None)
    assert pipe.max_n_results == 5 and pipe.session is None and pipe.num_indexed() == 0
    params.model_path = str(tmp_path / "model.onnx")
    try:
        import onnxruntime  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="onnxruntime"):
            APISearchPipeline(params, 5, str(tmp_path / "index"), model=None)
    clu = ClusteringPipeline(3, params, None)
    clu.set_n_clusters(4)
    assert clu.n_clusters == 4 and clu.method == "k-means"
    with pytest.raises(ValueError):
        ClusteringPipeline(3, params, None, method="dbscan")
    teacher = {0: [(1, "a"), (2, "b")], 1: [(5, "c"), (6, "d")]}
    student = {0: [(1, "a"), (9, "z")], 1: [(6, "d"), (5, "c")]}
    assert compare_models(["q0", "q1"], teacher, student) == pytest.approx(75.0)


def test_token_budget_bucketing_covers_every_sentence_once():
    """Length-bucketed batching (encoder._token_budget_batches): every sentence appears exactly once, with the same
    token ids a per-sentence tokenizer call gives, padded size within the budget, little padding."""
    from text_similarity_b200.config import ModelParameters, SearchConfiguration
    from text_similarity_b200.encoder import SentenceTransformerWrapper
    from text_similarity_b200.utils import SyntheticTokenizer, synthetic_sentences
    tok = SyntheticTokenizer()
    params = SearchConfiguration(model_parameters=ModelParameters(model_name="m", hidden_size=8), model="m", save_path=".",
                                 tokenizer=tok, sequence_max_len=32, batch_size=16, device=torch.device("cpu"),
                                 token_budget=256)

    class Dummy(torch.nn.Module):
        config = type("c", (), {"hidden_size": 8})()

    m = SentenceTransformerWrapper(pooler=None, merge_strategy=None, loss=None, params=params,
                                   context_embedder=Dummy(), parallel_mode=False)
    docs = synthetic_sentences(200, seed=3)
    seen, cells, pad = [], 0, 0
    for rows, feats in m._batches(docs):
        ids, mask = feats.input_ids, feats.attention_mask
        assert ids.shape[0] * ids.shape[1] <= 256 or ids.shape[0] == 1
        for r, i in enumerate(rows.tolist()):
            ref = tok([docs[i]], max_length=32)["input_ids"][0]
            n = int(mask[r].sum())
            assert ids[r, :n].tolist() == ref.tolist() and (ids[r, n:] == 0).all()
        seen += rows.tolist()
        cells += ids.numel()
        pad += int((mask == 0).sum())
    assert sorted(seen) == list(range(200))
    assert pad / cells < 0.1
    params.token_budget = None                      # default: the reference's fixed batches of 16
    assert [len(r) for r, _ in m._batches(docs)] == [16] * 12 + [8]
