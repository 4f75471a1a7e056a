"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference (cr1m5onk1ng/text_similarity) imports third-party packages that are not
installed here (nltk, sentence_transformers, hnswlib, onnxruntime); none of them
touches the arithmetic of the hot path, so they are replaced by empty stub modules.
With the stubs in place the following reference objects are imported and executed
unmodified:

* ``src.modules.modules.AvgPoolingStrategy.forward``            (modules.py:158-171)
* ``src.models.sentence_encoder.OnnxSentenceTransformerWrapper.forward`` (sentence_encoder.py:32-39)
* ``src.utils.metrics.cos_sim``                                 (metrics.py:81-101)
* ``src.dataset.dataset.EmbeddingsFeatures``                    (dataset.py:213-251)

``SentenceMiningPipeline._search`` raises as checked in, so for the
search itself the script issues the same two ATen calls the method makes
(``F.cosine_similarity`` on the expanded query, search_pipeline.py:76-77, and
``torch.topk(..., largest=True, sorted=False)``, :78) with the documented repairs
(dim of the 1-D top-k, k clamped by the corpus size).

Outputs are small (< 1 MB total) and committed; the GPU box never needs /root/reference.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


class _StubModule(types.ModuleType):
    """Empty stand-in: any attribute is a dummy class, any submodule another stub."""

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return type(item, (), {})


STUB_ROOTS = ("nltk", "sentence_transformers", "hnswlib", "onnxruntime", "matplotlib", "seaborn")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Resolves any module under the given (not installed) roots to an empty _StubModule."""

    def __init__(self, roots):
        self.roots = tuple(roots)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def import_reference():
    missing = []
    for n in STUB_ROOTS:
        try:
            __import__(n)
        except Exception:
            missing.append(n)
    if missing:
        sys.meta_path.insert(0, _StubFinder(missing))
    sys.path.insert(0, REF)
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    import src.modules.modules as ref_modules
    import src.utils.metrics as ref_metrics
    import src.models.sentence_encoder as ref_encoder
    import src.dataset.dataset as ref_dataset
    assert ref_modules.__file__.startswith(REF), ref_modules.__file__
    print("stubbed (not installed here):", missing)
    return ref_modules, ref_metrics, ref_encoder, ref_dataset


def main():
    ref_modules, ref_metrics, ref_encoder, ref_dataset = import_reference()
    torch.manual_seed(0)
    torch.set_num_threads(1)
    out = {}

    # ---- pooling: AvgPoolingStrategy.forward on ragged masks (incl. an all-masked row)
    pooler = ref_modules.AvgPoolingStrategy(params=None)
    for tag, (B, L, D) in {"small": (5, 9, 8), "minilm": (16, 21, 384), "wide": (3, 40, 768)}.items():
        g = torch.Generator().manual_seed(100 + B)
        emb = torch.randn(B, L, D, generator=g)
        lens = torch.randint(1, L + 1, (B,), generator=g)
        lens[0] = L
        if B > 2:
            lens[2] = 0  # all-masked row -> zero vector (clamp 1e-9)
        mask = (torch.arange(L)[None, :] < lens[:, None]).to(torch.int64)
        feats = ref_dataset.EmbeddingsFeatures(input_ids=torch.zeros(B, L, dtype=torch.int64),
                                               attention_mask=mask)
        pooled = pooler(emb, feats)
        out[f"pool_{tag}_emb"] = emb.numpy()
        out[f"pool_{tag}_mask"] = mask.numpy()
        out[f"pool_{tag}_out"] = pooled.numpy()

    # ---- OnnxSentenceTransformerWrapper.forward on a tiny random-init BERT
    from transformers import BertConfig, BertModel
    torch.manual_seed(0)
    cfg = BertConfig(vocab_size=97, hidden_size=32, num_hidden_layers=2, num_attention_heads=4,
                     intermediate_size=64, max_position_embeddings=64)
    bert = BertModel(cfg).eval()
    wrapper = ref_encoder.OnnxSentenceTransformerWrapper(params=None, context_embedder=bert).eval()
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(1, 97, (6, 12), generator=g)
    lens = torch.tensor([12, 3, 7, 1, 12, 9])
    mask = (torch.arange(12)[None, :] < lens[:, None]).to(torch.int64)
    with torch.no_grad():
        tok = bert(input_ids=ids, attention_mask=mask)[0]
        pooled = wrapper(ids, mask)
    out["onnx_tok"] = tok.numpy()
    out["onnx_mask"] = mask.numpy()
    out["onnx_out"] = pooled.numpy()

    # ---- cos_sim (metrics.py:81-101)
    g = torch.Generator().manual_seed(11)
    a = torch.randn(5, 16, generator=g)
    b = torch.randn(7, 16, generator=g) * 3.0
    out["cossim_a"] = a.numpy()
    out["cossim_b"] = b.numpy()
    out["cossim_out"] = ref_metrics.cos_sim(a, b).numpy()
    out["cossim_1d_out"] = ref_metrics.cos_sim(a[0], b).numpy()
    out["cossim_list_out"] = ref_metrics.cos_sim(a.tolist(), b.numpy()).numpy()

    # ---- the search's ATen calls (search_pipeline.py:76-78), with planted duplicates -> ties
    for tag, (Q, N, D, k) in {"tiny": (6, 50, 16, 5), "mid": (12, 1000, 64, 10)}.items():
        g = torch.Generator().manual_seed(21 + N)
        corpus = torch.randn(N, D, generator=g)
        queries = torch.randn(Q, D, generator=g)
        corpus[N // 2] = corpus[3]          # exact duplicate rows
        corpus[N - 1] = corpus[3]
        corpus[7] = queries[0] * 2.5        # an exact hit for query 0 (cos == 1)
        corpus[11] = 0.0                    # zero row -> cosine 0
        scores = torch.stack([F.cosine_similarity(q.unsqueeze(0).expand_as(corpus), corpus, dim=-1)
                              for q in queries])
        tv, ti = [], []
        for s in scores:
            top = torch.topk(s, min(k, N), dim=0, sorted=False, largest=True)
            tv.append(top[0])
            ti.append(top[1])
        out[f"search_{tag}_corpus"] = corpus.numpy()
        out[f"search_{tag}_queries"] = queries.numpy()
        out[f"search_{tag}_k"] = np.int64(k)
        out[f"search_{tag}_scores"] = scores.numpy()
        out[f"search_{tag}_topk_val"] = torch.stack(tv).numpy()
        out[f"search_{tag}_topk_idx"] = torch.stack(ti).numpy()

    path = os.path.join(HERE, "reference_outputs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})
    print("torch", torch.__version__)


if __name__ == "__main__":
    main()
