"""Time-boxed randomized sweep of the Python layers above the C ABI on one B200, each against one float64 exact scan
of the whole matrix (Gaussian rows: no ties, so the expected order is unique):

 * EmbeddingStore -- random add / remove / search sequences against a host-side model of the live rows
   (labels must follow their rows through swap-removes and capacity growth);
 * SentenceMiningPipeline.search_tensors -- random corpus_chunk_size (chunk loop + hierarchical K3 merge);
 * contiguous fake shards (shard_bounds: ragged and empty shards, k above a shard's rows) + ops.merge_topk.

    python scripts/fuzz_layers.py --seconds 60 --seed 1

Exit status 1 on any mismatch."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_similarity_b200 import ops  # noqa: E402
from text_similarity_b200.config import ModelParameters, SearchConfiguration  # noqa: E402
from text_similarity_b200.pipeline import SentenceMiningPipeline  # noqa: E402
from text_similarity_b200.sharded import shard_bounds  # noqa: E402
from text_similarity_b200.store import EmbeddingStore  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=60.0)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--max-cases", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda", 0)


def exact(q, c, k):
    s, i, s64 = ops.search_topk(q, c, k, mode="exact", return_score64=True)
    return s, i, s64


def case_store(rng, g):
    D = int(rng.choice([32, 64, 384]))
    dtype = [torch.bfloat16, torch.float32, torch.float8_e4m3fn][int(rng.integers(0, 3))]
    st = EmbeddingStore(D, dtype=dtype, device=dev, capacity=int(rng.choice([1, 8, 64])))
    model = {}                                   # label -> row (store dtype, on the GPU)
    bad = 0
    for step in range(int(rng.integers(3, 12))):
        op = int(rng.integers(0, 4))
        if op <= 1 or not model:
            m = int(rng.integers(1, 400))
            rows = torch.randn(m, D, generator=g, device=dev)
            if dtype == torch.float8_e4m3fn:
                rows = rows * 4
            ids = None
            if rng.integers(0, 3) == 0:
                base = int(rng.integers(10_000, 1_000_000)) * 1000 + step * 100_000_000
                ids = [base + j for j in range(m)]
            lab = st.add(rows, ids=ids)
            stored = rows.to(dtype)
            for j, l in enumerate(lab.tolist()):
                model[int(l)] = stored[j]
        elif op == 2:
            labs = list(model)
            take = [labs[int(x)] for x in rng.integers(0, len(labs), size=int(rng.integers(1, max(2, len(labs) // 3))))]
            take += [-5, 10 ** 15]               # unknown labels are skipped
            done = st.remove(take)
            bad += int(done != len(set(take) & set(model)))
            for l in set(take):
                model.pop(l, None)
        bad += int(len(st) != len(model))
        if model:
            labs = list(model)
            M = torch.stack([model[l].view(torch.uint8) for l in labs]).view(dtype)   # (float8 tensors do not stack)
            Q = int(rng.integers(1, 40))
            q = torch.randn(Q, D, generator=g, device=dev).to(dtype)
            k = int(rng.choice([1, 5, 10, 30]))
            s, got = st.search(q, k)
            kk = min(k, len(labs))
            es, ei, _ = exact(q, M, kk)
            want = torch.tensor(labs, device=dev)[ei]
            torch.cuda.synchronize()
            bad += int(got.shape != (Q, kk)) or int((got != want).sum()) + int((s != es).sum())
    return bad, dict(kind="store", D=D, dtype=str(dtype)[6:], rows=len(model))


_PARAMS = None


def case_pipeline(rng, g):
    global _PARAMS
    if _PARAMS is None:
        _PARAMS = SearchConfiguration(model_parameters=ModelParameters(model_name="none", hidden_size=64), model="none",
                                      save_path="./results", tokenizer=None, sequence_max_len=64, batch_size=16,
                                      device=dev)
    dtype = [torch.bfloat16, torch.float32][int(rng.integers(0, 2))]
    D = int(rng.choice([64, 128, 384]))
    N = int(rng.integers(10, 30_000))
    Q = int(rng.choice([1, 9, 40, 200]))
    k = int(rng.choice([1, 10, 50, 100]))
    chunk = int(rng.choice([7, 100, 1000, 4096, 50_000]))
    if N / chunk > 300:
        chunk = N // 300 + 1
    c = torch.randn(N, D, generator=g, device=dev).to(dtype)
    q = torch.randn(Q, D, generator=g, device=dev).to(dtype)
    pipe = SentenceMiningPipeline(chunk, params=_PARAMS, model=None)
    s, i = pipe.search_tensors(q, k, corpus=c)
    kk = min(k, N)
    es, ei, _ = exact(q, c, kk)
    torch.cuda.synchronize()
    bad = int(i.shape != (Q, kk)) or int((i != ei).sum()) + int((s != es).sum())
    return bad, dict(kind="pipeline", dtype=str(dtype)[6:], N=N, Q=Q, D=D, k=k, chunk=chunk)


def case_shards(rng, g):
    dtype = [torch.bfloat16, torch.float8_e4m3fn][int(rng.integers(0, 2))]
    D = int(rng.choice([64, 384]))
    N = int(rng.choice([1, 3, 7, 50, 1000, 40_000, 400_000]))
    G = int(rng.choice([2, 3, 4, 8]))
    Q = int(rng.choice([1, 8, 32, 100, 300]))
    k = int(rng.choice([1, 10, 100]))
    c = (torch.randn(N, D, generator=g, device=dev) * (4 if dtype == torch.float8_e4m3fn else 1)).to(dtype)
    q = (torch.randn(Q, D, generator=g, device=dev) * (4 if dtype == torch.float8_e4m3fn else 1)).to(dtype)
    s64s, idxs = [], []
    for r in range(G):
        b, e = shard_bounds(N, G, r)
        _, ix, s64 = ops.search_topk(q, c[b:e], k, idx_base=b, return_score64=True)
        s64s.append(s64)
        idxs.append(ix)
    ms, ms64, mi = ops.merge_topk(torch.cat(s64s, 1), torch.cat(idxs, 1), k, G)
    es, ei, es64 = exact(q, c, k)
    torch.cuda.synchronize()
    bad = int((mi != ei).sum()) + int((ms64 != es64).sum()) + int((ms != es).sum())
    return bad, dict(kind="shards", dtype=str(dtype)[6:], N=N, G=G, Q=Q, D=D, k=k)


KINDS = [case_store, case_pipeline, case_shards]
t0 = time.time()
n_cases = fails = 0
case = 0
while time.time() - t0 < a.seconds:
    rng = np.random.default_rng([a.seed, case])
    g = torch.Generator(device=dev).manual_seed(int(rng.integers(0, 2 ** 31)))
    fn = KINDS[case % len(KINDS)]
    try:
        bad, desc = fn(rng, g)
    except Exception as e:  # noqa: BLE001
        print(f"ERROR case {case} ({fn.__name__}): {type(e).__name__}: {e}", flush=True)
        bad, desc = -1, {"kind": fn.__name__}
        if "CUDA" in str(e) or "cuda" in str(e):
            fails += 1
            break
    n_cases += 1
    if bad:
        fails += 1
        print(f"MISMATCH bad={bad} case={case} {desc}", flush=True)
    elif n_cases % 20 == 0:
        print(f"ok case={case} {desc}", flush=True)
    case += 1
    if a.max_cases and n_cases >= a.max_cases:
        break
print(f"fuzz_layers: {n_cases} cases, {fails} failing, seed {a.seed}, {time.time() - t0:.0f} s")
sys.exit(1 if fails else 0)
