"""The two extra baselines SURVEY.md 8(d) asks for, next to bench.py's own arms.

  (1) stock PyTorch on the same B200: rows already unit-norm bf16, `scores = Q @ chunk.T` (cuBLAS) ->
      `torch.topk` per chunk -> concatenate the per-chunk lists -> `torch.topk` again.  This is what a user of
      the reference gets by moving its tensors to the GPU and batching the query loop; the [Q, chunk] score
      block goes through HBM (the thing the fused tcgen05 epilogue avoids).  bf16 scores: indices are NOT
      exact (ties / rounding at bf16 resolution) -- the recall against the exact result is printed.
  (2) the reference's per-query CPU loop as written (search_pipeline.py:73-79, repaired per Appendix A):
      `F.cosine_similarity(q.expand_as(C), C, -1)` + `torch.topk` for every query, fp32, all host threads,
      on a bounded sample, scaled linearly in N and Q.

    python scripts/time_baselines.py [--rows 10000000] [--queries 4096] [--k 10] [--chunk 500000]
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_shard  # noqa: E402
from text_similarity_b200 import ops  # noqa: E402


def stock_torch_search(q, corpus, k, chunk):
    vals, idxs = [], []
    for s in range(0, corpus.shape[0], chunk):
        sc = q @ corpus[s:s + chunk].T                     # [Q, chunk] bf16, written to HBM
        v, i = torch.topk(sc, min(k, sc.shape[1]), dim=1)
        vals.append(v)
        idxs.append(i + s)
    v = torch.cat(vals, 1)
    i = torch.cat(idxs, 1)
    top, pos = torch.topk(v.float(), k, dim=1)
    return top, torch.gather(i, 1, pos)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--chunk", type=int, default=500_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    corpus = make_shard(a.rows, a.dim, 1234, dev)
    g = torch.Generator().manual_seed(4321)
    x = torch.randn(a.queries, a.dim, generator=g)
    q = (x / x.norm(dim=-1, keepdim=True)).to(torch.bfloat16).to(dev)

    def timed(fn, reps):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, out

    ms_stock, (_, idx_stock) = timed(lambda: stock_torch_search(q, corpus, a.k, a.chunk), a.reps)
    inv = ops.row_inv_norm(corpus)
    ms_ours, (_, idx_ours) = timed(lambda: ops.search_topk(q, corpus, a.k, corpus_inv_norm=inv), a.reps)
    same_rows = (idx_stock == idx_ours).all(dim=1).float().mean().item()
    recall = sum(len(set(r0) & set(r1)) for r0, r1 in zip(idx_stock.tolist(), idx_ours.tolist())) / idx_ours.numel()
    flops = 2.0 * a.queries * a.rows * a.dim
    out = {"workload": f"{a.rows}x{a.dim} bf16, Q={a.queries}, k={a.k}",
           "stock_torch_b200": {"ms_per_search": ms_stock, "queries_per_s": a.queries / ms_stock * 1e3,
                                "tflops": flops / ms_stock / 1e9, "chunk_rows": a.chunk,
                                "score_block_bytes_through_hbm": 2 * 2 * a.queries * a.rows,
                                "queries_whose_top_k_equals_exact": same_rows, "recall_vs_exact": recall},
           "this_build": {"ms_per_search": ms_ours, "queries_per_s": a.queries / ms_ours * 1e3,
                          "tflops": flops / ms_ours / 1e9},
           "speedup_vs_stock_torch": ms_stock / ms_ours}

    # (2) reference per-query loop on the host, bounded sample
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_s, q_s = min(a.rows, 200_000), 16
    gc = torch.Generator().manual_seed(1)
    C = torch.randn(n_s, a.dim, generator=gc)
    Qc = torch.randn(q_s, a.dim, generator=gc)
    t0, done = time.perf_counter(), 0
    while time.perf_counter() - t0 < a.cpu_seconds:
        for qi in range(q_s):
            s = F.cosine_similarity(Qc[qi].expand_as(C), C, -1)
            torch.topk(s, a.k, largest=True, sorted=False)
        done += 1
    el = time.perf_counter() - t0
    qps_sample = done * q_s / el
    out["reference_cpu_per_query_loop"] = {
        "queries_per_s_scaled": qps_sample * n_s / a.rows, "cores": cores,
        "sample": f"{q_s} queries x {n_s} rows x {a.dim} fp32, {done} passes in {el:.1f} s; scaled linearly in N to {a.rows} rows"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
