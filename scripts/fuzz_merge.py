"""Time-boxed randomized sweep of K3's merge (tsim_merge_topk / tsim_merge_topk_strided: per-shard or per-chunk
result lists -> top-k by (score desc, row asc), the merge the reference lacks, search_pipeline.py:83,88) against a
stable torch sort: best-first lists as the search emits them, unordered lists, heavy score ties, short lists padded
with row -1, k_out above and below what is available, and the rank-major all-gather layout read in place.

    python scripts/fuzz_merge.py --seconds 60 --seed 1

Exit status 1 on any mismatch."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_similarity_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=60.0)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--max-cases", type=int, default=0, help="stop after this many cases (0: run out the clock)")
a = ap.parse_args()
dev = torch.device("cuda", 0)


def expect(s, ix, k):
    s = s.clone()
    s[ix < 0] = -float("inf")
    by_idx = torch.sort(ix, dim=1, stable=True)[1]
    s, ix = torch.gather(s, 1, by_idx), torch.gather(ix, 1, by_idx)
    order = torch.sort(s, dim=1, descending=True, stable=True)[1][:, :k]
    es, ei = torch.gather(s, 1, order), torch.gather(ix, 1, order)
    if es.shape[1] < k:
        pad = k - es.shape[1]
        es = torch.cat([es, torch.full((es.shape[0], pad), -float("inf"), dtype=es.dtype, device=dev)], 1)
        ei = torch.cat([ei, torch.full((ei.shape[0], pad), -1, dtype=ei.dtype, device=dev)], 1)
    ei = torch.where(torch.isinf(es) & (es < 0), torch.full_like(ei, -1), ei)
    return es, ei


def run(case: int):
    rng = np.random.default_rng([a.seed, case])
    Q = int(rng.choice([1, 2, 31, 100, 1024, 4096]))
    n_lists = int(rng.choice([1, 2, 3, 8, 16, 37, 148, 300]))
    k_in = int(rng.choice([1, 5, 10, 16, 100, 128]))
    n_lists = max(1, min(n_lists, 4096 // k_in))         # the C ABI's bound: n_lists * k_in <= 4096
    while Q * n_lists * k_in > 2e7 and Q > 1:
        Q //= 2
    k_out = int(rng.choice([1, 5, 10, k_in, min(2 * k_in, 1024), 100]))
    g = torch.Generator(device=dev).manual_seed(int(rng.integers(0, 2 ** 31)))
    total = n_lists * k_in
    ties = int(rng.choice([0, 3, 50]))
    if ties:
        s = torch.randint(0, ties, (Q, total), generator=g, device=dev).double() / ties
    else:
        s = torch.rand(Q, total, generator=g, device=dev, dtype=torch.float64) * 2 - 1
    ix = torch.argsort(torch.rand(Q, total, generator=g, device=dev), dim=1) + int(rng.integers(0, 1 << 34))   # distinct rows per query
    # short lists: the tail of each list is padding
    fill = torch.randint(0, k_in + 1, (Q, n_lists), generator=g, device=dev)
    if rng.integers(0, 2):
        fill[:] = k_in
    pad = torch.arange(k_in, device=dev)[None, None, :] >= fill[..., None]
    ordered = bool(rng.integers(0, 2))
    s3, i3 = s.view(Q, n_lists, k_in).clone(), ix.view(Q, n_lists, k_in).clone()
    s3[pad] = -float("inf")
    i3[pad] = -1
    if ordered:
        # best first within every list, ties by lower row, padding last -- what tsim_search_topk writes
        by_idx = torch.sort(torch.where(i3 < 0, torch.full_like(i3, 1 << 62), i3), dim=2, stable=True)[1]
        s3, i3 = torch.gather(s3, 2, by_idx), torch.gather(i3, 2, by_idx)
        o = torch.sort(s3, dim=2, descending=True, stable=True)[1]
        s3, i3 = torch.gather(s3, 2, o), torch.gather(i3, 2, o)
    s2, i2 = s3.reshape(Q, total).contiguous(), i3.reshape(Q, total).contiguous()
    es, ei = expect(s2, i2, k_out)
    gs, gs64, gi = ops.merge_topk(s2, i2, k_out, n_lists)
    torch.cuda.synchronize()
    bad = int((gi != ei).sum()) + int((gs64 != es).sum()) + int((gs != es.float()).sum())
    desc = dict(case=case, Q=Q, n_lists=n_lists, k_in=k_in, k_out=k_out, ties=ties, ordered=ordered)
    if ordered and k_out == k_in:
        # the all-gather layout: per rank [Q, k] score bits, then [Q, k] rows
        recv = torch.stack([torch.stack([s3[:, r].contiguous().view(torch.int64), i3[:, r].contiguous()]) for r in range(n_lists)])
        rs, rs64, ri = ops.merge_gathered(recv.reshape(n_lists * 2, Q, k_in).contiguous(), Q, k_in, n_lists)
        torch.cuda.synchronize()
        bad += int((ri != ei).sum()) + int((rs64 != es).sum())
        desc["gathered"] = True
    return bad, desc


t0 = time.time()
n_cases = fails = 0
case = 0
while time.time() - t0 < a.seconds:
    try:
        bad, desc = run(case)
    except Exception as e:  # noqa: BLE001
        print(f"ERROR case {case}: {type(e).__name__}: {e}", flush=True)
        bad, desc = -1, {"case": case}
        if "CUDA" in str(e) or "cuda" in str(e):
            fails += 1
            break
    n_cases += 1
    if bad:
        fails += 1
        print(f"MISMATCH bad={bad} {desc}", flush=True)
    elif n_cases % 50 == 0:
        print(f"ok {desc}", flush=True)
    case += 1
    if a.max_cases and n_cases >= a.max_cases:
        break
print(f"fuzz_merge: {n_cases} cases, {fails} failing, seed {a.seed}, {time.time() - t0:.0f} s")
sys.exit(1 if fails else 0)
