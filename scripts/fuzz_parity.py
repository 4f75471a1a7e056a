"""Time-boxed randomized parity sweep on one B200: whichever kernels the planner picks for a random (shape, dtype,
k, data distribution, view, shard offset, self exclusion) must give exactly what the float64 exact scan of the same
library gives -- same indices, same float64 score bits (tests/test_gpu_parity.py pins that scan to the CPU oracle).

    python scripts/fuzz_parity.py --seconds 240 --seed 1 [--max-rows 2500000] [--big]

Prints one line per failing case (with everything needed to replay it: --only CASE) and a summary; exit status 1 on
any mismatch.  Data distributions: Gaussian, unit rows, tight clusters (near-duplicates closer than bf16 resolution),
exact duplicates, a handful of distinct rows (massive ties), zero rows, wide dynamic range of row norms, one dense
cluster around a query."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from text_similarity_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=240.0)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--max-cases", type=int, default=0, help="stop after this many cases (0: run out the clock)")
ap.add_argument("--max-rows", type=int, default=2_500_000)
ap.add_argument("--big", action="store_true", help="shards of >= 290K rows and batches up to 4096 queries only "
                "(round-robin plans, sample passes, append lists: the headline's code path)")
ap.add_argument("--only", type=int, default=-1, help="replay one case number of this seed")
ap.add_argument("--explain", action="store_true", help="on a mismatch print, for a few queries, the rows either answer "
                "holds with their float64 torch scores (exact for small-integer rows)")
ap.add_argument("--cases", default="", help="replay a comma list of case numbers of this seed")
a = ap.parse_args()
dev = torch.device("cuda", 0)

DIST = ["gauss", "unit", "clusters", "dups", "few_distinct", "zeros", "range", "dense_cluster"]
DIST2 = ["scale", "ints", "onehot"]      # added later, drawn from a second generator so that old case numbers replay


def make(case: int):
    rng = np.random.default_rng([a.seed, case])
    dtype = [torch.bfloat16, torch.float8_e4m3fn, torch.float32, torch.float16][int(rng.integers(0, 4))]
    step = 16 if dtype == torch.float8_e4m3fn else 8
    D = int(rng.choice([1, 2, 3, 5, 8, 16, 17, 24, 32, 33, 48, 64, 96, 97, 128])) * step
    size_class = int(rng.integers(2, 4)) if a.big else int(rng.integers(0, 4))
    rng3 = np.random.default_rng([a.seed, case, 11])       # (a later addition, own generator: old case numbers replay)
    if rng3.integers(0, 16) == 0:
        D = int(rng3.choice([200, 256, 520, 1030])) * step   # wide rows: approx_eps grows with D beyond 1664 (bf16)
        size_class = min(size_class, 2)
    hi = [2_000, 60_000, 420_000, a.max_rows][size_class]
    lo = [1, 2_000, 290_000, 420_000][size_class]
    # keep the corpus under ~2 GB of elements
    hi = max(lo + 1, min(hi, int(2.0e9 / (D * 4))))
    N = int(rng.integers(lo, hi))
    Q = int(rng.choice([1, 2, 7, 8, 9, 31, 32, 33, 64, 100, 127, 128, 129, 200, 240, 256, 257, 511, 600, 1024, 1500]))
    if a.big:
        Q = int(rng.choice([129, 600, 1200, 1700, 2500, 4096, 5000]))
    if N * Q * D > 6e12:
        Q = max(1, int(6e12 / (N * D)))
    k = int(rng.choice([1, 2, 5, 10, 16, 24, 25, 40, 64, 100, 101, 128]))
    dist = DIST[int(rng.integers(0, len(DIST)))]
    rng2 = np.random.default_rng([a.seed, case, 7])
    if rng2.integers(0, 11) < len(DIST2):
        dist = DIST2[int(rng2.integers(0, len(DIST2)))]
    g = torch.Generator(device=dev).manual_seed(int(rng.integers(0, 2 ** 31)))
    pitch = D + (step * int(rng.integers(0, 3)) if rng.integers(0, 3) == 0 else 0)
    base = torch.randn(N, pitch, generator=g, device=dev)
    c = base[:, :D]
    q = torch.randn(Q, D, generator=g, device=dev)
    if dist == "unit":
        c /= c.norm(dim=-1, keepdim=True).clamp_min(1e-20)
        q /= q.norm(dim=-1, keepdim=True).clamp_min(1e-20)
    elif dist == "clusters" and N > 64:
        cent = torch.randn(16, D, generator=g, device=dev)
        owner = torch.randint(0, 16, (N,), generator=g, device=dev)
        c.copy_(cent[owner] + c * float(rng.choice([1e-1, 1e-2, 1e-3])))
        q.copy_(cent[torch.randint(0, 16, (Q,), generator=g, device=dev)] + q * 0.05)
    elif dist == "dups" and N > 4:
        n_d = max(1, N // int(rng.choice([2, 10, 1000])))
        src = torch.randint(0, N, (n_d,), generator=g, device=dev)
        dst = torch.randint(0, N, (n_d,), generator=g, device=dev)
        c[dst] = c[src]
    elif dist == "few_distinct":
        m = int(rng.choice([1, 3, 50]))
        proto = torch.randn(m, D, generator=g, device=dev)
        c.copy_(proto[torch.randint(0, m, (N,), generator=g, device=dev)])
    elif dist == "zeros":
        c[torch.rand(N, generator=g, device=dev) < 0.3] = 0
        q[torch.rand(Q, generator=g, device=dev) < 0.3] = 0
    elif dist == "range":
        c *= torch.exp(torch.randn(N, 1, generator=g, device=dev) * 2.0)
        q *= torch.exp(torch.randn(Q, 1, generator=g, device=dev) * 2.0)
    elif dist == "dense_cluster" and N > 10_000:
        n_c = int(min(N // 2, rng.choice([300, 6000, 50_000])))
        rows = torch.randperm(N, generator=g, device=dev)[:n_c]
        c[rows] = q[0] + torch.randn(n_c, D, generator=g, device=dev) * 0.02
    elif dist == "scale" and dtype in (torch.bfloat16, torch.float32):
        # rows 24 orders of magnitude apart (cosine is scale free; nothing under- or overflows in fp32)
        c *= 10.0 ** (torch.rand(N, 1, generator=g, device=dev) * 24 - 12)
        q *= 10.0 ** (torch.rand(Q, 1, generator=g, device=dev) * 24 - 12)
    elif dist == "ints":
        # small integers: distinct rows with exactly equal cosines (the tie rule decides), exact zeros
        c.copy_(torch.randint(-2, 3, (N, D), generator=g, device=dev).float())
        q.copy_(torch.randint(-2, 3, (Q, D), generator=g, device=dev).float())
    elif dist == "onehot":
        c.zero_()
        c[torch.arange(N, device=dev), torch.randint(0, D, (N,), generator=g, device=dev)] = 1.0
        if D > 1:
            c[torch.arange(N, device=dev), torch.randint(0, D, (N,), generator=g, device=dev)] += 0.5
    if dtype == torch.float8_e4m3fn:
        c.clamp_(-30, 30); q.clamp_(-30, 30)
        c *= 8; q *= 8
    # the cast keeps the (possibly padded) row pitch: search through a strided view
    cb = torch.empty(N, pitch, dtype=dtype, device=dev)
    cb[:, :D] = c.to(dtype)
    c = cb[:, :D]
    q = q.to(dtype)
    kw = {}
    if rng.integers(0, 3) == 0:
        kw["idx_base"] = int(rng.integers(0, 1 << 33))
    if rng.integers(0, 5) == 0 and N >= Q:
        off = int(rng.integers(0, N - Q + 1))
        q = c[off:off + Q].contiguous()
        kw["exclude_self_base"] = kw.get("idx_base", 0) + off
    use_inv = bool(rng.integers(0, 2))
    shadow = None
    if dtype in (torch.float32, torch.float16) and D % 8 == 0 and N > 0:
        shadow = "split" if (k > 24 or rng.integers(0, 3) == 0) else "rounded"
    desc = dict(case=case, dtype=str(dtype).replace("torch.", ""), N=N, Q=Q, D=D, pitch=pitch, k=k, dist=dist,
                inv=use_inv, shadow=shadow, **kw)
    return q, c, k, kw, use_inv, shadow, desc


def run(case: int):
    q, c, k, kw, use_inv, shadow, desc = make(case)
    extra = {}
    if shadow:
        sh, sinv = ops.make_shadow(c, split=(shadow == "split"))
        extra = dict(corpus_shadow=sh, shadow_inv_norm=sinv)
        mode = "auto"
    else:
        mode = "auto"
        if use_inv and c.dtype in (torch.bfloat16, torch.float8_e4m3fn):
            extra["corpus_inv_norm"] = ops.row_inv_norm(c)
    got = ops.search_topk(q, c, k, mode=mode, return_score64=True, return_flags=True, **kw, **extra)
    ref = ops.search_topk(q, c, k, mode="exact", return_score64=True, **kw)
    torch.cuda.synchronize()
    bad_i = int((got[1] != ref[1]).sum())
    if a.explain and bad_i:
        base = kw.get("idx_base", 0)
        c64 = c.double()
        cn = c64.pow(2).sum(-1).sqrt().clamp_min(1e-8)
        for qi in (got[1] != ref[1]).any(1).nonzero().flatten()[:3].tolist():
            q64 = q[qi].double()
            sc = (c64 @ q64) / (q64.pow(2).sum().sqrt().clamp_min(1e-8) * cn)
            if "exclude_self_base" in kw:
                sc[kw["exclude_self_base"] - base + qi] = -float("inf")
            order = torch.sort(sc, descending=True, stable=True)[1][:k]
            g, r = (got[1][qi] - base).tolist(), (ref[1][qi] - base).tolist()
            t = order.tolist()
            print(f"  query {qi} flag={int(got[3][qi])}: tensor-path == torch {g == t}, scan == torch {r == t}")
            for pos in range(k):
                if g[pos] != r[pos] or g[pos] != t[pos]:
                    print(f"    pos {pos}: tensor {g[pos]} ({float(sc[g[pos]]).hex()}, returned {float(got[2][qi][pos]).hex()})  "
                          f"scan {r[pos]} ({float(sc[r[pos]]).hex()}, returned {float(ref[2][qi][pos]).hex()})  "
                          f"torch {t[pos]} ({float(sc[t[pos]]).hex()})")
            kth = float(sc[t[-1]])
            tied = ((sc - kth).abs() <= 1e-12).nonzero().flatten().tolist()
            print(f"    rows within 1e-12 of the k-th score: {len(tied)}: {tied[:12]}  "
                  f"distinct bit patterns {len(set(float(sc[x]).hex() for x in tied))}")
    # (-inf == -inf; NaN never appears in the output)
    bad_s = int((got[2] != ref[2]).sum())
    bad_f = int((got[0] != ref[0]).sum())
    desc["flag1"] = int((got[3] == 1).sum())
    desc["flag2"] = int((got[3] == 2).sum())
    return bad_i, bad_s, bad_f, desc


t0 = time.time()
n_cases = fails = 0
replay = [int(x) for x in a.cases.split(",") if x] or ([a.only] if a.only >= 0 else [])
case = replay[0] if replay else 0
while True:
    try:
        bad_i, bad_s, bad_f, desc = run(case)
    except Exception as e:  # noqa: BLE001 -- a planner / argument error on a legal input is a finding too
        print(f"ERROR case {case}: {type(e).__name__}: {e}", flush=True)
        bad_i = bad_s = bad_f = -1
        desc = {"case": case}
        if "CUDA" in str(e) or "cuda" in str(e):
            fails += 1
            break
    n_cases += 1
    if bad_i or bad_s or bad_f:
        fails += 1
        print(f"MISMATCH idx={bad_i} s64={bad_s} s32={bad_f} {desc}", flush=True)
    elif replay or n_cases % 20 == 0:
        print(f"ok {desc}", flush=True)
    if replay:
        if n_cases == len(replay):
            break
        case = replay[n_cases]
        continue
    case += 1
    if time.time() - t0 > a.seconds or (a.max_cases and n_cases >= a.max_cases):
        break
print(f"fuzz: {n_cases} cases, {fails} failing, seed {a.seed}, {time.time() - t0:.0f} s")
sys.exit(1 if fails else 0)
