"""Time-boxed randomized sweep of K1 (tsim_pool_norm: masked mean-pool -> optional L2 normalise -> cast, reference
src/modules/modules.py:158-171) on one B200 against the same formula evaluated in float64 with plain torch ops:
random shapes (odd widths included), token / mask dtypes, real-valued weights, interior holes, all-masked sentences,
NaN garbage under masked tokens, batch- and token-strided views, scattered output rows, every output dtype.

    python scripts/fuzz_pool.py --seconds 120 --seed 1 [--cases 3,17]

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
ap.add_argument("--seconds", type=float, default=120.0)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--max-cases", type=int, default=0, help="stop after this many cases (0: run out the clock)")
ap.add_argument("--cases", default="")
a = ap.parse_args()
dev = torch.device("cuda", 0)


def run(case: int):
    rng = np.random.default_rng([a.seed, case])
    in_dt = [torch.float32, torch.float16, torch.bfloat16][int(rng.integers(0, 3))]
    out_dt = [torch.float32, torch.bfloat16, torch.float8_e4m3fn][int(rng.integers(0, 3))]
    D = int(rng.choice([1, 3, 8, 30, 64, 100, 128, 384, 385, 512, 768, 1000, 1024, 2048, 4096]))
    L = int(rng.choice([1, 2, 7, 16, 31, 64, 100, 256, 300]))
    B = int(rng.choice([1, 2, 15, 16, 33, 128, 500, 2000]))
    while B * L * D > 4e7 and B > 1:
        B = max(1, B // 2)
    normalize = bool(rng.integers(0, 2)) or out_dt == torch.float8_e4m3fn
    mask_dt = [torch.bool, torch.uint8, torch.int32, torch.int64, torch.float32][int(rng.integers(0, 5))]
    view = ["plain", "batch_stride", "token_stride", "both"][int(rng.integers(0, 4))]
    g = torch.Generator(device=dev).manual_seed(int(rng.integers(0, 2 ** 31)))
    Lp = L + (int(rng.integers(1, 5)) if view in ("batch_stride", "both") else 0)
    Dp = D + (8 * int(rng.integers(1, 4)) if view in ("token_stride", "both") else 0)
    big = (torch.randn(B, Lp, Dp, generator=g, device=dev) * float(rng.choice([1e-3, 1.0, 30.0]))).to(in_dt)
    tok = big[:, :L, :D]
    lens = torch.randint(0, L + 1, (B,), generator=g, device=dev)
    if B > 1:
        lens[0] = L
    m = (torch.arange(L, device=dev)[None] < lens[:, None])
    if rng.integers(0, 2):
        m = m & (torch.rand(B, L, generator=g, device=dev) > 0.2)      # interior holes
    w = m.to(torch.float32)
    if mask_dt == torch.float32 and rng.integers(0, 2):
        w = w * torch.rand(B, L, generator=g, device=dev)               # real-valued weights
    mask = w.to(mask_dt) if mask_dt == torch.float32 else m.to(mask_dt)
    w64 = mask.to(torch.float64)
    garbage = bool(rng.integers(0, 3) == 0)
    e64 = tok.to(torch.float64)
    if garbage:
        tok.masked_fill_((w64 == 0)[..., None], float("nan"))           # must never reach the sum
    exp = (torch.where(w64[..., None] != 0, e64 * w64[..., None], torch.zeros((), dtype=torch.float64, device=dev))).sum(1)
    exp = exp / w64.sum(1, keepdim=True).clamp_min(1e-9)
    pooled_norm = exp.norm(dim=-1, keepdim=True)
    if normalize:
        exp = exp / pooled_norm.clamp_min(1e-8)
    scatter = bool(rng.integers(0, 3) == 0)
    if scatter:
        pad = int(rng.integers(0, 3)) * 16
        store = torch.full((B + 7, D + pad), 7.0, device=dev).to(out_dt)
        rows = torch.randperm(B + 7, generator=g, device=dev)[:B]
        out, inv = ops.pool_norm(tok, mask, normalize=normalize, out=store[:, :D], out_rows=rows)
        sf = store.float()                                    # (float8 tensors do not index)
        got = sf[rows, :D].to(torch.float64)
        inv = inv[rows]
        untouched = torch.ones(B + 7, dtype=torch.bool, device=dev)
        untouched[rows] = False
        clean = bool((sf[untouched] == 7.0).all()) and bool((sf[:, D:] == 7.0).all())
    else:
        out, inv = ops.pool_norm(tok, mask, out_dtype=out_dt, normalize=normalize)
        got = out.float().to(torch.float64)
        clean = True
    torch.cuda.synchronize()
    desc = dict(case=case, B=B, L=L, D=D, in_dt=str(in_dt)[6:], out_dt=str(out_dt)[6:], mask=str(mask_dt)[6:], view=view,
                normalize=normalize, garbage=garbage, scatter=scatter)
    # error scale of a row: the mean absolute term of its sums (fp32 accumulation), amplified by the normalisation
    # when the pooled vector is short against its terms (cancellation)
    A = ((e64 * w64[..., None]).abs().sum(1) / w64.sum(1, keepdim=True).clamp_min(1e-9)).amax(-1, keepdim=True)
    if normalize:
        A = A / pooled_norm.clamp_min(1e-300)
    if out_dt == torch.float8_e4m3fn:
        # stored times a per-row power of two: compare directions of the non-zero rows
        nz = exp.norm(dim=-1) > 1e-30
        gn = got / got.norm(dim=-1, keepdim=True).clamp_min(1e-300)
        en = exp / exp.norm(dim=-1, keepdim=True).clamp_min(1e-300)
        cos = (gn * en).sum(-1)[nz]
        bad = int((~(cos > 0.995)).sum()) + int((got[~nz] != 0).sum())
        err = float((1 - cos).max()) if cos.numel() else 0.0
    else:
        d = (got - exp).abs()
        tol = 1e-5 * A + 1e-30 + (2 ** -8 * exp.abs() if out_dt == torch.bfloat16 else 0.0)
        bad = int((~(d <= tol)).sum())
        err = float((d / A.clamp_min(1e-30)).max())
    # inv_norm describes the STORED row
    n = got.norm(dim=-1)
    live = n > 1e-6
    inv_bad = int((~((inv.double() * n - 1.0).abs()[live] <= 1e-4)).sum())
    return bad, inv_bad, clean, err, desc


t0 = time.time()
n_cases = fails = 0
replay = [int(x) for x in a.cases.split(",") if x]
case = replay[0] if replay else 0
while True:
    try:
        bad, inv_bad, clean, err, desc = run(case)
    except Exception as e:  # noqa: BLE001
        print(f"ERROR case {case}: {type(e).__name__}: {e}", flush=True)
        bad, inv_bad, clean, err, desc = -1, 0, True, 0.0, {"case": case}
        if "CUDA" in str(e) or "cuda" in str(e):
            fails += 1
            break
    n_cases += 1
    if bad or inv_bad or not clean:
        fails += 1
        print(f"MISMATCH elems={bad} inv={inv_bad} clean={clean} maxerr={err:.3e} {desc}", flush=True)
    elif replay or n_cases % 50 == 0:
        print(f"ok maxerr={err:.3e} {desc}", flush=True)
    if replay:
        if n_cases == len(replay):
            break
        case = replay[n_cases]
        continue
    case += 1
    if time.time() - t0 > a.seconds or (a.max_cases and n_cases >= a.max_cases):
        break
print(f"fuzz_pool: {n_cases} cases, {fails} failing, seed {a.seed}, {time.time() - t0:.0f} s")
sys.exit(1 if fails else 0)
