import os, sys, torch
sys.path.insert(0, "/root/repo")
from text_similarity_b200 import _lib, build, ops
build.build(experiment=True); _lib.use_experiment_build()
dev = torch.device("cuda")
N, D = 25_000_000, 384
x = torch.empty(N, D, dtype=torch.float8_e4m3fn, device=dev)
x.view(torch.uint8).random_(0, 120)
for knob in ("0", "1", "0", "1"):
    os.environ["TSIM_NO_MMA_NORM"] = knob
    ops.row_inv_norm(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): r = ops.row_inv_norm(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"NO_MMA_NORM={knob}: {ms:.3f} ms  {(N * D + N * 4) / ms / 1e6:.0f} GB/s", flush=True)
os.environ["TSIM_NO_MMA_NORM"] = "0"; a = ops.row_inv_norm(x)
os.environ["TSIM_NO_MMA_NORM"] = "1"; b = ops.row_inv_norm(x)
print("max rel diff", ((a - b).abs() / b).max().item())
# ragged shapes
for (n, d) in [(1, 16), (17, 48), (1000, 400), (33, 1040)]:
    y = (torch.randn(n, d, device=dev) * 3).to(torch.float8_e4m3fn)
    os.environ["TSIM_NO_MMA_NORM"] = "0"; a = ops.row_inv_norm(y)
    ref = 1.0 / y.double().norm(dim=-1).clamp_min(1e-8)
    print(n, d, "max rel err vs f64", ((a.double() - ref).abs() / ref).max().item())
