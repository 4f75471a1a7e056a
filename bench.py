#!/usr/bin/env python
"""bench.py -- queries/sec of exact cosine top-10 over a 10M x 768 bf16 corpus (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the C ABI)
    python bench.py --impl reference ...                     # the reference's CPU search, timed
                                                             # on the box's host cores

One "step" = one search of a query batch (default 4096 queries) against the whole corpus.
N > 1 (torchrun, one rank per GPU, NCCL): corpus rows are sharded contiguously, every rank
searches its shard, ONE all-gather moves the per-shard (score, index) lists and every rank
merges them (strong scaling: the 10M-row corpus is fixed, per-GPU work shrinks with N).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (queries already in HBM),
`e2e` = the same through the public Python API with the query batch copied from pinned host
memory and the results copied back every step.  `roofline` describes the dominant kernel
(the tcgen05 search kernel) timed with CUDA events on its own stream via the C-ABI hook.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N rows, D, Q per batch, k)
    "10Mx768_q4096_top10": (10_000_000, 768, 4096, 10),      # BASELINE.json metric / north_star target
    "1Mx768_q1024_top10": (1_000_000, 768, 1024, 10),        # BASELINE.json configs[1]
}
DEFAULT_WORKLOAD = "10Mx768_q4096_top10"
METRIC = "queries/sec exact top-10 over 10Mx768 bf16 corpus"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # under load = samples in the upper half of what was seen
        load = [x for x in sm if x >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def make_shard(rows: int, D: int, seed: int, dev) -> torch.Tensor:
    """Synthetic unit-norm bf16 corpus rows, generated on the device in chunks."""
    g = torch.Generator(device=dev).manual_seed(seed)
    out = torch.empty(rows, D, dtype=torch.bfloat16, device=dev)
    step = 1 << 18
    for s in range(0, rows, step):
        n = min(step, rows - s)
        x = torch.randn(n, D, generator=g, device=dev, dtype=torch.float32)
        x = x / x.norm(dim=-1, keepdim=True)
        out[s:s + n] = x.to(torch.bfloat16)
    return out


def cpu_reference_qps(N_full: int, D: int, k: int, budget_s: float = 12.0):
    """The reference's CPU search (cos_sim + topk, metrics.py:99-101 + search_pipeline.py:78, as
    restated in oracle/oracle.py) on a bounded sample of the workload, all host threads, repeated
    for about `budget_s` seconds; scaled linearly in N to the full corpus (stated in `sample`)."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_s = min(N_full, 500_000)
    q_s = 128
    g = torch.Generator().manual_seed(1234)
    corpus = torch.randn(n_s, D, generator=g)
    queries = torch.randn(q_s, D, generator=g)
    O.search_cos_sim_literal(queries[:4], corpus[:1000], k)  # warm-up
    t0 = time.perf_counter()
    done = 0
    while True:
        O.search_cos_sim_literal(queries, corpus, k)
        done += 1
        el = time.perf_counter() - t0
        if el >= budget_s or done >= 200:
            break
    per_call = el / done
    qps_full = (q_s / per_call) * n_s / N_full
    sample = (f"{q_s} fp32 queries x {n_s} rows x {D} (cos_sim + torch.topk), {done} reps in {el:.1f} s "
              f"({per_call:.3f} s each) on {cores} threads; scaled linearly in N to {N_full} rows")
    return qps_full, cores, sample, el


def run_reference(args, out_fd):
    N, D, Q, k = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    qps, cores, sample, spent = cpu_reference_qps(N, D, k, budget_s=min(90.0, max(10.0, 3.0 * steps)))
    ms_per_step = Q / qps * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "corpus_rows": N, "dim": D, "queries_per_step": Q, "k": k,
                   "rows_per_gpu": N, "sharding": "none (host CPU)", "cache": "n/a (CPU)",
                   "scores": "fp32 cos_sim + torch.topk, as the reference computes them",
                   "note": "CPU arm: oracle port of the reference's cos_sim+topk search on host cores"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(out_fd, line)


def _claim_stdout() -> int:
    """Keep stdout for the ONE JSON line: libraries (NCCL prints its version banner to stdout on some
    boxes) get stderr instead.  Returns the fd to write the JSON line to."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def _emit(fd: int, line: dict) -> None:
    os.write(fd, (json.dumps(line) + "\n").encode())


def main():
    out_fd = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--queries", type=int, default=0, help="override queries per step")
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--small-q", default="1,32", help="extra HBM-regime batch sizes reported under 'regimes' (N=1)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, out_fd)

    from text_similarity_b200 import _lib, ops
    from text_similarity_b200.sharded import ShardedCorpus

    N, D, Q, k = WORKLOADS[args.workload]
    Q = args.queries or Q
    k = args.k or k
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    lib = _lib.load()
    peaks = load_peaks()
    steps, warmup = args.steps, max(args.warmup, 0)

    # ---- data: contiguous row shard of the synthetic corpus, query batches -------------------
    rows_per = (N + world - 1) // world
    r0 = rank * rows_per
    rows = max(0, min(N, r0 + rows_per) - r0)
    shard = make_shard(rows, D, seed=1234 + rank, dev=dev)
    corpus = ShardedCorpus(shard, idx_base=r0, group=dist.group.WORLD if dist else None)
    nbatch = 4
    gq = torch.Generator(device="cpu").manual_seed(4321)
    host_batches = []
    for b in range(nbatch):
        x = torch.randn(Q, D, generator=gq)
        x = (x / x.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
        host_batches.append(x.pin_memory())
    dev_batches = [x.to(dev) for x in host_batches]
    host_scores = torch.empty(Q, k, dtype=torch.float32).pin_memory()
    host_idx = torch.empty(Q, k, dtype=torch.int64).pin_memory()
    torch.cuda.synchronize()

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if not dist:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- small-batch (HBM-bound) regime, 1 GPU only: same kernels, small Q.  Measured BEFORE the heavy
    # tensor-bound loop: a small-batch search alone does not hit the 1 kW power cap, the loop below does.
    regimes = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def small_batch_regime(corp, qbatch, n_rows, dim, esize, label):
        """One HBM-regime entry: whole call via CUDA-graph replay, candidate-pass kernels via the timing hook."""
        sq = qbatch.shape[0]
        run, graphed = (lambda: corp.search_graphed(qbatch, k)), True
        try:
            run()
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] CUDA-graph capture failed ({exc}); eager launches", file=sys.stderr)
            run, graphed = (lambda: corp.search(qbatch, k)), False
        for _ in range(3):
            run()
        reps = 10
        a0 = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
        a1 = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
        for ev in a0 + a1:
            ev.record()
        torch.cuda.synchronize()
        e0.record()
        for i in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        # candidate-pass kernels alone (eager calls: the timing hook records events around them)
        for i in range(reps):
            lib.tsim_set_timing_events(a0[i].cuda_event, a1[i].cuda_event)
            corp.search(qbatch, k)
        lib.tsim_set_timing_events(None, None)
        torch.cuda.synchronize()
        km = statistics.mean(x.elapsed_time(y) for x, y in zip(a0, a1))
        b_alg = n_rows * dim * esize + n_rows * 4 + sq * dim * esize + sq * k * 12
        f_alg = 2.0 * sq * n_rows * dim
        return {"workload": label, "queries_per_step": sq, "queries_per_s": reps * sq / (e0.elapsed_time(e1) * 1e-3),
                "ms_per_step": e0.elapsed_time(e1) / reps, "kernel_ms": km, "cuda_graph": graphed,
                "roofline": {"bound": "hbm", "achieved": b_alg / (km * 1e-3) / 1e9,
                             "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": b_alg / (km * 1e-3) / 1e9 / peaks["hbm_gbs"],
                             "achieved_whole_call": b_alg / (e0.elapsed_time(e1) / reps * 1e-3) / 1e9,
                             "tflops": f_alg / (km * 1e-3) / 1e12}}

    if world == 1 and args.small_q:
        small = [int(x) for x in args.small_q.split(",") if x]
        for sq in small:
            regimes.append(small_batch_regime(corpus, dev_batches[0][:sq].contiguous(), rows, D, 2, args.workload))
        if args.workload == DEFAULT_WORKLOAD:
            # BASELINE config 4's per-GPU shape: one 12.5M x 384 e4m3 shard of the 100M-row corpus (4.8 GB)
            n8, d8 = 12_500_000, 384
            g8 = torch.Generator(device=dev).manual_seed(99)
            c8 = torch.empty(n8, d8, dtype=torch.float8_e4m3fn, device=dev)
            for s0 in range(0, n8, 1 << 20):
                n = min(1 << 20, n8 - s0)
                x = torch.randn(n, d8, generator=g8, device=dev)
                c8[s0:s0 + n] = (x / x.norm(dim=-1, keepdim=True) * 64).to(torch.float8_e4m3fn)
            corpus8 = ShardedCorpus(c8)
            for sq in small:
                x = torch.randn(sq, d8, generator=g8, device=dev)
                q8 = (x / x.norm(dim=-1, keepdim=True) * 64).to(torch.float8_e4m3fn)
                regimes.append(small_batch_regime(corpus8, q8, n8, d8, 1, "12.5Mx384_e4m3_shard_top10"))
            del corpus8, c8
            torch.cuda.empty_cache()
        torch.cuda.synchronize()
        time.sleep(0.5)

    # ---- device-resident timed region ---------------------------------------------------------
    ev_k0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ev_k1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    for ev in ev_k0 + ev_k1:
        ev.record()  # torch creates the CUDA event lazily: materialise the handles up front
    for i in range(warmup):
        corpus.search(dev_batches[i % nbatch], k)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.tsim_launch_count()
    e0.record()
    for i in range(steps):
        lib.tsim_set_timing_events(ev_k0[i].cuda_event, ev_k1[i].cuda_event)
        corpus.search(dev_batches[i % nbatch], k)
    lib.tsim_set_timing_events(None, None)
    e1.record()
    gpu_launches = int(lib.tsim_launch_count() - launches0)   # libtsim kernels launched in the timed region
    barrier()
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    kern_ms = statistics.mean(a.elapsed_time(b) for a, b in zip(ev_k0, ev_k1))
    kern_ms = max_over_ranks(kern_ms)
    value = steps * Q / (total_ms * 1e-3)

    # ---- end-to-end timed region: pinned host queries in, host results out, every step --------
    for i in range(min(warmup, 2)):
        corpus.search_host(host_batches[i % nbatch], k, host_scores, host_idx)
    barrier()
    e0.record()
    for i in range(steps):
        corpus.search_host(host_batches[i % nbatch], k, host_scores, host_idx)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e = steps * Q / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel (per launch, this rank's shard) ----------------------
    flops = 2.0 * Q * rows * D
    bytes_alg = rows * D * 2 + rows * 4 + Q * D * 2 + Q * k * 12
    tflops = flops / (kern_ms * 1e-3) / 1e12
    gbs = bytes_alg / (kern_ms * 1e-3) / 1e9
    ridge = peaks["bf16_tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    tensor_bound = (flops / bytes_alg) > ridge
    if tensor_bound:
        roofline = {"bound": "tensor", "achieved": tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": tflops / peaks["bf16_tflops"],
                    "frac_of_sustained": tflops / peaks["bf16_tflops_sustained"], "traffic": None}
    else:
        roofline = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": gbs / peaks["hbm_gbs"], "traffic": None}
    if world == 1 and args.workload == DEFAULT_WORKLOAD and Q == WORKLOADS[DEFAULT_WORKLOAD][2]:
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload, one ncu --set full
        # capture (profiles/r01i_search_tc_bench_q4096.ncu_raw.txt)
        roofline["traffic"] = 16.87e9 + 0.33e9   # main-pass launch (98.4 % of the rows; the sample passes are not in this figure)
        roofline["traffic_source"] = "profiles/r01i_search_tc_bench_q4096.ncu_raw.txt"
    roofline.update({"kernel": "search_tc_kernel", "kernel_ms": kern_ms, "peak_source": peaks["source"],
                     "algorithmic_flops": flops, "algorithmic_bytes": bytes_alg})

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload ----------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        qps, cores, sample, _ = cpu_reference_qps(N, D, k, budget_s=12.0)
        cpu_baseline = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample}

    # ---- stock PyTorch on the same GPU (rank 0, N = 1 only; SURVEY.md 8d's second baseline): cuBLAS
    # `Q @ chunk.T` -> torch.topk per chunk -> topk of the concatenated lists.  bf16 scores, so its indices are
    # not exact; reported for scale only, after both timed regions.
    torch_gpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        def stock(qb):
            vals, idxs = [], []
            for s0 in range(0, rows, 500_000):
                sc = qb @ shard[s0:s0 + 500_000].T
                v, ix = torch.topk(sc, min(k, sc.shape[1]), dim=1)
                vals.append(v)
                idxs.append(ix + s0)
            v, ix = torch.cat(vals, 1), torch.cat(idxs, 1)
            top, pos = torch.topk(v.float(), min(k, v.shape[1]), dim=1)
            return top, torch.gather(ix, 1, pos)
        try:
            stock(dev_batches[0])
            e0.record()
            for i in range(2):
                stock(dev_batches[i % nbatch])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 2
            torch_gpu = {"value": Q / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms,
                         "what": "torch bf16 matmul (cuBLAS) + torch.topk, 500k-row chunks, [Q, chunk] scores through HBM; "
                                 "bf16 scores, indices not exact"}
        except Exception as exc:  # noqa: BLE001  (informational leg: never fail the bench on it)
            torch_gpu = {"value": None, "error": str(exc)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": total_ms / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": args.workload, "corpus_rows": N, "dim": D, "queries_per_step": Q, "k": k,
                       "rows_per_gpu": rows_per, "sharding": f"contiguous rows x{world}, 1 all-gather + merge",
                       "cache": "inputs larger than L2 (corpus shard %.1f GB per step)" % (rows * D * 2 / 1e9),
                       "scores": "float64 re-scored, exact index match vs oracle"},
            "e2e": {"value": e2e, "unit": "queries/s", "h2d_bytes_per_step": Q * D * 2,
                    "d2h_bytes_per_step": Q * k * 12, "ms_per_step": e2e_ms / steps},
            "gpu_launches": gpu_launches,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "torch_gpu_baseline": torch_gpu, "clocks": clocks,
            "regimes": regimes,
        }
        _emit(out_fd, line)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
