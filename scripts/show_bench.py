"""Pretty-print a bench.py JSON line (last line of the given file): headline, verification, regimes."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ("n_gpus", "value", "ms_per_step", "gpu_launches")}, "e2e", round(d["e2e"]["value"], 1))
v = d["verified"]
print("verified:", {k: v[k] for k in ("queries", "mismatches", "rescored", "max_rescore_err", "e2e_host_copy_matches")})
print("roofline:", {k: (round(x, 3) if isinstance(x, float) else x) for k, x in d["roofline"].items() if k not in ("regimes", "verified")})
for r in d["regimes"]:
    rf = r.get("roofline", {})
    print(f"{r['name']:44s} n={r.get('n_gpus')} {r['value']:12.1f} {r['unit']:11s} ms {r.get('ms_per_step', 0):9.4f} eager {r.get('ms_per_step_eager') or 0:8.4f} "
          f"{rf.get('bound', '-'):6s} {rf.get('achieved', 0):8.1f} frac {rf.get('frac', 0):.3f} whole {rf.get('achieved_whole_call') or 0:7.1f} "
          f"mism {(r.get('verified') or {}).get('mismatches')}"
          + (f"  sustained {r['sustained']['ms_per_step']:.4f} ms = {r['sustained']['whole_call_gbs']:.0f} GB/s whole call" if r.get("sustained") else ""))
print("cpu:", d.get("cpu_baseline"))
print("clocks:", d.get("clocks"))
