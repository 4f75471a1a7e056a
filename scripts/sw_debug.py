"""Debug driver for the small-batch kernel: one search, then the thresholds / append counts / sample lists it left in
the workspace (layout of make_search_plan for a swapped plan)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_shard  # noqa: E402
from text_similarity_b200 import ops  # noqa: E402

rows, D, Q = int(os.environ.get("ROWS", "2000000")), int(os.environ.get("D", "384")), int(os.environ.get("Q", "8"))
dtype = torch.float8_e4m3fn if os.environ.get("DT", "fp8") == "fp8" else torch.bfloat16
dev = torch.device("cuda")
corpus = make_shard(rows, D, 1, dev, dtype)
inv = ops.row_inv_norm(corpus)
q = make_shard(Q, D, 2, dev, dtype)
s, i, s64, fl = ops.search_topk(q, corpus, 10, corpus_inv_norm=inv, return_score64=True, return_flags=True)
torch.cuda.synchronize()
ws = [v for (d, st, tag), v in ops._workspaces.items()][0]
al = lambda x: (x + 255) // 256 * 256
NC, KP, cap = 192, 16, 4096
off_app = al(Q * NC * KP * 8)
off_thr = al(off_app + Q * cap * 8)
off_flag = off_thr + al(Q * 4)
off_appcnt = off_flag + 256
cand = ws[:Q * NC * KP * 8].view(torch.int64).view(Q, NC * KP)
thr = ws[off_thr:off_thr + Q * 4].view(torch.int32)
cnt = ws[off_appcnt:off_appcnt + Q * 4].view(torch.int32)
flagcnt = ws[off_flag:off_flag + 4].view(torch.int32)
off_rthr = off_appcnt + al(Q * 4)
off_rflag = off_rthr + al(128 * 4)
print("flag_cnt", flagcnt.tolist(), "r_flag_cnt", ws[off_rflag:off_rflag + 4].view(torch.int32).tolist())
print("flags", fl.tolist())
print("thr(ord)", [hex(x & 0xffffffff) for x in thr.tolist()])
print("app_cnt", cnt.tolist())
print("live sample keys per query", (cand != 0).sum(1).tolist())
top = (cand.view(torch.int64) >> 32) & 0xffffffff
print("best sample ord", [hex(int(x)) for x in top.max(1).values.tolist()])
es, ei = ops.search_topk(q, corpus, 10, corpus_inv_norm=inv, mode="exact")
print("match exact:", torch.equal(ei, i), "top scores", s[:, 0].tolist()[:4], s[:, 9].tolist()[:4])
