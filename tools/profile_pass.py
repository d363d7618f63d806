"""One build + one query batch of BASELINE configs[1] at reduced query count: the command profiled under ncu
(launch list / --set full).  Prints stage times so the plain run documents itself."""
import argparse
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from similaritysearchbyrdf_b200 import DPFIndex, synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--nq", type=int, default=1000)
ap.add_argument("--d", type=int, default=128)
ap.add_argument("--topk", type=int, default=10)
ap.add_argument("--repeat", type=int, default=1)
ap.add_argument("--world", type=int, default=1, help="emulate one rank of a multi-GPU partition on this GPU")
ap.add_argument("--rank", type=int, default=0)
ap.add_argument("--tc", type=int, default=0, help="1: the tcgen05 scoring kernel")
a = ap.parse_args()
X, Q = synth.config2(a.n, a.nq, a.d)
A, chain = synth.angle_family(a.d, max(100, a.d), 10, 3, 32, 88387 + 2)
Ap = synth.partitioner_family(30, 3, 88387 + 3)
ix = DPFIndex(d=a.d, L=30, k=32, pb=3)
ix.set_family(A, chain)
ix.set_partitioners(Ap)
ix.set_profiling(True)
Xd = torch.from_numpy(X).cuda()
for _ in range(a.repeat):
    ix2 = DPFIndex(d=a.d, L=30, k=32, pb=3, rank=a.rank, world=a.world)
    ix2.set_balanced_partition(a.world > 1)
    ix2.set_family(A, chain); ix2.set_partitioners(Ap); ix2.set_profiling(True)
    if a.tc:
        ix2.set_debug_option(3, 3)
    ix2.fit_dense_dev(Xd.data_ptr(), a.n)
    print("build stage ms:", {k: round(v, 3) for k, v in ix2.stage_times_ms().items() if v}, ix2.stats())
    ix = ix2
for _ in range(a.repeat + 1):
    ids, sc = ix.query_topk_dense(Q, None, 0, a.topk, 0)
    st = ix.stats()
    print("query stage ms:", {k: round(v, 3) for k, v in ix.stage_times_ms().items() if v},
          "cand/q", st["last_candidates"] / a.nq, "dups/q", st["last_cand_with_dups"] / a.nq,
          "bm pairs", st["bm_pairs"], "runs", st["bm_runs"], "rows staged", st["bm_rows_staged"], "survivors/q", st["bm_survivors"] / a.nq,
          "direct", st["bm_direct"])

d = ix.tc_diag()
print("tc watchdog", d[:8].tolist())
print("tc wait cycles by barrier tag (1 a_full@mma 2 acc_empty@mma 3 acc_full@epi 4 a_empty@prod 6 rec_full 7 rec_empty@loader), kernel cycles x CTAs:",
      d[8:].tolist())
# how many bucket rows a batch stages as a function of the unit width (queries scored per pass over a bucket)
off, ln = ix.leaf_pairs()
cnt = np.diff(off.astype(np.int64))
print("leaves", len(ln), "probed", int((cnt > 0).sum()), "pairs", int(cnt.sum()), "pairs/probed leaf mean", cnt[cnt > 0].mean(),
      "p50/p90/p99/max", np.percentile(cnt[cnt > 0], [50, 90, 99]).tolist(), int(cnt.max()))
for uq in (8, 16, 32, 64, 128, 256):
    units = -(-cnt // uq)
    rows = int((units * ln).sum())
    rows128 = int((units * (-(-ln // 128)) * 128).sum())
    print(f"unit width {uq:4d}: units {int(units.sum()):9d} rows staged {rows:11d} (padded to 128-row tiles {rows128:11d})")
