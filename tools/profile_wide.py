"""One build + query batches of the GIST shape (d = 960, k = 100) at reduced size: the command profiled under ncu for
the wide bucket-major kernels (k_score_wide, k_threshold_wide).  Prints stage times so the plain run documents itself."""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from similaritysearchbyrdf_b200 import DPFIndex, synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=300_000)
ap.add_argument("--nq", type=int, default=4000)
ap.add_argument("--d", type=int, default=960)
ap.add_argument("--topk", type=int, default=100)
ap.add_argument("--steps", type=int, default=0)
ap.add_argument("--repeat", type=int, default=2)
a = ap.parse_args()
X, Q = synth.config3(a.n, a.nq, a.d)
A, chain = synth.angle_family(a.d, max(100, a.d), 10, 3, 32, 88387 + 3)
Ap = synth.partitioner_family(30, 3, 88387 + 4)
ix = DPFIndex(d=a.d, L=30, k=32, pb=3)
ix.set_family(A, chain); ix.set_partitioners(Ap); ix.set_profiling(True)
Xd, Qd = torch.from_numpy(X).cuda(), torch.from_numpy(Q).cuda()
ix.fit_dense_dev(Xd.data_ptr(), a.n)
ids = torch.empty((a.nq, a.topk), dtype=torch.int32, device="cuda")
sc = torch.empty((a.nq, a.topk), dtype=torch.float64, device="cuda")
for _ in range(a.repeat):
    ix.query_topk_dense_dev(Qd.data_ptr(), a.nq, 0, a.steps, a.topk, 0, ids.data_ptr(), sc.data_ptr())
    torch.cuda.synchronize()
    st = ix.stats()
    print("query stage ms:", {k: round(v, 3) for k, v in ix.stage_times_ms().items() if v}, "dups/q", st["last_cand_with_dups"] / a.nq,
          "pairs", st["bm_pairs"], "units", st["bm_runs"], "rows staged", st["bm_rows_staged"], "survivors/q", st["bm_survivors"] / a.nq,
          "direct", st["bm_direct"])
