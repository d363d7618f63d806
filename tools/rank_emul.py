"""Per-rank cost of a multi-GPU partition, emulated on ONE GPU: every rank of `--world` builds its shard of configs[1]
and answers the same 10k-query batch; prints each rank's stage times and counters (where the imbalance comes from)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from similaritysearchbyrdf_b200 import DPFIndex, synth

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--nq", type=int, default=10000)
ap.add_argument("--ranks", type=str, default="")
ap.add_argument("--fork", type=int, default=0, help="DPF_DBG_TAU_FORK: 2 = fork the threshold stream after the pair fill")
a = ap.parse_args()
X, Q = synth.config2(1_000_000, a.nq, 128)
A, chain = synth.angle_family(128, 128, 10, 3, 32, 88387 + 2)
Ap = synth.partitioner_family(30, 3, 88387 + 3)
Xd = torch.from_numpy(X).cuda()
Qd = torch.from_numpy(Q).cuda()
ids = torch.empty((a.nq, 10), dtype=torch.int32, device="cuda")
sc = torch.empty((a.nq, 10), dtype=torch.float64, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ranks = [int(r) for r in a.ranks.split(",")] if a.ranks else range(a.world)
for r in ranks:
    ix = DPFIndex(d=128, L=30, k=32, pb=3, rank=r, world=a.world)
    ix.set_balanced_partition(True)
    ix.set_family(A, chain); ix.set_partitioners(Ap)
    if a.fork:
        ix.set_debug_option(13, a.fork)
    stm = torch.cuda.Stream()
    ix.set_stream(stm.cuda_stream)
    torch.cuda.synchronize()
    ix.fit_dense_dev(Xd.data_ptr(), 1_000_000)
    for _ in range(3):
        ix.query_topk_dense_dev(Qd.data_ptr(), a.nq, 0, 0, 10, 0, ids.data_ptr(), sc.data_ptr())
    torch.cuda.synchronize()
    e0.record(stm)
    for _ in range(5):
        ix.query_topk_dense_dev(Qd.data_ptr(), a.nq, 0, 0, 10, 0, ids.data_ptr(), sc.data_ptr())
    e1.record(stm)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ix.set_profiling(True)
    ix.query_topk_dense_dev(Qd.data_ptr(), a.nq, 0, 0, 10, 0, ids.data_ptr(), sc.data_ptr())
    st = ix.stage_times_ms()
    s = ix.stats()
    print(f"rank {r}: {ms:.3f} ms/step", {k: round(v, 3) for k, v in st.items() if v}, "owned cells", int(ix.owned_subindexes().sum()),
          "pairs", s["bm_pairs"], "rows", s["bm_rows_staged"], "surv/q", round(s["bm_survivors"] / a.nq, 1), "direct", s["bm_direct"],
          "entries/q", round(s["last_cand_with_dups"] / a.nq), flush=True)
    ix.close()
