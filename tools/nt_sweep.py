"""Step time of configs[1] against the number of threshold samples per query (DPF_DBG_TAU_TABLES); with --world / --rank
the shard one rank of a multi-GPU partition holds (its local step, without the final all-gather)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from similaritysearchbyrdf_b200 import DPFIndex, synth, _lib as B
ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=1)
ap.add_argument("--rank", type=int, default=0)
ap.add_argument("--nts", type=str, default="6,5,4,3,2,6")
a = ap.parse_args()
X, Q = synth.config2(1_000_000, 10_000, 128)
A, chain = synth.angle_family(128, 128, 10, 3, 32, 88387 + 2)
Ap = synth.partitioner_family(30, 3, 88387 + 3)
Xd, Qd = torch.from_numpy(X).cuda(), torch.from_numpy(Q).cuda()
ids = torch.empty((10000, 10), dtype=torch.int32, device="cuda"); sc = torch.empty((10000, 10), dtype=torch.float64, device="cuda")
ix = DPFIndex(d=128, L=30, k=32, pb=3, rank=a.rank, world=a.world)
if a.world > 1:
    ix.set_balanced_partition(True)
ix.set_family(A, chain); ix.set_partitioners(Ap)
st = torch.cuda.Stream()
ix.set_stream(st.cuda_stream)
with torch.cuda.stream(st):
    ix.fit_dense_dev(Xd.data_ptr(), 1_000_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for nt in [int(v) for v in a.nts.split(',')]:
        ix.set_debug_option(B.DBG_TAU_TABLES, nt)
        ix.set_profiling(False)
        for _ in range(20):
            ix.query_topk_dense_dev(Qd.data_ptr(), 10000, 0, 0, 10, 0, ids.data_ptr(), sc.data_ptr())
        torch.cuda.synchronize()
        e0.record(st)
        for _ in range(20):
            ix.query_topk_dense_dev(Qd.data_ptr(), 10000, 0, 0, 10, 0, ids.data_ptr(), sc.data_ptr())
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        ix.set_profiling(True)
        ix.query_topk_dense_dev(Qd.data_ptr(), 10000, 0, 0, 10, 0, ids.data_ptr(), sc.data_ptr())
        t = ix.stage_times_ms(); s = ix.stats()
        print(f"world {a.world} rank {a.rank} NT={nt}: {ms:.3f} ms/step", {k: round(v, 3) for k, v in t.items() if v}, "survivors/q", round(s["bm_survivors"] / 10000, 1), "direct", s["bm_direct"], flush=True)
