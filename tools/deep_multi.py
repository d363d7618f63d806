"""BASELINE configs[4]: Deep shape (n x 96-d FP64, generated on every GPU from the same seed = replicated vectors),
content-partitioned forest across the GPUs of one box (rank r owns sub-indexes p % world == r of every table),
queries replicated, per-GPU top-k merged after one NCCL all-gather.  One JSON line from rank 0.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
         tools/deep_multi.py --vectors 100000000 --queries 10000
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from similaritysearchbyrdf_b200 import DPFIndex, synth
from similaritysearchbyrdf_b200 import _lib as B

ap = argparse.ArgumentParser()
ap.add_argument("--vectors", dest="n", type=int, default=100_000_000)
ap.add_argument("--queries", dest="nq", type=int, default=10_000)
ap.add_argument("--timed-steps", dest="steps", type=int, default=3)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n, nq, d, K = a.n, a.nq, 96, 10
g = torch.Generator(device=dev); g.manual_seed(1005)
centres = torch.randn((max(1, n // 10000), d), generator=g, device=dev, dtype=torch.float64)
Xd = torch.empty((n, d), dtype=torch.float64, device=dev)
t0 = time.time()
for s in range(0, n, 1 << 20):
    m = min(1 << 20, n - s)
    c = torch.randint(0, centres.shape[0], (m,), generator=g, device=dev)
    Xd[s:s + m] = centres[c] + 0.35 * torch.randn((m, d), generator=g, device=dev, dtype=torch.float64)
c = torch.randint(0, centres.shape[0], (nq,), generator=g, device=dev)
Qd = centres[c] + 0.35 * torch.randn((nq, d), generator=g, device=dev, dtype=torch.float64)
torch.cuda.synchronize(); gen_s = time.time() - t0
A, chain = synth.angle_family(d, 100, 10, 3, 32, 88387 + 5)
Ap = synth.partitioner_family(30, 3, 88387 + 6)
stream = torch.cuda.Stream(device=dev)
with torch.cuda.stream(stream):
    build_s = []
    for rep in range(2):                 # first build grows the device memory pool (cold), second is warm
        ix = DPFIndex(d=d, L=30, k=32, pb=3, device=local, rank=rank, world=world)
        ix.set_family(A, chain); ix.set_partitioners(Ap); ix.set_stream(stream.cuda_stream); ix.set_profiling(True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(); t0 = time.time()
        ix.fit_dense_dev(Xd.data_ptr(), n)
        torch.cuda.synchronize()
        bt = torch.tensor([time.time() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(bt, op=dist.ReduceOp.MAX)
        build_s.append(float(bt.item()))
        if rep == 0:
            ix.close()
    bst, stats = ix.stage_times_ms(), ix.stats()
    ids_d = torch.empty((nq, K), dtype=torch.int32, device=dev)
    sc_d = torch.empty((nq, K), dtype=torch.float64, device=dev)
    g_ids = torch.empty((world, nq, K), dtype=torch.int32, device=dev)
    g_sc = torch.empty((world, nq, K), dtype=torch.float64, device=dev)
    m_ids = torch.empty((nq, K), dtype=torch.int32, device=dev)
    m_sc = torch.empty((nq, K), dtype=torch.float64, device=dev)

    def step():
        ix.query_topk_dense_dev(Qd.data_ptr(), nq, 0, 0, K, B.METRIC_DOT, ids_d.data_ptr(), sc_d.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(g_ids, ids_d)
            dist.all_gather_into_tensor(g_sc, sc_d)
            ix.merge_topk_dev(g_ids.data_ptr(), g_sc.data_ptr(), world, nq, K, B.METRIC_DOT, m_ids.data_ptr(), m_sc.data_ptr())

    step(); step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    qt = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(qt, op=dist.ReduceOp.MAX)
    st, s2 = ix.stage_times_ms(), ix.stats()
    res = (m_ids if world > 1 else ids_d)
    sub = min(nq, 200)
    hits = 0
    if rank == 0:
        for s in range(0, sub, 4):
            gt = (Qd[s:s + 4] @ Xd.T).topk(K, dim=1).indices.cpu().numpy()
            got = res[s:s + 4].cpu().numpy()
            hits += sum(len(set(gt[i]) & set(got[i])) for i in range(len(gt)))
free, total = torch.cuda.mem_get_info()
if rank == 0:
    rec = {"config": "configs[4] Deep shape, content-partitioned across GPUs", "n_gpus": world, "n": n, "d": d, "nq": nq, "topk": K,
           "datagen_s": round(gen_s, 1), "build_s_cold_max_over_ranks": build_s[0], "build_s_max_over_ranks": build_s[1], "build_vectors_per_s": n / build_s[1],
           "build_stage_ms_rank0": {k: round(v, 2) for k, v in bst.items() if v}, "owned_entries_rank0": int(sum(stats["occupancy"]) * 30) if "occupancy" in stats else None,
           "splits_rank0": stats["splits"], "dir_nodes_rank0": stats["dir_nodes"],
           "query_ms_per_step_max_over_ranks": float(qt.item()), "queries_per_s": nq / (float(qt.item()) * 1e-3),
           "query_stage_ms_rank0": {k: round(v, 3) for k, v in st.items() if v},
           "with_dups_per_query_rank0": s2["last_cand_with_dups"] / nq, "bm_rows_staged_rank0": s2["bm_rows_staged"],
           "recall_at_10_first_200": hits / (sub * K), "device_mem_used_gb_rank0": (total - free) / 1e9}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/config_runs.jsonl", "a") as f:
        f.write(json.dumps(rec) + "\n")
    print(json.dumps(rec), flush=True)
if world > 1:
    dist.destroy_process_group()
