"""Runs the BASELINE.json configurations that are not the bench line (configs[2], [3], and [4] on one GPU) at full
or stated size and appends one JSON line per run to gpurun_out/config_runs.jsonl: stage times, throughput,
algorithmic traffic against the measured HBM peak, recall against exact FP64 brute force where it applies.

  python tools/config_runs.py gist   [--n 1000000] [--nq 1000]
  python tools/config_runs.py sparse [--n 2000000]
  python tools/config_runs.py deep   [--n 25000000] [--nq 10000]     # vectors generated on the device
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from similaritysearchbyrdf_b200 import DPFIndex, synth
from similaritysearchbyrdf_b200 import _lib as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "config_runs.jsonl")
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


def emit(rec):
    rec["hbm_peak_gbs"] = PEAK
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "a") as f:
        f.write(json.dumps(rec) + "\n")
    print(json.dumps(rec), flush=True)


def r3(d):
    return {k: round(float(v), 3) for k, v in d.items() if v}


def mk(d, A, chain, Ap, **kw):
    ix = DPFIndex(d=d, L=chain.shape[0], k=chain.shape[1], pb=Ap.shape[1], **kw)
    ix.set_family(A, chain)
    ix.set_partitioners(Ap)
    ix.set_profiling(True)
    return ix


def recall_at(Xd, Qd, ids, K, chunk=100):
    hits = 0
    for s in range(0, Qd.shape[0], chunk):
        gt = (Qd[s:s + chunk] @ Xd.T).topk(K, dim=1).indices.cpu().numpy()
        hits += sum(len(set(gt[i]) & set(ids[s + i])) for i in range(len(gt)))
    return hits / (Qd.shape[0] * K)


def run_gist(a):
    n, nq, d, K = a.n or 1_000_000, a.nq or 1000, 960, 100
    t0 = time.time()
    X, Q = synth.config3(n, nq, d)
    A, chain = synth.angle_family(d, max(100, d), 10, 3, 32, 88387 + 3)
    Ap = synth.partitioner_family(30, 3, 88387 + 4)
    gen_s = time.time() - t0
    Xd, Qd = torch.from_numpy(X).cuda(), torch.from_numpy(Q).cuda()
    for _ in range(2):                                   # second build = warm pool
        ix = mk(d, A, chain, Ap)
        torch.cuda.synchronize(); t0 = time.time()
        ix.fit_dense_dev(Xd.data_ptr(), n)
        build_ms = 1e3 * (time.time() - t0)
        bst = ix.stage_times_ms(); stats = ix.stats()
        if _ == 0:
            ix.close()
    P = A.shape[0]
    rec = {"config": "configs[2] GIST shape", "n": n, "d": d, "nq": nq, "topk": K, "datagen_s": round(gen_s, 1),
           "build_ms": round(build_ms, 2), "build_vectors_per_s": n / (build_ms * 1e-3), "build_stage_ms": r3(bst),
           "hash_tflops": 2.0 * d * P * n / (bst["hash"] * 1e-3) / 1e12, "distinct_functions": P,
           "splits": stats["splits"], "dir_nodes": stats["dir_nodes"], "near_zero_fixups": stats["near_zero_fixups"],
           "steps": {}}
    ids_d = torch.empty((nq, K), dtype=torch.int32, device="cuda")
    sc_d = torch.empty((nq, K), dtype=torch.float64, device="cuda")
    for steps in (0, 1, 2, 3):
        rec["steps"][str(steps)] = {}
        ref_ids = None
        for mode, wide_opt in (("bucket_major_wide", 0), ("row_major", 1)):      # rerank_wide.cu against the row-major gather
            ix.set_debug_option(B.DBG_WIDE, wide_opt)
            for rep in range(3):
                torch.cuda.synchronize(); t0 = time.time()
                ix.query_topk_dense_dev(Qd.data_ptr(), nq, 0, steps, K, B.METRIC_DOT, ids_d.data_ptr(), sc_d.data_ptr())
                torch.cuda.synchronize(); ms = 1e3 * (time.time() - t0)
            st, s2 = ix.stage_times_ms(), ix.stats()
            ids_h = ids_d.cpu().numpy().copy()
            r = {"ms": round(ms, 3), "queries_per_s": nq / (ms * 1e-3), "stage_ms": r3(st),
                 "with_dups_per_query": s2["last_cand_with_dups"] / nq}
            if wide_opt == 0:
                rows = s2["bm_rows_staged"]
                r.update({"pairs": s2["bm_pairs"], "units": s2["bm_runs"], "rows_staged": rows, "survivors_per_query": s2["bm_survivors"] / nq,
                          "answered_exhaustively": s2["bm_direct"],
                          "rerank_alg_gbs": rows * 8 * d / (st["rerank"] * 1e-3) / 1e9 if st["rerank"] else None,
                          "rerank_frac_of_hbm_peak": rows * 8 * d / (st["rerank"] * 1e-3) / 1e9 / PEAK if st["rerank"] else None,
                          "rerank_tflops": 2.0 * d * 16 * rows / (st["rerank"] * 1e-3) / 1e12 if st["rerank"] else None,
                          "recall_at_100": recall_at(Xd, Qd, ids_h, K)})
                ref_ids = ids_h
            else:
                uniq = s2["last_candidates"]
                r.update({"unique_candidates_per_query": uniq / nq,
                          "rerank_alg_gbs": uniq * (8 * d + 4) / (st["rerank"] * 1e-3) / 1e9 if st["rerank"] else None,
                          "rerank_frac_of_hbm_peak": uniq * (8 * d + 4) / (st["rerank"] * 1e-3) / 1e9 / PEAK if st["rerank"] else None,
                          "ids_equal_to_bucket_major_frac": float((ids_h == ref_ids).mean())})
            rec["steps"][str(steps)][mode] = r
        ix.set_debug_option(B.DBG_WIDE, 0)
    emit(rec)


def run_sparse(a):
    n, D = a.n or 2_000_000, 100_000
    t0 = time.time()
    indptr, idx, val = synth.config4_csr_fast(n, D)
    A, chain = synth.angle_family(D, 100, 10, 3, 32, 88387 + 4)
    Ap = synth.partitioner_family(30, 3, 88387 + 5)
    gen_s = time.time() - t0
    nnz = int(indptr[-1])
    for rep in range(2):
        ix = mk(D, A, chain, Ap)
        t0 = time.time()
        ix.fit_csr(indptr, idx, val)
        fit_ms = 1e3 * (time.time() - t0)
        bst, stats = ix.stage_times_ms(), ix.stats()
        if rep == 0:
            ix.close()
    P = A.shape[0]
    alg = 12.0 * nnz + 8.0 * nnz * P + 4.0 * 30 * n
    rec = {"config": "configs[3] sparse CSR hashing", "n": n, "D": D, "nnz": nnz, "mean_nnz": nnz / n, "distinct_functions": P,
           "datagen_s": round(gen_s, 1), "fit_ms_host_api": round(fit_ms, 1), "build_stage_ms": r3(bst),
           "hash_vectors_per_s": n / (bst["hash"] * 1e-3), "hash_alg_bytes": alg,
           "hash_alg_gbs": alg / (bst["hash"] * 1e-3) / 1e9, "hash_frac_of_hbm_peak": alg / (bst["hash"] * 1e-3) / 1e9 / PEAK,
           "hash_gflops": 2.0 * nnz * P / (bst["hash"] * 1e-3) / 1e9,
           "device_build_vectors_per_s": n / (sum(bst.values()) * 1e-3), "splits": stats["splits"], "dir_nodes": stats["dir_nodes"]}
    qids = np.arange(0, n, n // 10000, dtype=np.int32)[:10000]
    for steps in (0, 1):
        for _ in range(2):
            t0 = time.time()
            off, cand = ix.query_candidates_by_id(qids, steps)
            ms = 1e3 * (time.time() - t0)
        rec[f"query_by_id_steps{steps}"] = {"nq": len(qids), "ms_host_api": round(ms, 2), "stage_ms": r3(ix.stage_times_ms()),
                                            "candidates_per_query": float(len(cand)) / len(qids),
                                            "self_found_frac": float(np.mean([qids[i] in cand[off[i]:off[i + 1]] for i in range(0, len(qids), 50)]))}
    emit(rec)


def run_deep(a):
    n, nq, d, K = a.n or 25_000_000, a.nq or 10_000, 96, 10
    g = torch.Generator(device="cuda"); g.manual_seed(1005)
    centres = torch.randn((max(1, n // 10000), d), generator=g, device="cuda", dtype=torch.float64)
    Xd = torch.empty((n, d), dtype=torch.float64, device="cuda")
    t0 = time.time()
    for s in range(0, n, 1 << 20):
        m = min(1 << 20, n - s)
        c = torch.randint(0, centres.shape[0], (m,), generator=g, device="cuda")
        Xd[s:s + m] = centres[c] + 0.35 * torch.randn((m, d), generator=g, device="cuda", dtype=torch.float64)
    c = torch.randint(0, centres.shape[0], (nq,), generator=g, device="cuda")
    Qd = centres[c] + 0.35 * torch.randn((nq, d), generator=g, device="cuda", dtype=torch.float64)
    torch.cuda.synchronize(); gen_s = time.time() - t0
    A, chain = synth.angle_family(d, 100, 10, 3, 32, 88387 + 5)
    Ap = synth.partitioner_family(30, 3, 88387 + 6)
    ix = mk(d, A, chain, Ap)
    torch.cuda.synchronize(); t0 = time.time()
    ix.fit_dense_dev(Xd.data_ptr(), n)
    build_ms = 1e3 * (time.time() - t0)
    bst, stats = ix.stage_times_ms(), ix.stats()
    rec = {"config": "configs[4] Deep shape on ONE GPU (all 8 sub-indexes)", "n": n, "d": d, "nq": nq, "topk": K,
           "datagen_s": round(gen_s, 1), "build_ms_cold": round(build_ms, 1), "build_vectors_per_s_cold": n / (build_ms * 1e-3),
           "build_stage_ms": r3(bst), "device_stage_vectors_per_s": n / (sum(bst.values()) * 1e-3),
           "splits": stats["splits"], "dir_nodes": stats["dir_nodes"], "singleton_splits": stats["singleton_splits"],
           "mem_allocated_gb": torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9}
    ids_d = torch.empty((nq, K), dtype=torch.int32, device="cuda")
    sc_d = torch.empty((nq, K), dtype=torch.float64, device="cuda")
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        ix.query_topk_dense_dev(Qd.data_ptr(), nq, 0, 0, K, B.METRIC_DOT, ids_d.data_ptr(), sc_d.data_ptr())
        torch.cuda.synchronize(); ms = 1e3 * (time.time() - t0)
    st, s2 = ix.stage_times_ms(), ix.stats()
    rec["query"] = {"ms": round(ms, 2), "queries_per_s": nq / (ms * 1e-3), "stage_ms": r3(st),
                    "with_dups_per_query": s2["last_cand_with_dups"] / nq, "bm_pairs": s2["bm_pairs"], "bm_rows_staged": s2["bm_rows_staged"],
                    "survivors_per_query": s2["bm_survivors"] / nq, "answered_exhaustively": s2["bm_direct"]}
    sub = min(nq, 200)
    rec["recall_at_10_first_200"] = recall_at(Xd, Qd[:sub], ids_d[:sub].cpu().numpy(), K, chunk=8)
    emit(rec)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["gist", "sparse", "deep"])
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--nq", type=int, default=0)
    a = ap.parse_args()
    {"gist": run_gist, "sparse": run_sparse, "deep": run_deep}[a.which](a)
