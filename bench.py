#!/usr/bin/env python
"""bench.py — headline benchmark of the DPF hot path on B200 (contract: see the task statement / DESIGN.md).

Metric (BASELINE.json): batch kNN queries/s at fixed recall@10 (+ index build vectors/s as extra keys).
Workload at every N: BASELINE.json configs[1] — 1M x 128-d synthetic dense vectors (SIFT shape), 10k-query
batch, k=10, steps=0, reference test defaults (L=30 tables, k=32 bits, 8 sub-indexes, T=500).
A "step" = one pass of the whole query batch through hash -> probe -> de-dup -> gather/re-rank -> top-k
(for N>1: + NCCL all-gather of the per-GPU top-k and the merge kernel).

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # the CPU restatement of the reference on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# ncu figures of the scoring kernels (dram__bytes_read.sum + dram__bytes_write.sum per launch, tensor-pipe and L2
# percentages) are NOT typed in here: tools/ncu_traffic.py extracts them from an `ncu --set full` capture of this
# workload into profiles/r2_ncu_kernels.json, with the command and commit it was taken at; a kernel without a capture
# of the same workload reports null
NCU_FILE = os.path.join(ROOT, "profiles", "r2_ncu_kernels.json")


def ncu_record(kernel, workload_key):
    try:
        rec = json.load(open(NCU_FILE))
    except Exception:
        return None
    if rec.get("workload_key") != workload_key:
        return None
    k = rec.get("kernels", {}).get(kernel)
    if k is not None:
        k = dict(k, commit=rec.get("commit"), command=rec.get("command"))
    return k

METRIC_NAME = "batch kNN queries/sec at fixed recall@10"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--topk", type=int, default=10)
    ap.add_argument("--qsteps", type=int, default=0, help="multi-step sub-index search radius (reference `steps`)")
    ap.add_argument("--metric", default="dot", choices=["dot", "angular", "l2"])
    ap.add_argument("--cpu-sample", type=int, default=512, help="queries in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(args):
    from similaritysearchbyrdf_b200 import synth
    t0 = time.time()
    X, Q = synth.config2(args.n, args.nq, args.d)
    A, chain = synth.angle_family(args.d, max(100, args.d), 10, 3, 32, 88387 + 2)
    Ap = synth.partitioner_family(chain.shape[0], 3, 88387 + 3)
    return X, Q, A, chain, Ap, time.time() - t0


def config_dict(args, n_gpus):
    return {
        "workload": "configs[1]: 1M x 128-d synthetic dense (SIFT shape), 10k-query batch, k=10",
        "n_vectors": args.n, "dim": args.d, "n_queries": args.nq, "topk": args.topk, "steps": args.qsteps,
        "rerank_metric": args.metric, "tables": 30, "chain_length": 32, "partition_bits": 3, "bucket_overflow": 500,
        "dir_node_size": 32, "probe": "dense multi-probe",
        "partitioning": f"8 sub-indexes per table dealt to {n_gpus} GPUs by occupancy, vectors replicated" if n_gpus > 1 else "single GPU",
        "cache": "no explicit flush: a step reads the 128 MB byte store, 120 MB of bucket ids, 8 MB of forest nodes and ~0.3 GB of "
                 "per-batch scratch (> 126 MB L2); rows are re-staged many times inside a step by design",
    }


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's algorithm on the host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_arm(args, X, Q, A, chain, Ap, sample, steps, warmup, threads=None):
    """Builds the full index with the CPU oracle, then times `steps` passes over a bounded query sample.  `threads`:
    None = every host core; 5 = the reference's own insertThreadNum / queryThreadNum (table-sliced workers,
    DensevectorRDFInit.scala:183-188, 344-349; src/test/scala/mclab/TestSettings.scala:43-44)."""
    from oracle import oracle_py as O
    metric = {"dot": 0, "angular": 1, "l2": 2}[args.metric]
    cores = threads or os.cpu_count() or 1
    o = O.Oracle(d=args.d, L=chain.shape[0], k=chain.shape[1], P=A.shape[0], pb=Ap.shape[1])
    o.set_family(A, chain)
    o.set_partitioners(Ap)
    t0 = time.time()
    o.fit_dense(X, nthreads=cores)
    build_s = time.time() - t0
    Qs = np.ascontiguousarray(Q[:sample])
    for _ in range(warmup):
        o.query_topk_dense(Qs[: max(8, sample // 8)], None, args.qsteps, args.topk, metric, nthreads=cores)
    t0 = time.time()
    for _ in range(steps):
        ids, _ = o.query_topk_dense(Qs, None, args.qsteps, args.topk, metric, nthreads=cores)
    dt = (time.time() - t0) / steps
    _, sc = o.query_topk_dense(Qs, None, args.qsteps, args.topk, metric, nthreads=cores)
    return {"qps": sample / dt, "ms_per_step": dt * 1e3, "build_vectors_per_s": len(X) / build_s, "cores": cores,
            "sample": f"{sample} of the {len(Q)} queries per step against the full {len(X)}-vector index "
                      f"(index built by the same oracle in {build_s:.1f} s)", "ids": ids, "scores": sc}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    X, Q, A, chain, Ap, _ = workload(args)
    r = cpu_arm(args, X, Q, A, chain, Ap, min(args.cpu_sample, args.nq), args.steps, min(args.warmup, 1))
    r5 = cpu_arm(args, X, Q, A, chain, Ap, min(args.cpu_sample, args.nq) // 4, 1, 0, threads=5)
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": r["qps"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, 1),
        "cpu_baseline": {"value": r["qps"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                         "reference_thread_counts": {"value": r5["qps"], "cores": 5, "build_vectors_per_s": r5["build_vectors_per_s"],
                                                     "note": "insertThreadNum = queryThreadNum = 5, table-sliced as in the reference"},
                         "note": "C++ restatement of the reference algorithm (no JVM in the image); omits the "
                                 "reference's (de)serialisation/boxing, so it is faster than the JVM path"},
        "e2e": {"value": r["qps"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "build_vectors_per_s": r["build_vectors_per_s"], "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_id):
        self.gpu_id, self.rows, self.proc = gpu_id, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_id), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from similaritysearchbyrdf_b200 import DPFIndex
    from similaritysearchbyrdf_b200 import _lib as B

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created: point fd 1 at stderr until the first
        # collective is through, so that stdout carries the one JSON line only
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    metric = {"dot": B.METRIC_DOT, "angular": B.METRIC_ANGULAR, "l2": B.METRIC_L2}[args.metric]

    X, Q, A, chain, Ap, gen_s = workload(args)
    n, nq, d, K = args.n, args.nq, args.d, args.topk
    stream = torch.cuda.Stream(device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        Xd = torch.from_numpy(X).to(dev)
        Qd = torch.from_numpy(Q).to(dev)
        ix = DPFIndex(d=d, L=chain.shape[0], k=chain.shape[1], pb=Ap.shape[1], device=local, rank=rank, world=world)
        ix.set_balanced_partition(world > 1)         # sub-indexes dealt to the GPUs by occupancy (same split on every rank)
        ix.set_family(A, chain)
        ix.set_partitioners(Ap)
        ix.set_stream(stream.cuda_stream)
        ix.set_profiling(True)

        def join_comm(index):
            # the library's own NCCL plane: rank 0 draws the id, torch.distributed only carries it to the other ranks
            uid = torch.zeros(B.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
            if rank == 0:
                uid.copy_(torch.from_numpy(DPFIndex.comm_unique_id()))
            dist.broadcast(uid, 0)
            index.comm_init(uid.cpu().numpy())

        # warm-up build (separate handle, same size): CUDA lazy module loading happens here and the device memory
        # pool the library allocates from (cudaMallocAsync) grows to its working size, as in a long-lived process
        for _ in range(3):                       # (the pool's best-fit reuse settles after two builds of the same size)
            wix = DPFIndex(d=d, L=chain.shape[0], k=chain.shape[1], pb=Ap.shape[1], device=local, rank=rank, world=world)
            wix.set_balanced_partition(world > 1)
            wix.set_family(A, chain)
            wix.set_partitioners(Ap)
            wix.set_stream(stream.cuda_stream)
            if world > 1 and _ == 2:                 # the last warm-up goes through the sharded build (NCCL channels warm)
                join_comm(wix)
                wix.fit_dense_sharded_dev(Xd.data_ptr(), n)
                wix.sync()
                wix.comm_destroy()
            else:
                wix.fit_dense_dev(Xd.data_ptr(), n)
            wix.close()
        if world > 1:
            join_comm(ix)
        # ---- index build (inputs resident in HBM), device-timed ------------------------------------------------
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if world > 1:
            ix.fit_dense_sharded_dev(Xd.data_ptr(), n)     # every rank hashes n / world vectors, one all-gather of the keys
        else:
            ix.fit_dense_dev(Xd.data_ptr(), n)
        e1.record(stream)
        torch.cuda.synchronize()
        tb = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        build_ms = float(tb.item())
        build_stage = ix.stage_times_ms()
        bstats = ix.stats()

        ids_d = torch.empty((nq, K), dtype=torch.int32, device=dev)
        sc_d = torch.empty((nq, K), dtype=torch.float64, device=dev)
        m_ids, m_sc = ids_d, sc_d

        def step_device():
            if world > 1:     # one C-ABI call: local top k -> the library's NCCL all-gather -> merge, no host sync
                ix.query_topk_dense_all_dev(Qd.data_ptr(), nq, 0, args.qsteps, K, metric, ids_d.data_ptr(), sc_d.data_ptr())
            else:
                ix.query_topk_dense_dev(Qd.data_ptr(), nq, 0, args.qsteps, K, metric, ids_d.data_ptr(), sc_d.data_ptr())

        # ---- value: device-resident steps ---------------------------------------------------------------------
        uuid = getattr(torch.cuda.get_device_properties(dev), "uuid", None)
        clocks = ClockSampler(f"GPU-{uuid}" if uuid is not None else local)
        clocks.start()                          # sampled from the warm-up steps to the end of the stage pass: the same
        # steps throughout, so every sample is "under load".  The number of warm-up steps is the same on every rank (the
        # steps are collective at N > 1): >= W, and enough of them (~0.5 s) for nvidia-smi to take a few samples
        for nw in range(max(args.warmup, 200)):
            step_device()
            if nw % 8 == 7:
                stream.synchronize()
        launches0 = ix.stats()["kernel_launches"]
        barrier()
        ix.set_profiling(False)                # the timed region runs without per-stage events and without any host sync
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
        e1.record(stream)
        barrier()
        dt_ms = e0.elapsed_time(e1)
        launches_total = ix.stats()["kernel_launches"] - launches0          # this library's kernels inside the timed region
        # per-stage device times from a separate pass of the same steps (CUDA events on the launching stream around each
        # stage; reading them synchronises, which is why they are not taken inside the timed region)
        ix.set_profiling(True)
        rerank_ms, stage_acc = [], {}
        for _ in range(args.steps):
            step_device()
            st = ix.stage_times_ms()
            rerank_ms.append(st["rerank"])
            for k_, v in st.items():
                stage_acc[k_] = stage_acc.get(k_, 0.0) + v / args.steps
        clk = clocks.stop()
        launches = launches_total // max(args.steps, 1)
        t = torch.tensor([dt_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t.item()) / args.steps
        qstats = ix.stats()
        int_queries = bool(np.all(Q == np.floor(Q)) and Q.min() >= 0 and Q.max() <= 255)
        unit_rec_bytes = 16 + 16 * 4 + 32 * 4                                    # sizeof(UnitRec), rerank_units.cuh
        result_ids = (m_ids if world > 1 else ids_d).cpu().numpy()
        result_sc = (m_sc if world > 1 else sc_d).cpu().numpy()

        # ---- the same batch on the FP64 rows only (what real-valued data take): second index, same timing recipe ------
        f64 = None
        if world == 1 and qstats["store_kind"] != B.STORE_KIND_F64:
            ixf = DPFIndex(d=d, L=chain.shape[0], k=chain.shape[1], pb=Ap.shape[1], device=local)
            ixf.set_store_mode(B.STORE_F64_ONLY)
            ixf.set_family(A, chain)
            ixf.set_partitioners(Ap)
            ixf.set_stream(stream.cuda_stream)
            ixf.fit_dense_dev(Xd.data_ptr(), n)
            idf = torch.empty((nq, K), dtype=torch.int32, device=dev)
            scf = torch.empty((nq, K), dtype=torch.float64, device=dev)
            for _ in range(max(args.warmup, 3)):
                ixf.query_topk_dense_dev(Qd.data_ptr(), nq, 0, args.qsteps, K, metric, idf.data_ptr(), scf.data_ptr())
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(args.steps):
                ixf.query_topk_dense_dev(Qd.data_ptr(), nq, 0, args.qsteps, K, metric, idf.data_ptr(), scf.data_ptr())
            e1.record(stream)
            torch.cuda.synchronize()
            msf = e0.elapsed_time(e1) / args.steps
            ixf.set_profiling(True)
            ixf.query_topk_dense_dev(Qd.data_ptr(), nq, 0, args.qsteps, K, metric, idf.data_ptr(), scf.data_ptr())
            stf = ixf.stage_times_ms()
            f64 = {"value": nq / (msf * 1e-3), "ms_per_step": msf, "kernel": "k_score_stream", "kernel_ms": stf["rerank"],
                   "stage_ms": {k_: v for k_, v in stf.items() if v},
                   "ids_equal_to_byte_store": bool(np.array_equal(idf.cpu().numpy(), result_ids)),
                   "scores_equal_to_byte_store": bool(np.array_equal(scf.cpu().numpy(), result_sc))}
            ixf.close()
            del idf, scf

        # ---- cross-check outside the timed region: the row-major gather/re-rank kernel over the same batch --------
        # (per-candidate-row kernel of DESIGN 4; gives the unique-candidate count the roofline is quoted on, its own
        # HBM rate, and a full-size parity check of the bucket-major path: ids equal, scores within 1e-12 relative)
        ix.set_debug_option(B.DBG_RERANK, 1)
        for _ in range(2):
            step_device()
        st = ix.stage_times_ms()
        rstats = ix.stats()
        ix.set_debug_option(B.DBG_RERANK, 0)
        rm_ids = (m_ids if world > 1 else ids_d).cpu().numpy()
        rm_sc = (m_sc if world > 1 else sc_d).cpu().numpy()
        ok = rm_ids == result_ids
        close = np.abs(rm_sc - result_sc) <= 1e-12 * np.maximum(np.abs(rm_sc), 1e-300)
        rowmajor = {"kernel": "k_rerank_units", "kernel_ms": st["rerank"], "expand_ms": st["expand"],
                    "unique_candidates": int(rstats["last_candidates"]),
                    "achieved_gbs": int(rstats["last_candidates"]) * (8 * d + 4) / (st["rerank"] * 1e-3) / 1e9
                    if st["rerank"] else None,
                    "ids_equal_frac": float(ok.mean()), "scores_within_1e-12_frac": float(close[ok].mean()) if ok.any() else None}

        # ---- e2e: the reference-facing call with HOST buffers (pinned), copies inside the timed region ----------
        Qh = torch.from_numpy(Q).pin_memory()
        ids_h = torch.empty((nq, K), dtype=torch.int32).pin_memory()
        sc_h = torch.empty((nq, K), dtype=torch.float64).pin_memory()
        Qh_np, ids_np, sc_np = Qh.numpy(), ids_h.numpy(), sc_h.numpy()

        def step_e2e():      # the reference-facing call, host buffers in and out, at every N
            fn = ix.lib.dpf_query_topk_dense if world == 1 else ix.lib.dpf_query_topk_dense_all
            ix._ck(fn(ix.h, Qh_np.ctypes.data, nq, None, args.qsteps, B.PROBE_DENSE, K, metric, ids_np.ctypes.data,
                      sc_np.ctypes.data))

        ix.set_profiling(False)                # (the stage pass above left the per-stage events on: not part of the product call)
        for _ in range(max(3, args.warmup)):
            step_e2e()
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step_e2e()
        e1.record(stream)
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_ms = float(te.item()) / args.steps

        # ---- e2e build through the host API (fresh index each time, pinned host buffer -> H2D inside the timed region)
        # first pass = cold (the library's device memory pool grows to hold the vector store), then 3 warm passes
        e2e_build_ms, e2e_build_cold_ms = None, None
        if world == 1:
            Xh = torch.from_numpy(X).pin_memory()
            times = []
            for _ in range(4):
                ix2 = DPFIndex(d=d, L=chain.shape[0], k=chain.shape[1], pb=Ap.shape[1], device=local)
                ix2.set_family(A, chain)
                ix2.set_partitioners(Ap)
                ix2.set_stream(stream.cuda_stream)
                torch.cuda.synchronize()
                e0.record(stream)
                ix2.fit_dense(Xh.numpy())
                e1.record(stream)
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
                ix2.close()
            e2e_build_cold_ms, e2e_build_ms = times[0], float(np.mean(times[1:]))
            del Xh

        # ---- recall@10 against exact FP64 brute force (outside every timed region) -----------------------------
        recall = None
        if rank == 0:
            hits = 0
            for s in range(0, nq, 500):
                qb = Qd[s:s + 500]
                if args.metric == "l2":
                    sc = -(torch.cdist(qb, Xd) ** 2)
                else:
                    sc = qb @ Xd.T
                    if args.metric == "angular":
                        sc = sc / (Xd.norm(dim=1)[None, :] * qb.norm(dim=1)[:, None])
                gt = sc.topk(K, dim=1).indices.cpu().numpy()
                got = result_ids[s:s + 500]
                hits += sum(len(set(gt[i]) & set(got[i])) for i in range(len(gt)))
                del sc
            recall = hits / (nq * K)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------
    # The re-rank is bucket-major: every leaf bucket probed by the batch is scored once per unit of <= 16 queries that
    # probe it.  `achieved` = the bytes one launch of the scoring kernel has to move for that: staged rows x
    # (row bytes of the compact store + 4 B id) + one record per unit + the query operand of each unit + 16 B per
    # survivor (or 8 B per score on the dense FP64 pipeline), divided by the kernel's time from CUDA events.
    # `survey_8d` restates SURVEY 8(d)'s per-candidate figure (nC_q x (8d + 4) B per query, FP64 rows fetched once per
    # (query, unique candidate)) for comparison: it is far above the HBM peak precisely because the kernel neither
    # fetches a row once per query nor keeps it in FP64.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    ncand_unique = int(rowmajor["unique_candidates"])
    rr_ms = float(np.mean(rerank_ms)) if rerank_ms else None
    bm = qstats["bm_pairs"] > 0
    store_kind = {0: "f64", 1: "f32", 2: "u8"}[int(qstats["store_kind"])]
    row_bytes = int(qstats["store_row_bytes"])
    filtered = bm
    units, rows_staged, survivors = int(qstats["bm_runs"]), int(qstats["bm_rows_staged"]), int(qstats["bm_survivors"])
    if bm:
        kernel = ("k_score_u8s" if int_queries else "k_score_u8d") if store_kind == "u8" else "k_score_stream"
        q_operand = 16 * (128 if (store_kind == "u8" and int_queries) else 8 * d)
        # bytes the kernel REQUESTS per launch (every staged row, its id, the unit records, the query operands, the survivors)
        requested = rows_staged * (row_bytes + 4) + units * (unit_rec_bytes + q_operand) + survivors * 16
        # bytes it MUST move per launch = the algorithmic bytes of the roofline: every row of the store and every entry of
        # ids_sorted once (a row staged again for another unit is a re-read the L2 is there to absorb), the unit records,
        # the batch's queries once, the survivors out
        entries = int(chain.shape[0]) * n if world == 1 else None
        algorithmic = (n * row_bytes + (entries or rows_staged) * 4 + units * unit_rec_bytes + nq * (128 if store_kind == "u8" else 8 * d)
                       + survivors * 16)
    else:
        kernel = "k_rerank_units"
        requested = algorithmic = ncand_unique * (8 * d + 4)
    workload_key = f"configs[1] n={n} d={d} nq={nq} k={K} steps={args.qsteps} metric={args.metric} gpus={world}"
    ncu = ncu_record(kernel, workload_key)
    traffic = ncu["dram_bytes"] if ncu else None
    achieved = algorithmic / (rr_ms * 1e-3) / 1e9 if rr_ms else None
    survey_bytes = ncand_unique * (8 * d + 4)
    roofline = {"kernel": kernel, "bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": achieved / peak_gbs if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algorithmic, "kernel_ms": rr_ms, "step_share": rr_ms / ms_per_step if rr_ms else None,
                "note": "achieved = algorithmic (compulsory) bytes / CUDA-event time of the scoring kernel: the kernel is not bound by "
                        "compulsory HBM bytes — every bucket row is staged once per unit of <= 16 queries that probe its bucket and the "
                        "128 MB byte store is mostly L2-resident, so the binding resource is the gather path from L2 to the SMs: "
                        "l2_to_sm_gbs (ncu l1tex__m_xbar2l1tex_read_bytes / CUDA-event time) sits at the rate of the measured device copy; "
                        "17 % fewer instructions and whole-line copy requests left the time unchanged (DESIGN.md section 4)",
                "requested_bytes_per_launch": requested, "requested_gbs": requested / (rr_ms * 1e-3) / 1e9 if rr_ms else None,
                "dram_gbs": traffic / (rr_ms * 1e-3) / 1e9 if traffic and rr_ms else None,
                "dram_frac": traffic / (rr_ms * 1e-3) / 1e9 / peak_gbs if traffic and rr_ms else None,
                "l2_to_sm_gbs": ncu["l2_to_sm_bytes"] / (rr_ms * 1e-3) / 1e9 if ncu and ncu.get("l2_to_sm_bytes") and rr_ms else None,
                "l2_to_sm_frac_of_copy_peak": (ncu["l2_to_sm_bytes"] / (rr_ms * 1e-3) / 1e9 / peak_gbs
                                               if ncu and ncu.get("l2_to_sm_bytes") and rr_ms else None),
                "ncu": ncu,
                "store_kind": store_kind, "store_row_bytes": row_bytes, "rows_staged_per_launch": rows_staged,
                "units_per_launch": units, "survivors_per_query": survivors / nq, "queries_answered_exhaustively": int(qstats["bm_direct"]),
                "unique_candidates_per_query": ncand_unique / nq,
                "survey_8d": {"algorithmic_bytes_per_launch": survey_bytes,
                              "gbs": survey_bytes / (rr_ms * 1e-3) / 1e9 if rr_ms else None,
                              "frac_of_peak": survey_bytes / (rr_ms * 1e-3) / 1e9 / peak_gbs if rr_ms else None,
                              "note": "SURVEY 8(d)'s per-candidate figure nC_q x (8d + 4) B: what a row-major FP64 gather would move"},
                "row_major_kernel": rowmajor}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_arm(args, X, Q, A, chain, Ap, min(args.cpu_sample, nq), 1, 1)
        r5 = cpu_arm(args, X, Q, A, chain, Ap, min(args.cpu_sample, nq) // 4, 1, 0, threads=5)
        m = len(r["ids"])
        ids_equal = bool(np.array_equal(r["ids"], result_ids[:m]))
        sc_close = bool(np.all(np.abs(r["scores"] - result_sc[:m]) <= 1e-12 * np.maximum(np.abs(r["scores"]), 1e-300)))
        if not (ids_equal and sc_close):
            raise SystemExit("bench.py: the GPU top-k differs from the oracle's on the CPU-baseline sample")
        cpu = {"value": r["qps"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "build_vectors_per_s": r["build_vectors_per_s"],
               "reference_thread_counts": {"value": r5["qps"], "cores": 5, "build_vectors_per_s": r5["build_vectors_per_s"],
                                           "note": "insertThreadNum = queryThreadNum = 5, table-sliced as in the reference "
                                                   "(DensevectorRDFInit.scala:183-188, 344-349)"},
               "gpu_topk_ids_equal_oracle": ids_equal, "gpu_topk_scores_within_1e-12": sc_close}

    line = {
        "metric": METRIC_NAME, "value": nq / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None,
        # what the scoring kernel multiplies: exact u8 x u8 -> s32 on the byte copy of the FP64 store (scores are the exact
        # FP64 dot products), FP64 otherwise; `value_f64_store` is the same batch with the byte copy switched off
        "dtype": "u8" if (store_kind == "u8" and int_queries) else "f64", "value_f64_store": f64["value"] if f64 else None,
        "f64_store": f64, "data": "synthetic", "config": config_dict(args, world),
        "recall_at_10": recall, "clocks": clk,
        "e2e": {"value": nq / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(nq * d * 8),
                "d2h_bytes_per_step": int(nq * K * 12), "ms_per_step": e2e_ms},
        "gpu_launches": int(launches_total), "gpu_launches_per_step": int(launches),
        "roofline": roofline, "cpu_baseline": cpu,
        "build": {"vectors_per_s": n / (build_ms * 1e-3), "ms": build_ms, "stage_ms": build_stage,
                  "e2e_vectors_per_s": n / (e2e_build_ms * 1e-3) if e2e_build_ms else None,
                  "e2e_ms": e2e_build_ms, "e2e_cold_ms": e2e_build_cold_ms, "e2e_h2d_bytes": int(n * d * 8),
                  "near_zero_fixups": bstats["near_zero_fixups"], "splits": bstats["splits"],
                  "singleton_splits": bstats["singleton_splits"], "dir_nodes": bstats["dir_nodes"]},
        "query_stage_ms": stage_acc, "candidates_with_dups_per_query": qstats["last_cand_with_dups"] / nq,
        "store": {"kind": store_kind, "row_bytes": row_bytes},
        "datagen_s": gen_s,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
