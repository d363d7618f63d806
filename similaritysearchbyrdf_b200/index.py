"""DPFIndex — thin Python handle over the C ABI (include/dpf.h).  All compute happens in libdpf_b200.so on the GPU;
nothing here falls back to the CPU."""
import ctypes as C

import numpy as np

from . import _lib as B


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


class DPFIndex:
    """One forest of L tables on one GPU.

    Parameters mirror the reference configuration keys (src/test/scala/mclab/TestSettings.scala:9-45):
    d=mclab.lsh.vectorDim, L=tableNum*permutationNum, k=lshTable.chainLength, pb=lsh.partitionBits,
    bucket_bits=lshTable.bucketBits, dir_node_size=lshTable.dirNodeSize, bucket_overflow=lshTable.bufferOverflow.
    """

    def __init__(self, d, L, k=32, pb=3, bucket_bits=28, dir_node_size=32, bucket_overflow=500,
                 family_kind=B.FAMILY_ANGLE, key_transform=B.KEY_ORIGINAL, self_exclude_small_ids=1, device=0, rank=0,
                 world=1):
        self.lib = B.load()
        self.cfg = B.Config(B.ABI_VERSION, device, d, L, k, pb, bucket_bits, dir_node_size, bucket_overflow,
                            family_kind, key_transform, self_exclude_small_ids, rank, world)
        self.h = C.c_void_p()
        rc = self.lib.dpf_create(C.byref(self.cfg), C.byref(self.h))
        if rc != B.OK:
            raise B.DpfError(rc, self.lib.dpf_strerror(rc).decode())
        self.d, self.L, self.k, self.pb = d, L, k, pb

    # ---- persist / reload ------------------------------------------------------------------------------------
    def save(self, path):
        """Write configuration, hash functions, vectors and keys to `path` (dpf_save)."""
        self._ck(self.lib.dpf_save(self.h, str(path).encode()))

    @classmethod
    def load(cls, path, device=0):
        """Restore an index written by save() on `device`; the forest is re-created from the stored keys."""
        self = cls.__new__(cls)
        self.lib = B.load()
        self.h = C.c_void_p()
        rc = self.lib.dpf_load(str(path).encode(), device, C.byref(self.h))
        if rc != B.OK:
            raise B.DpfError(rc, self.lib.dpf_strerror(rc).decode())
        import struct
        with open(path, "rb") as f:
            f.read(8)
            vals = struct.unpack("14i", f.read(56))
        self.cfg = B.Config(*vals)
        self.cfg.device = device
        self.d, self.L, self.k, self.pb = self.cfg.d, self.cfg.L, self.cfg.k, self.cfg.pb
        return self

    # ---- plumbing -------------------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != B.OK:
            raise B.DpfError(rc, self.lib.dpf_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None) and self.h:
            self.lib.dpf_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(self.lib.dpf_size(self.h))

    def sync(self):
        self._ck(self.lib.dpf_sync(self.h))

    def set_stream(self, cuda_stream):
        """Run this index's work on the caller's CUDA stream (an integer cudaStream_t, e.g.
        torch.cuda.current_stream().cuda_stream); 0/None restores the index's own stream."""
        self._ck(self.lib.dpf_set_stream(self.h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def set_store_mode(self, mode):
        """B.STORE_AUTO (default): keep a lossless uint8 copy of the dense store for the re-rank kernels when every value
        is a byte; B.STORE_NARROWEST: uint8, else float32 when lossless; B.STORE_F64_ONLY: always read the FP64 rows.
        Call before fit."""
        self._ck(self.lib.dpf_set_store_mode(self.h, mode))

    def set_debug_option(self, option, value):
        """Test / profiling hook (B.DBG_*): selects a kernel variant on this handle; the library reads no environment."""
        self._ck(self.lib.dpf_set_debug_option(self.h, option, int(value)))

    def debug_options(self, **kw):
        """Context manager: set B.DBG_<NAME>=value for the duration of a with-block, then restore the default."""
        import contextlib

        @contextlib.contextmanager
        def cm():
            keys = {getattr(B, "DBG_" + k.upper()): v for k, v in kw.items()}
            for k, v in keys.items():
                self.set_debug_option(k, v)
            try:
                yield self
            finally:
                for k in keys:
                    self.set_debug_option(k, B.DBG_DEFAULTS.get(k, 0))
        return cm()

    def set_balanced_partition(self, on=True):
        """Multi-GPU: deal the (table, sub-index) cells to the ranks by occupancy at the first fit instead of p % world."""
        self._ck(self.lib.dpf_set_balanced_partition(self.h, 1 if on else 0))

    def owned_subindexes(self):
        """Flags (uint8, L x 2^pb) of the (table, sub-index) cells this handle owns."""
        out = np.zeros((self.L, 1 << self.pb), np.uint8)
        self._ck(self.lib.dpf_owned_subindexes(self.h, _p(out)))
        return out

    # ---- multi-GPU plane (NCCL inside the library) ------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        """Rank 0: the id every rank passes to comm_init (bytes of length B.COMM_ID_BYTES)."""
        lib = B.load()
        try:                       # in a PyTorch process the library must bind the NCCL torch ships, so torch loads it first
            import torch           # noqa: F401
        except ImportError:
            pass
        buf = np.zeros(B.COMM_ID_BYTES, np.uint8)
        rc = lib.dpf_comm_unique_id(_p(buf))
        if rc != B.OK:
            raise B.DpfError(rc, lib.dpf_strerror(rc).decode())
        return buf

    def comm_init(self, unique_id):
        """Collective over the `world` ranks of this index's configuration."""
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        assert uid.size == B.COMM_ID_BYTES
        self._ck(self.lib.dpf_comm_init(self.h, _p(uid)))

    def comm_destroy(self):
        self._ck(self.lib.dpf_comm_destroy(self.h))

    def fit_dense_sharded(self, X):
        """Collective fit: same X on every rank, hashing split across the ranks, keys all-gathered."""
        X = _f64(X)
        self._ck(self.lib.dpf_fit_dense_sharded(self.h, _p(X), X.shape[0]))

    def fit_dense_sharded_dev(self, dev_ptr, n):
        self._ck(self.lib.dpf_fit_dense_sharded_dev(self.h, C.c_void_p(dev_ptr), n))

    def query_topk_dense_all(self, Q, qids=None, steps=0, topk=10, metric=B.METRIC_DOT, probe_mode=B.PROBE_DENSE):
        """Collective query: same Q on every rank, the merged global top k on every rank."""
        Q = _f64(Q)
        qids = None if qids is None else _i32(qids)
        nq = Q.shape[0]
        ids = np.empty((nq, topk), np.int32)
        sc = np.empty((nq, topk), np.float64)
        self._ck(self.lib.dpf_query_topk_dense_all(self.h, _p(Q), nq, _p(qids), steps, probe_mode, topk, metric, _p(ids),
                                                   _p(sc)))
        return ids, sc

    def query_topk_dense_all_dev(self, q_ptr, nq, qids_ptr, steps, topk, metric, ids_ptr, score_ptr,
                                 probe_mode=B.PROBE_DENSE):
        self._ck(self.lib.dpf_query_topk_dense_all_dev(self.h, C.c_void_p(q_ptr), nq,
                                                       C.c_void_p(qids_ptr) if qids_ptr else None, steps, probe_mode,
                                                       topk, metric, C.c_void_p(ids_ptr), C.c_void_p(score_ptr)))

    # ---- hash functions -------------------------------------------------------------------------------------
    def set_family(self, A, chain_idx, b=None, w=None):
        A, chain_idx = _f64(A), _i32(chain_idx)
        if A.ndim != 2 or A.shape[1] != self.d or chain_idx.shape != (self.L, self.k):
            raise ValueError("A must be P x d and chain_idx L x k")
        b = None if b is None else _f64(b)
        w = None if w is None else _i32(w)
        self.P = A.shape[0]
        self._ck(self.lib.dpf_set_family(self.h, _p(A), A.shape[0], _p(chain_idx), _p(b), _p(w)))

    def set_partitioners(self, Ap, b=None, w=None):
        """Ap: L x pb x 32.  A pStable index passes the offsets b and widths w (L x pb each) of its partitioner chains
        as well (dpf_set_partitioners_pstable)."""
        Ap = _f64(Ap)
        if Ap.shape != (self.L, self.pb, 32):
            raise ValueError("Ap must be L x pb x 32")
        if b is None and w is None:
            self._ck(self.lib.dpf_set_partitioners(self.h, _p(Ap)))
            return
        b, w = _f64(b), _i32(w)
        if b.shape != (self.L, self.pb) or w.shape != (self.L, self.pb):
            raise ValueError("b and w must be L x pb")
        self._ck(self.lib.dpf_set_partitioners_pstable(self.h, _p(Ap), _p(b), _p(w)))

    # ---- hashing --------------------------------------------------------------------------------------------
    def hash_dense(self, X):
        X = _f64(X)
        n = X.shape[0]
        keys = np.empty((self.L, n), np.int32)
        pids = np.empty((self.L, n), np.int32)
        self._ck(self.lib.dpf_hash_dense(self.h, _p(X), n, _p(keys), _p(pids)))
        return keys, pids

    def hash_csr(self, indptr, indices, values):
        indptr, indices, values = _i64(indptr), _i32(indices), _f64(values)
        n = len(indptr) - 1
        keys = np.empty((self.L, n), np.int32)
        pids = np.empty((self.L, n), np.int32)
        self._ck(self.lib.dpf_hash_csr(self.h, _p(indptr), _p(indices), _p(values), n, _p(keys), _p(pids)))
        return keys, pids

    # ---- fit ------------------------------------------------------------------------------------------------
    def fit_dense(self, X):
        X = _f64(X)
        self._ck(self.lib.dpf_fit_dense(self.h, _p(X), X.shape[0]))

    def fit_dense_dev(self, dev_ptr, n):
        """Borrow an n x d FP64 device buffer (e.g. torch tensor .data_ptr()); it must outlive the index."""
        self._ck(self.lib.dpf_fit_dense_dev(self.h, C.c_void_p(dev_ptr), n))

    def fit_csr(self, indptr, indices, values):
        indptr, indices, values = _i64(indptr), _i32(indices), _f64(values)
        self._ck(self.lib.dpf_fit_csr(self.h, _p(indptr), _p(indices), _p(values), len(indptr) - 1))

    def remove(self, ids):
        """RandomDrawTreeMap.remove for a batch of ids in every table; returns the (table, id) entries removed."""
        ids = _i32(ids)
        gone = C.c_int64(0)
        self._ck(self.lib.dpf_remove(self.h, _p(ids), len(ids), C.byref(gone)))
        return int(gone.value)

    # ---- queries --------------------------------------------------------------------------------------------
    def _cand_call(self, fn, nq, *args):
        off = np.empty(nq + 1, np.int64)
        total = C.c_int64(0)
        cap = max(1 << 16, 64 * nq)
        while True:
            ids = np.empty(cap, np.int32)
            rc = fn(self.h, *args, _p(off), _p(ids), cap, C.byref(total))
            if rc == B.ERR_CAPACITY:
                cap = int(total.value)
                continue
            self._ck(rc)
            return off, ids[:total.value].copy()

    def query_candidates_dense(self, Q, qids=None, steps=0, probe_mode=B.PROBE_DENSE):
        Q = _f64(Q)
        qids = None if qids is None else _i32(qids)
        return self._cand_call(self.lib.dpf_query_candidates_dense, Q.shape[0], _p(Q), Q.shape[0], _p(qids), steps,
                               probe_mode)

    def query_candidates_csr(self, indptr, indices, values, qids=None, steps=0):
        indptr, indices, values = _i64(indptr), _i32(indices), _f64(values)
        qids = None if qids is None else _i32(qids)
        nq = len(indptr) - 1
        return self._cand_call(self.lib.dpf_query_candidates_csr, nq, _p(indptr), _p(indices), _p(values), nq,
                               _p(qids), steps)

    def query_candidates_by_id(self, qids, steps=0):
        qids = _i32(qids)
        return self._cand_call(self.lib.dpf_query_candidates_by_id, len(qids), _p(qids), len(qids), steps)

    def query_topk_dense(self, Q, qids=None, steps=0, topk=10, metric=B.METRIC_DOT, probe_mode=B.PROBE_DENSE):
        Q = _f64(Q)
        qids = None if qids is None else _i32(qids)
        nq = Q.shape[0]
        ids = np.empty((nq, topk), np.int32)
        sc = np.empty((nq, topk), np.float64)
        self._ck(self.lib.dpf_query_topk_dense(self.h, _p(Q), nq, _p(qids), steps, probe_mode, topk, metric, _p(ids),
                                               _p(sc)))
        return ids, sc

    def query_topk_dense_dev(self, q_ptr, nq, qids_ptr, steps, topk, metric, ids_ptr, score_ptr,
                             probe_mode=B.PROBE_DENSE):
        """Device-resident variant: all pointers are device addresses on this index's GPU (asynchronous)."""
        self._ck(self.lib.dpf_query_topk_dense_dev(self.h, C.c_void_p(q_ptr), nq,
                                                   C.c_void_p(qids_ptr) if qids_ptr else None, steps, probe_mode,
                                                   topk, metric, C.c_void_p(ids_ptr), C.c_void_p(score_ptr)))

    def rerank_dense(self, Q, offsets, cand, topk, metric=B.METRIC_DOT):
        Q, offsets, cand = _f64(Q), _i64(offsets), _i32(cand)
        nq = Q.shape[0]
        ids = np.empty((nq, topk), np.int32)
        sc = np.empty((nq, topk), np.float64)
        self._ck(self.lib.dpf_rerank_dense(self.h, _p(Q), nq, _p(offsets), _p(cand), topk, metric, _p(ids), _p(sc)))
        return ids, sc

    def merge_topk_dev(self, gids_ptr, gsc_ptr, G, nq, topk, metric, ids_ptr, score_ptr):
        self._ck(self.lib.dpf_merge_topk_dev(self.h, C.c_void_p(gids_ptr), C.c_void_p(gsc_ptr), G, nq, topk, metric,
                                             C.c_void_p(ids_ptr), C.c_void_p(score_ptr)))

    # ---- introspection --------------------------------------------------------------------------------------
    def dump_buckets(self, table):
        nb, nid = C.c_int64(0), C.c_int64(0)
        self._ck(self.lib.dpf_dump_buckets(self.h, table, C.byref(nb), C.byref(nid), None, None, None))
        desc = np.empty((max(nb.value, 1), 3), np.int32)
        off = np.empty(nb.value + 1, np.int64)
        ids = np.empty(max(nid.value, 1), np.int32)
        self._ck(self.lib.dpf_dump_buckets(self.h, table, C.byref(nb), C.byref(nid), _p(desc), _p(off), _p(ids)))
        return desc[:nb.value], off, ids[:nid.value]

    def leaf_pairs(self):
        """(pair_off[nleaves + 1], leaf_len[nleaves]) of the last bucket-major query chunk (dpf_debug_leaf_pairs)."""
        n = C.c_int64(0)
        self._ck(self.lib.dpf_debug_leaf_pairs(self.h, C.byref(n), None, None))
        off = np.zeros(n.value + 1, np.uint32)
        ln = np.zeros(max(n.value, 1), np.int32)
        self._ck(self.lib.dpf_debug_leaf_pairs(self.h, C.byref(n), _p(off), _p(ln)))
        return off, ln[:n.value]

    def tc_diag(self):
        """Watchdog record of the tcgen05 scoring kernel (all zero in a correct run)."""
        out = np.zeros(24, np.uint64)
        self._ck(self.lib.dpf_debug_tc_diag(self.h, _p(out)))
        return out

    def stats(self):
        s = np.zeros(B.STAT_COUNT, np.int64)
        occ = np.zeros(1 << self.pb, np.float64)
        self._ck(self.lib.dpf_stats(self.h, _p(s), _p(occ)))
        out = {n: int(s[i]) for i, n in enumerate(B.STAT_NAMES)}
        out["occupancy"] = occ
        return out

    def set_profiling(self, on=True):
        self._ck(self.lib.dpf_set_profiling(self.h, 1 if on else 0))

    def stage_times_ms(self):
        ms = np.zeros(B.T_COUNT, np.float32)
        self._ck(self.lib.dpf_stage_times_ms(self.h, _p(ms)))
        return {n: float(ms[i]) for i, n in enumerate(B.STAGE_NAMES)}
