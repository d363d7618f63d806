"""ctypes binding of the CPU parity oracle (oracle/dpf_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package (similaritysearchbyrdf_b200) must never import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdpf_oracle.so")

METRIC_DOT, METRIC_ANGULAR, METRIC_L2 = 0, 1, 2
PROBE_NONE, PROBE_DENSE = 0, 1
TRANSFORMS = {"original": 0, "sampling": 1, "continueBitsCount": 2, "angleNewMethod": 3}


class Cfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "d", "L", "k", "P", "pb", "bucket_bits", "dir_node_size", "bucket_overflow", "family_kind",
        "key_transform", "self_exclude_small_ids", "rank", "world")]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("dpf_oracle.cpp", "dpf_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.dpfo_create.restype = vp
        L.dpfo_create.argtypes = [C.POINTER(Cfg)]
        L.dpfo_destroy.argtypes = [vp]
        L.dpfo_set_family.argtypes = [vp, vp, vp, vp, vp]
        L.dpfo_set_partitioners.argtypes = [vp, vp]
        L.dpfo_set_owned.argtypes = [vp, vp]
        L.dpfo_set_partitioners_pstable.argtypes = [vp, vp, vp, vp]
        L.dpfo_set_owned_cells.argtypes = [vp, vp]
        L.dpfo_hash_dense.argtypes = [vp, vp, i64, vp, vp, C.c_int]
        L.dpfo_hash_csr.argtypes = [vp, vp, vp, vp, i64, vp, vp, C.c_int]
        L.dpfo_fit_dense.argtypes = [vp, vp, i64, C.c_int]
        L.dpfo_fit_csr.argtypes = [vp, vp, vp, vp, i64, C.c_int]
        L.dpfo_size.restype = i64
        L.dpfo_size.argtypes = [vp]
        L.dpfo_remove.restype = i64
        L.dpfo_remove.argtypes = [vp, vp, i64]
        L.dpfo_query_candidates_dense.restype = i64
        L.dpfo_query_candidates_dense.argtypes = [vp, vp, i64, vp, C.c_int, C.c_int, C.c_int]
        L.dpfo_query_candidates_csr.restype = i64
        L.dpfo_query_candidates_csr.argtypes = [vp, vp, vp, vp, i64, vp, C.c_int, C.c_int]
        L.dpfo_query_candidates_by_id.restype = i64
        L.dpfo_query_candidates_by_id.argtypes = [vp, vp, i64, C.c_int, C.c_int]
        L.dpfo_get_candidates.argtypes = [vp, vp, vp]
        L.dpfo_rerank_dense.argtypes = [vp, vp, i64, vp, vp, C.c_int, C.c_int, vp, vp, C.c_int]
        L.dpfo_query_topk_dense.argtypes = [vp, vp, i64, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int]
        L.dpfo_dump_buckets.restype = i64
        L.dpfo_dump_buckets.argtypes = [vp, C.c_int, vp, vp, vp]
        L.dpfo_num_dir_nodes.restype = i64
        L.dpfo_num_dir_nodes.argtypes = [vp, C.c_int]
        L.dpfo_stats.argtypes = [vp, vp, vp]
        L.dpfo_dot_dense.restype = dbl
        L.dpfo_dot_dense.argtypes = [vp, vp, C.c_int]
        L.dpfo_dot_sparse.restype = dbl
        L.dpfo_dot_sparse.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int]
        L.dpfo_angle_key_from_dots.restype = i32
        L.dpfo_angle_key_from_dots.argtypes = [vp, C.c_int]
        L.dpfo_pstable_key_from_dots.restype = i32
        L.dpfo_pstable_key_from_dots.argtypes = [vp, vp, vp, C.c_int]
        L.dpfo_sampling_key.restype = i32
        L.dpfo_sampling_key.argtypes = [i32]
        L.dpfo_sampling_index.argtypes = [vp]
        L.dpfo_continue_bits_count.restype = i32
        L.dpfo_continue_bits_count.argtypes = [i32]
        L.dpfo_angle_new_method.restype = i32
        L.dpfo_angle_new_method.argtypes = [i32]
        L.dpfo_partition_id.restype = i32
        L.dpfo_partition_id.argtypes = [i32, vp, C.c_int, C.c_int]
        L.dpfo_partition_id_pstable.restype = i32
        L.dpfo_partition_id_pstable.argtypes = [i32, vp, C.c_int, C.c_int, vp, vp]
        L.dpfo_default_hasher.restype = i32
        L.dpfo_default_hasher.argtypes = [i32]
        L.dpfo_dir_offset_from_slot.restype = i32
        L.dpfo_dir_offset_from_slot.argtypes = [vp, C.c_int, C.c_int]
        L.dpfo_tree_params.argtypes = [C.c_int, C.c_int, C.c_int, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


# ---- known-answer hooks -------------------------------------------------------------------------------------
def dot_dense(a, x):
    a, x = _f64(a), _f64(x)
    return lib().dpfo_dot_dense(_p(a), _p(x), len(a))


def dot_sparse(ia, va, ib, vb):
    ia, va, ib, vb = _i32(ia), _f64(va), _i32(ib), _f64(vb)
    return lib().dpfo_dot_sparse(_p(ia), _p(va), len(ia), _p(ib), _p(vb), len(ib))


def angle_key_from_dots(dots):
    dots = _f64(dots)
    return lib().dpfo_angle_key_from_dots(_p(dots), len(dots))


def pstable_key_from_dots(dots, b, w):
    dots, b, w = _f64(dots), _f64(b), _i32(w)
    return lib().dpfo_pstable_key_from_dots(_p(dots), _p(b), _p(w), len(dots))


def sampling_key(key):
    return lib().dpfo_sampling_key(int(np.int32(key)))


def sampling_index():
    s = np.zeros(32, np.int32)
    lib().dpfo_sampling_index(_p(s))
    return s


def continue_bits_count(key):
    return lib().dpfo_continue_bits_count(int(np.int32(key)))


def angle_new_method(key):
    return lib().dpfo_angle_new_method(int(np.int32(key)))


def partition_id(h, Ap_t, transform=0, b=None, w=None):
    Ap_t = _f64(Ap_t)
    if b is None:
        return lib().dpfo_partition_id(int(np.int32(h)), _p(Ap_t), Ap_t.shape[0], transform)
    b, w = _f64(b), np.ascontiguousarray(w, dtype=np.int32)
    return lib().dpfo_partition_id_pstable(int(np.int32(h)), _p(Ap_t), Ap_t.shape[0], transform, _p(b), _p(w))


def default_hasher(key):
    return lib().dpfo_default_hasher(int(np.int32(key)))


def dir_offset_from_slot(bitmap, slot):
    bitmap = _i32(bitmap)
    return lib().dpfo_dir_offset_from_slot(_p(bitmap), len(bitmap), slot)


def tree_params(bucket_bits, dir_node_size, chain_length):
    out = np.zeros(4, np.int32)
    lib().dpfo_tree_params(bucket_bits, dir_node_size, chain_length, _p(out))
    return dict(SEG=int(out[0]), nb=int(out[1]), mask=int(out[2]), MAXL=int(out[3]))


class Oracle:
    """One forest (L tables) of the CPU oracle; mirrors the C-ABI handle of the product library."""

    def __init__(self, d, L, k, P, pb=3, bucket_bits=28, dir_node_size=32, bucket_overflow=500, family_kind=0,
                 key_transform=0, self_exclude_small_ids=1, rank=0, world=1):
        self.cfg = Cfg(d, L, k, P, pb, bucket_bits, dir_node_size, bucket_overflow, family_kind, key_transform,
                       self_exclude_small_ids, rank, world)
        self.h = lib().dpfo_create(C.byref(self.cfg))
        if not self.h:
            raise ValueError("dpfo_create rejected the configuration")
        self.L, self.d, self.k, self.P, self.pb = L, d, k, P, pb

    def close(self):
        if self.h:
            lib().dpfo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_family(self, A, chain_idx, b=None, w=None):
        A, chain_idx = _f64(A), _i32(chain_idx)
        assert A.shape == (self.P, self.d) and chain_idx.shape == (self.L, self.k)
        b = None if b is None else _f64(b)
        w = None if w is None else _i32(w)
        rc = lib().dpfo_set_family(self.h, _p(A), _p(chain_idx), _p(b), _p(w))
        assert rc == 0, rc

    def set_partitioners(self, Ap, b=None, w=None):
        """Ap: L x pb x 32.  With b, w (L x pb each) the partitioner chains are pStable functions (the reference builds
        the partitioner's LSH from the main family's configuration)."""
        Ap = _f64(Ap)
        assert Ap.shape == (self.L, self.pb, 32)
        if b is None:
            lib().dpfo_set_partitioners(self.h, _p(Ap))
        else:
            b, w = _f64(b), np.ascontiguousarray(w, dtype=np.int32)
            assert b.shape == (self.L, self.pb) and w.shape == (self.L, self.pb)
            rc = lib().dpfo_set_partitioners_pstable(self.h, _p(Ap), _p(b), _p(w))
            assert rc == 0, rc

    def set_owned(self, owned):
        """Explicit shard: flags of the sub-indexes (shape 2^pb: the same in every table) or of the (table, sub-index)
        cells (shape L x 2^pb) this instance owns; None = p % world == rank."""
        if owned is None:
            lib().dpfo_set_owned(self.h, None)
            return
        owned = np.ascontiguousarray(owned, dtype=np.uint8)
        if owned.ndim == 2:
            assert owned.shape == (self.L, 1 << self.pb)
            lib().dpfo_set_owned_cells(self.h, _p(owned))
        else:
            assert owned.shape == (1 << self.pb,)
            lib().dpfo_set_owned(self.h, _p(owned))

    def hash_dense(self, X, nthreads=0):
        X = _f64(X)
        n = X.shape[0]
        keys = np.empty((self.L, n), np.int32)
        pids = np.empty((self.L, n), np.int32)
        lib().dpfo_hash_dense(self.h, _p(X), n, _p(keys), _p(pids), nthreads)
        return keys, pids

    def hash_csr(self, indptr, indices, values, nthreads=0):
        indptr, indices, values = _i64(indptr), _i32(indices), _f64(values)
        n = len(indptr) - 1
        keys = np.empty((self.L, n), np.int32)
        pids = np.empty((self.L, n), np.int32)
        lib().dpfo_hash_csr(self.h, _p(indptr), _p(indices), _p(values), n, _p(keys), _p(pids), nthreads)
        return keys, pids

    def fit_dense(self, X, nthreads=0):
        X = _f64(X)
        rc = lib().dpfo_fit_dense(self.h, _p(X), X.shape[0], nthreads)
        assert rc == 0, rc

    def fit_csr(self, indptr, indices, values, nthreads=0):
        indptr, indices, values = _i64(indptr), _i32(indices), _f64(values)
        rc = lib().dpfo_fit_csr(self.h, _p(indptr), _p(indices), _p(values), len(indptr) - 1, nthreads)
        assert rc == 0, rc

    def size(self):
        return lib().dpfo_size(self.h)

    def remove(self, ids):
        ids = _i32(ids)
        return lib().dpfo_remove(self.h, _p(ids), len(ids))

    def _fetch(self, nq, total):
        assert total >= 0, total
        off = np.empty(nq + 1, np.int64)
        ids = np.empty(max(total, 1), np.int32)
        lib().dpfo_get_candidates(self.h, _p(off), _p(ids))
        return off, ids[:total]

    def query_candidates_dense(self, Q, qids=None, steps=0, probe_mode=PROBE_DENSE, nthreads=0):
        Q = _f64(Q)
        qids = None if qids is None else _i32(qids)
        tot = lib().dpfo_query_candidates_dense(self.h, _p(Q), Q.shape[0], _p(qids), steps, probe_mode, nthreads)
        return self._fetch(Q.shape[0], tot)

    def query_candidates_csr(self, indptr, indices, values, qids=None, steps=0, nthreads=0):
        indptr, indices, values = _i64(indptr), _i32(indices), _f64(values)
        qids = None if qids is None else _i32(qids)
        nq = len(indptr) - 1
        tot = lib().dpfo_query_candidates_csr(self.h, _p(indptr), _p(indices), _p(values), nq, _p(qids), steps,
                                              nthreads)
        return self._fetch(nq, tot)

    def query_candidates_by_id(self, qids, steps=0, nthreads=0):
        qids = _i32(qids)
        tot = lib().dpfo_query_candidates_by_id(self.h, _p(qids), len(qids), steps, nthreads)
        return self._fetch(len(qids), tot)

    def rerank_dense(self, Q, offsets, cand, topk, metric=METRIC_DOT, nthreads=0):
        Q, offsets, cand = _f64(Q), _i64(offsets), _i32(cand)
        nq = Q.shape[0]
        ids = np.empty((nq, topk), np.int32)
        sc = np.empty((nq, topk), np.float64)
        rc = lib().dpfo_rerank_dense(self.h, _p(Q), nq, _p(offsets), _p(cand), topk, metric, _p(ids), _p(sc), nthreads)
        assert rc == 0, rc
        return ids, sc

    def query_topk_dense(self, Q, qids=None, steps=0, topk=10, metric=METRIC_DOT, probe_mode=PROBE_DENSE, nthreads=0):
        Q = _f64(Q)
        qids = None if qids is None else _i32(qids)
        nq = Q.shape[0]
        ids = np.empty((nq, topk), np.int32)
        sc = np.empty((nq, topk), np.float64)
        rc = lib().dpfo_query_topk_dense(self.h, _p(Q), nq, _p(qids), steps, probe_mode, topk, metric, _p(ids), _p(sc),
                                         nthreads)
        assert rc == 0, rc
        return ids, sc

    def dump_buckets(self, table):
        nb = lib().dpfo_dump_buckets(self.h, table, None, None, None)
        desc = np.empty((max(nb, 1), 3), np.int32)
        off = np.empty(nb + 1, np.int64)
        ids = np.empty(max(self.size(), 1), np.int32)
        lib().dpfo_dump_buckets(self.h, table, _p(desc), _p(off), _p(ids))
        return desc[:nb], off, ids[:off[nb]]

    def num_dir_nodes(self, table):
        return lib().dpfo_num_dir_nodes(self.h, table)

    def stats(self):
        s = np.zeros(8, np.int64)
        occ = np.zeros(1 << self.pb, np.float64)
        lib().dpfo_stats(self.h, _p(s), _p(occ))
        return dict(singleton_splits=int(s[0]), nlz_gt28=int(s[1]), splits=int(s[2]), MAXL=int(s[3]), nb=int(s[4]),
                    SEG=int(s[5]), n=int(s[6]), occupancy=occ)
