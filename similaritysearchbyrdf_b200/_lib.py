"""ctypes loader for libdpf_b200.so (the C ABI of include/dpf.h).  There is no CPU fallback: if the library is
missing or cannot be loaded the import of any product entry point fails loudly."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libdpf_b200.so")

ABI_VERSION = 1
OK, ERR_INVALID, ERR_STATE, ERR_CUDA, ERR_NOMEM, ERR_CAPACITY = range(6)
FAMILY_ANGLE, FAMILY_PSTABLE = 0, 1
KEY_ORIGINAL, KEY_SAMPLING, KEY_CONTINUE_BITS, KEY_ANGLE_NEW = 0, 1, 2, 3
METRIC_DOT, METRIC_ANGULAR, METRIC_L2 = 0, 1, 2
PROBE_NONE, PROBE_DENSE = 0, 1
STORE_KIND_F64, STORE_KIND_F32, STORE_KIND_U8 = 0, 1, 2
STORE_AUTO, STORE_F64_ONLY, STORE_NARROWEST = 0, 1, 2
STAT_COUNT, T_COUNT = 16, 16
# dpf_set_debug_option keys (include/dpf.h): test / profiling hooks, defaults are the product path
(DBG_RERANK, DBG_BM_KERNEL, DBG_U8_IMMA, DBG_U8I_KERNEL, DBG_TAU_TABLES, DBG_TAU_KERNEL, DBG_HASH_EXACT, DBG_CAND_BUDGET,
 DBG_TRACE, DBG_STORE, DBG_POOL_RECORDS, DBG_APPEND, DBG_WIDE, DBG_TAU_FORK) = range(14)
DBG_DEFAULTS = {DBG_U8_IMMA: 1}
STAT_NAMES = ["size", "near_zero_fixups", "singleton_splits", "splits", "dir_nodes", "nlz_gt28", "last_candidates",
              "last_cand_with_dups", "kernel_launches", "bm_pairs", "bm_runs",
              "bm_rows_staged", "store_kind", "store_row_bytes", "bm_survivors", "bm_direct"]
STAGE_NAMES = ["hash", "fixup", "pack", "sort", "split", "probe_count", "expand", "rerank", "cand_sort", "select", "narrow",
               "comm", "threshold"]
COMM_ID_BYTES = 128

# every symbol include/dpf.h declares
EXPORTS = [
    "dpf_create", "dpf_destroy", "dpf_last_error", "dpf_strerror", "dpf_sync", "dpf_set_stream", "dpf_set_family",
    "dpf_set_partitioners", "dpf_set_partitioners_pstable", "dpf_hash_dense", "dpf_hash_csr", "dpf_fit_dense", "dpf_fit_csr", "dpf_fit_dense_dev",
    "dpf_size", "dpf_query_candidates_dense", "dpf_query_candidates_csr", "dpf_query_candidates_by_id",
    "dpf_query_topk_dense", "dpf_query_topk_dense_dev", "dpf_rerank_dense", "dpf_merge_topk_dev", "dpf_dump_buckets",
    "dpf_stats", "dpf_set_profiling", "dpf_stage_times_ms", "dpf_set_store_mode", "dpf_save", "dpf_load", "dpf_set_balanced_partition", "dpf_owned_subindexes", "dpf_parse_dense_file", "dpf_parse_sparse_file",
    "dpf_set_debug_option", "dpf_debug_leaf_pairs", "dpf_debug_tc_diag", "dpf_remove", "dpf_comm_unique_id", "dpf_comm_init", "dpf_comm_destroy",
    "dpf_fit_dense_sharded", "dpf_fit_dense_sharded_dev", "dpf_query_topk_dense_all", "dpf_query_topk_dense_all_dev",
]


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "device", "d", "L", "k", "pb", "bucket_bits", "dir_node_size", "bucket_overflow", "family_kind",
        "key_transform", "self_exclude_small_ids", "rank", "world")]


class DpfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libdpf_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Loads the shared library and declares the prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
            "this package has no CPU fallback")
    L = C.CDLL(SO_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.dpf_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.dpf_destroy.argtypes = [vp]
    L.dpf_last_error.restype = C.c_char_p
    L.dpf_last_error.argtypes = [vp]
    L.dpf_strerror.restype = C.c_char_p
    L.dpf_strerror.argtypes = [C.c_int]
    L.dpf_sync.argtypes = [vp]
    L.dpf_set_stream.argtypes = [vp, vp]
    L.dpf_set_family.argtypes = [vp, vp, i32, vp, vp, vp]
    L.dpf_set_partitioners.argtypes = [vp, vp]
    L.dpf_set_partitioners_pstable.argtypes = [vp, vp, vp, vp]
    L.dpf_hash_dense.argtypes = [vp, vp, i64, vp, vp]
    L.dpf_hash_csr.argtypes = [vp, vp, vp, vp, i64, vp, vp]
    L.dpf_fit_dense.argtypes = [vp, vp, i64]
    L.dpf_fit_csr.argtypes = [vp, vp, vp, vp, i64]
    L.dpf_fit_dense_dev.argtypes = [vp, vp, i64]
    L.dpf_size.restype = i64
    L.dpf_size.argtypes = [vp]
    L.dpf_query_candidates_dense.argtypes = [vp, vp, i64, vp, i32, i32, vp, vp, i64, vp]
    L.dpf_query_candidates_csr.argtypes = [vp, vp, vp, vp, i64, vp, i32, vp, vp, i64, vp]
    L.dpf_query_candidates_by_id.argtypes = [vp, vp, i64, i32, vp, vp, i64, vp]
    L.dpf_query_topk_dense.argtypes = [vp, vp, i64, vp, i32, i32, i32, i32, vp, vp]
    L.dpf_query_topk_dense_dev.argtypes = [vp, vp, i64, vp, i32, i32, i32, i32, vp, vp]
    L.dpf_rerank_dense.argtypes = [vp, vp, i64, vp, vp, i32, i32, vp, vp]
    L.dpf_merge_topk_dev.argtypes = [vp, vp, vp, i32, i64, i32, i32, vp, vp]
    L.dpf_dump_buckets.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.dpf_stats.argtypes = [vp, vp, vp]
    L.dpf_set_profiling.argtypes = [vp, i32]
    L.dpf_stage_times_ms.argtypes = [vp, vp]
    L.dpf_set_store_mode.argtypes = [vp, i32]
    L.dpf_set_debug_option.argtypes = [vp, i32, i64]
    L.dpf_debug_leaf_pairs.argtypes = [vp, vp, vp, vp]
    L.dpf_debug_tc_diag.argtypes = [vp, vp]
    L.dpf_remove.argtypes = [vp, vp, i64, vp]
    L.dpf_comm_unique_id.argtypes = [vp]
    L.dpf_comm_init.argtypes = [vp, vp]
    L.dpf_comm_destroy.argtypes = [vp]
    L.dpf_fit_dense_sharded.argtypes = [vp, vp, i64]
    L.dpf_fit_dense_sharded_dev.argtypes = [vp, vp, i64]
    L.dpf_query_topk_dense_all.argtypes = [vp, vp, i64, vp, i32, i32, i32, i32, vp, vp]
    L.dpf_query_topk_dense_all_dev.argtypes = [vp, vp, i64, vp, i32, i32, i32, i32, vp, vp]
    L.dpf_set_balanced_partition.argtypes = [vp, i32]
    L.dpf_owned_subindexes.argtypes = [vp, vp]
    L.dpf_save.argtypes = [vp, C.c_char_p]
    L.dpf_parse_dense_file.argtypes = [C.c_char_p, i32, vp, i64, vp]
    L.dpf_parse_sparse_file.argtypes = [C.c_char_p, vp, vp, vp, i64, i64, vp, vp, vp]
    L.dpf_load.argtypes = [C.c_char_p, i32, C.POINTER(vp)]
    for name in EXPORTS:
        getattr(L, name)
    _lib = L
    return L
