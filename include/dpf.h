/*
 * dpf.h — C ABI of libdpf_b200: the B200-native (sm_100a) replacement for the data-parallel hot path of
 * Dynamic Partition Forest (reference: MacLLL/SimilaritySearchByRDF; all file:line below are relative to
 * the reference tree).
 *
 * The reference has no FFI of its own: its plug points are Scala/Java types.  Each entry
 * point below names the reference interface it replaces; INTEGRATION.md shows the JNI/Scala stub that binds it.
 *
 * Conventions
 *   - every function returns an int status (DPF_OK == 0); nothing aborts the process, no C++ exception
 *     crosses the boundary; dpf_last_error(h) gives the message of the last failure on that handle;
 *   - the caller owns every buffer it passes; the library owns all device memory behind a handle;
 *   - plain functions take HOST pointers (pinned memory is faster but not required) and include the
 *     host<->device copies; *_dev functions take DEVICE pointers on the handle's GPU and are asynchronous on
 *     the handle's stream until dpf_sync;
 *   - a handle is thread-compatible (calls are serialised internally); separate handles are independent;
 *   - there is no CPU fallback: with no usable CUDA device dpf_create fails with DPF_ERR_CUDA.
 *   - arrays are C-order, native endianness; ids are int32 running counters assigned in fit order
 *     (DensevectorRDFInit.scala:174-184: the id in the input file is ignored).
 */
#ifndef DPF_H
#define DPF_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define DPF_ABI_VERSION 1

typedef struct dpf_index* dpf_handle;

enum {
    DPF_OK = 0,
    DPF_ERR_INVALID = 1,   /* bad argument / configuration                                   */
    DPF_ERR_STATE = 2,     /* call order (family not set, not fitted, dense/sparse mixed)     */
    DPF_ERR_CUDA = 3,      /* CUDA runtime failure or no device                               */
    DPF_ERR_NOMEM = 4,     /* device or host allocation failed                                */
    DPF_ERR_CAPACITY = 5   /* caller's output buffer too small (needed size is reported)      */
};

enum { DPF_FAMILY_ANGLE = 0, DPF_FAMILY_PSTABLE = 1 };                 /* mclab.lsh.name (LSH.scala:29-53) */
enum { DPF_KEY_ORIGINAL = 0, DPF_KEY_SAMPLING = 1, DPF_KEY_CONTINUE_BITS = 2, DPF_KEY_ANGLE_NEW = 3 };
                                                                       /* mclab.lsh.typeOfIndex (LSH.scala:152-161) */
enum { DPF_METRIC_DOT = 0, DPF_METRIC_ANGULAR = 1, DPF_METRIC_L2 = 2 };
enum { DPF_PROBE_NONE = 0, DPF_PROBE_DENSE = 1 };
/* element type of the compact vector store the re-rank kernels read (see dpf_set_store_mode) */
enum { DPF_STORE_KIND_F64 = 0, DPF_STORE_KIND_F32 = 1, DPF_STORE_KIND_U8 = 2 };
enum { DPF_STORE_AUTO = 0, DPF_STORE_F64_ONLY = 1, DPF_STORE_NARROWEST = 2 };
    /* DPF_PROBE_DENSE: getSimilarWithStepWiseFaster(key, DenseVector, steps) multi-probe (RandomDrawTreeMap.java:742-797)
       DPF_PROBE_NONE : sparse overload / id-based getSimilarWithStepWise (RandomDrawTreeMap.java:686-732, 630-675) */

typedef struct {
    int32_t abi_version;      /* DPF_ABI_VERSION                                                              */
    int32_t device;           /* CUDA device ordinal                                                           */
    int32_t d;                /* mclab.lsh.vectorDim (dense) or feature-space size D (sparse)                  */
    int32_t L;                /* tableNum * permutationNum (DensevectorRDFInit.scala:107)                      */
    int32_t k;                /* mclab.lshTable.chainLength (1..32)                                            */
    int32_t pb;               /* mclab.lsh.partitionBits (0..8): 2^pb sub-indexes per table                    */
    int32_t bucket_bits;      /* mclab.lshTable.bucketBits   (RandomDrawTreeMap.java:435-438)                   */
    int32_t dir_node_size;    /* mclab.lshTable.dirNodeSize  (RandomDrawTreeMap.java:446-465), power of two     */
    int32_t bucket_overflow;  /* mclab.lshTable.bufferOverflow = BUCKET_OVERFLOW (RandomDrawTreeMap.java:1719) */
    int32_t family_kind;      /* DPF_FAMILY_*                                                                  */
    int32_t key_transform;    /* DPF_KEY_*                                                                     */
    int32_t self_exclude_small_ids; /* reproduce `ln.key != key` on boxed Integers (RandomDrawTreeMap.java:982) */
    int32_t rank;             /* content-based partition scheme on G GPUs: this handle owns the sub-indexes    */
    int32_t world;            /*   { p : p % world == rank } of every table (Partitioner.scala:27-64); 0,1 = all */
} dpf_config;

/* ---- lifetime ------------------------------------------------------------------------------------------- */
/* replaces DensevectorRDFInit.initializeRDFHashMap (DensevectorRDFInit.scala:50-118): table construction */
int dpf_create(const dpf_config* cfg, dpf_handle* out);
int dpf_destroy(dpf_handle h);                       /* clearAndClose (DensevectorRDFInit.scala:453-458) */
const char* dpf_last_error(dpf_handle h);
const char* dpf_strerror(int code);
int dpf_sync(dpf_handle h);
/* run the handle's work on the caller's CUDA stream (cudaStream_t; e.g. torch.cuda.current_stream().cuda_stream)
 * so that the caller's CUDA events bracket it; NULL restores the handle's own stream */
int dpf_set_stream(dpf_handle h, void* cuda_stream);

/* ---- hash functions: inputs, generated/loaded on the host side (AngleHashFamily.scala:121-177) ---------- */
/* A: P x d row-major distinct functions; chain_idx: L x k -> row of A (tableIndexGenerators, LSH.scala:27);
 * b[P], w[P]: pStable offsets/widths (PStableHashFamily.scala:185-191), NULL for the angle family. */
int dpf_set_family(dpf_handle h, const double* A, int32_t P, const int32_t* chain_idx, const double* b,
                   const int32_t* w);
/* Ap: L x pb x 32 — each table's private LocalitySensitivePartitioner functions (DensevectorRDFInit.scala:63-77) */
int dpf_set_partitioners(dpf_handle h, const double* Ap);
/* The same for a DPF_FAMILY_PSTABLE index.  The reference builds the partitioner's LSH from the main family's
 * configuration with vectorDim = 32 and chainLength = pb (DensevectorRDFInit.scala:63-70), so there the sub-index is
 * PStableHashChain.compute(bits of the key).hashCode >>> (32 - pb) (Partitioner.scala:40-64, PStableHashFamily.scala:
 * 155-177): b[L x pb], w[L x pb] are the offsets / widths of the chain functions.  dpf_set_partitioners on a pStable
 * index with pb > 0 returns DPF_ERR_INVALID. */
int dpf_set_partitioners_pstable(dpf_handle h, const double* Ap, const double* b, const int32_t* w);

/* ---- LSH.calculateIndex + LocalitySensitivePartitioner.getPartition, batched (parity hook) --------------
 * replaces LSH.calculateIndex (LSH.scala:93-166) / LocalitySensitiveHasher.hash (Hasher.scala:44-54) /
 * getPartition (Partitioner.scala:40-64).  keys_out, pids_out: L x n int32, table-major; pids_out may be NULL. */
int dpf_hash_dense(dpf_handle h, const double* X, int64_t n, int32_t* keys_out, int32_t* pids_out);
int dpf_hash_csr(dpf_handle h, const int64_t* indptr, const int32_t* indices, const double* values, int64_t n,
                 int32_t* keys_out, int32_t* pids_out);

/* ---- fit: DensevectorRDFInit.newMultiThreadFit / newFastFit (DensevectorRDFInit.scala:127-206) and the
 * sparse twin (SparsevectorRDFInit.scala:124-203), after the text has been parsed.  Appends n vectors with ids
 * size()..size()+n-1 and (re)builds the forest so that it equals sequential ascending-id RandomDrawTreeMap.put
 * (RandomDrawTreeMap.java:1558-1584, 1662-1790). */
int dpf_fit_dense(dpf_handle h, const double* X, int64_t n);
int dpf_fit_csr(dpf_handle h, const int64_t* indptr, const int32_t* indices, const double* values, int64_t n);
/* X_dev stays owned by the caller and must outlive the handle (no copy is made: 100M x 96 FP64 is 76.8 GB) */
int dpf_fit_dense_dev(dpf_handle h, const double* X_dev, int64_t n);
int64_t dpf_size(dpf_handle h);    /* ids handed out so far (removed ids are not reused) */
/* RandomDrawTreeMap.remove (RandomDrawTreeMap.java:1817-1932) for a batch of ids, in every table: an id leaves its bucket,
 * an emptied bucket frees its slot, a directory node left without children is deleted from its parent (the roots stay) —
 * with the configured tree geometry, not the reference's hard-coded 4 levels x 7 bits.
 * removed_entries_out: (table, id) entries that were found and removed.  Ids that are not in the index are ignored.
 * A small dpf_fit_dense afterwards is PUT id by id into the forest as it is (exactly the reference's sequence); an append
 * so large that the forest is rebuilt re-inserts the surviving ids in ascending order, which the reference's history-
 * dependent tree need not equal bucket for bucket (results are the same sets whenever no bucket is at the split limit). */
int dpf_remove(dpf_handle h, const int32_t* ids, int64_t m, int64_t* removed_entries_out);
/* The reference keeps every vector as double[] (vectorIdToVector, DensevectorRDFInit.scala:35-36).  A dense fit also
 * checks whether EVERY stored value survives a round trip through a narrower type and, if so, keeps a copy of the rows
 * in that type for the re-rank kernels, which widen it back to the identical doubles in registers (or, when the queries
 * of a batch are bytes too, multiply exact integers): the products and sums are the same, only 1/8 (1/2) of the bytes
 * cross HBM.  Lossy narrowing is never done.  DPF_STORE_AUTO (default): uint8 when lossless (SIFT-like descriptors);
 * DPF_STORE_NARROWEST: uint8, else float when lossless (fvecs data); DPF_STORE_F64_ONLY: no check, FP64 rows only.
 * Call before fit.  DPF_STAT_STORE_KIND reports the kind in use. */
int dpf_set_store_mode(dpf_handle h, int32_t mode);

/* ---- query: candidate sets = DensevectorRDFInit.NewMultiThreadQueryBatch (DensevectorRDFInit.scala:335-360),
 * SparsevectorRDFInit.NewMultiThreadQueryBatch (SparsevectorRDFInit.scala:324-348) --------------------------
 * qids may be NULL (no self-exclusion).  Output is CSR: offsets_out[nq+1], ids_out sorted unique per query.
 * If total > cap the call returns DPF_ERR_CAPACITY with *total_out set; offsets_out is still valid. */
int dpf_query_candidates_dense(dpf_handle h, const double* Q, int64_t nq, const int32_t* qids, int32_t steps,
                               int32_t probe_mode, int64_t* offsets_out, int32_t* ids_out, int64_t cap,
                               int64_t* total_out);
int dpf_query_candidates_csr(dpf_handle h, const int64_t* indptr, const int32_t* indices, const double* values,
                             int64_t nq, const int32_t* qids, int32_t steps, int64_t* offsets_out,
                             int32_t* ids_out, int64_t cap, int64_t* total_out);
int dpf_query_candidates_by_id(dpf_handle h, const int32_t* qids, int64_t nq, int32_t steps,
                               int64_t* offsets_out, int32_t* ids_out, int64_t cap, int64_t* total_out);

/* ---- query + re-rank + top-k = DensevectorRDFInit.topKAndPrecisionScore's gather/dgemv/argsort
 * (DensevectorRDFInit.scala:472-507).  metric DOT is the reference's (descending dot product); ANGULAR
 * (descending cosine) and L2 (ascending squared distance) are the north-star additions.  Ties: ascending id.
 * ids_out: nq x topk (-1 padded), score_out: nq x topk (NaN padded). */
int dpf_query_topk_dense(dpf_handle h, const double* Q, int64_t nq, const int32_t* qids, int32_t steps,
                         int32_t probe_mode, int32_t topk, int32_t metric, int32_t* ids_out, double* score_out);
int dpf_query_topk_dense_dev(dpf_handle h, const double* Q_dev, int64_t nq, const int32_t* qids_dev,
                             int32_t steps, int32_t probe_mode, int32_t topk, int32_t metric,
                             int32_t* ids_out_dev, double* score_out_dev);
/* re-rank caller-supplied candidate sets (parity hook: "ids exact given the same candidate set") */
int dpf_rerank_dense(dpf_handle h, const double* Q, int64_t nq, const int64_t* offsets, const int32_t* cand,
                     int32_t topk, int32_t metric, int32_t* ids_out, double* score_out);

/* ---- text loaders (host code, no handle): the file formats newMultiThreadFit reads — dense `[id,[v1,v2,...]]`
 * (Vectors.parseDense, Vector.scala:215-219; read loop DensevectorRDFInit.scala:172-181) and sparse
 * `[id, size, [i1,...], [v1,...]]` (Vectors.fromPythonString, Vector.scala:194-208; SparsevectorRDFInit.scala:164-176).
 * Rows come back in file order (the id in the file is ignored, as in the reference), empty lines are skipped, sparse
 * indices ascending.  Call once with NULL outputs for the sizes, then with buffers; DPF_ERR_CAPACITY if they are too
 * small, DPF_ERR_INVALID for a line that does not parse (dense: not exactly d values). */
int dpf_parse_dense_file(const char* path, int32_t d, double* X_out, int64_t cap_rows, int64_t* n_out);
int dpf_parse_sparse_file(const char* path, int64_t* indptr_out, int32_t* indices_out, double* values_out,
                          int64_t cap_rows, int64_t cap_nnz, int64_t* n_out, int64_t* nnz_out, int32_t* dim_out);

/* ---- persist / reload: the replacement for the reference's RAM-threshold spill of tree pages to disk
 * (RandomDrawTreeMap.java:2713-2773, StoreSegment.java:489-545).  dpf_save writes configuration, hash functions, the
 * vector store, every vector's keys / sub-index ids and the flat forest; dpf_load restores them on `device` (no re-hashing,
 * no rebuild), bit-identical to the saved index — also after incremental puts and removes. */
/* A file is the shard of the rank that wrote it: configuration, hash functions, ownership of sub-indexes (also a
 * balanced one), store mode, vectors, keys, removed ids and the forest arrays as they stand; every rank of a multi-GPU
 * index saves and loads its own file. */
int dpf_save(dpf_handle h, const char* path);
int dpf_load(const char* path, int32_t device, dpf_handle* out);

/* ---- multi-GPU: which sub-indexes a rank owns.  Default: sub-index p of every table lives on GPU p mod world.  The
 * content-based partition is skewed (sub-indexes hold 64k..224k of 1M ids in configs[1]), so with enable != 0 the first
 * dense fit deals the L x 2^pb (table, sub-index) cells — each one of the reference's per-partition stores — to the GPUs
 * by occupancy instead (largest first, to the least loaded GPU); every rank derives the same assignment from the
 * replicated vectors.  Results are unaffected.  Call before fit. */
int dpf_set_balanced_partition(dpf_handle h, int32_t enable);
/* owned_out[t * 2^pb + p] = 1 for the cells (sub-index p of table t) this handle owns, L x 2^pb flags (after the first
 * fit when the assignment is balanced; p mod world == rank in every table otherwise) */
int dpf_owned_subindexes(dpf_handle h, uint8_t* owned_out);

/* ---- multi-GPU data plane inside the library: one handle (rank) per GPU, NCCL over NVLink between them (the library
 * binds libnccl.so.2 at run time).  Replaces the in-process partition scheme of the reference (Partitioner.scala:27-64,
 * RandomDrawTreeMap.java:613-621, 1430-1459) for the GPUs of one box, whether the ranks are threads of one process or
 * one process each.  Rank 0 obtains an id, the caller hands it to every rank (any channel), every rank calls
 * dpf_comm_init (collective).  After that:
 *   dpf_fit_dense_sharded[_dev]   collective fit: every rank passes the same vectors, hashes 1/world of them, and one
 *                                 all-gather of the keys (5 L bytes per vector) lets each build the sub-forests it owns;
 *   dpf_query_topk_dense_all[_dev] collective query: same queries on every rank, per-rank top k -> one all-gather of
 *                                 12-byte entries -> merge; every rank gets the global result.  No host synchronisation
 *                                 in the _dev form. */
#define DPF_COMM_ID_BYTES 128
int dpf_comm_unique_id(uint8_t* id_out /* DPF_COMM_ID_BYTES */);
int dpf_comm_init(dpf_handle h, const uint8_t* id /* DPF_COMM_ID_BYTES */);
int dpf_comm_destroy(dpf_handle h);
int dpf_fit_dense_sharded(dpf_handle h, const double* X, int64_t n);
int dpf_fit_dense_sharded_dev(dpf_handle h, const double* X_dev, int64_t n);
int dpf_query_topk_dense_all(dpf_handle h, const double* Q, int64_t nq, const int32_t* qids, int32_t steps,
                             int32_t probe_mode, int32_t topk, int32_t metric, int32_t* ids_out, double* score_out);
int dpf_query_topk_dense_all_dev(dpf_handle h, const double* Q_dev, int64_t nq, const int32_t* qids_dev, int32_t steps,
                                 int32_t probe_mode, int32_t topk, int32_t metric, int32_t* ids_out_dev,
                                 double* score_out_dev);

/* ---- multi-GPU: merge per-GPU top-k lists after the NCCL all-gather -----------------------
 * gathered_*_dev: G x nq x topk as produced by all-gathering dpf_query_topk_dense_dev outputs; duplicates of
 * one id (same vector reached through tables owned by different GPUs) are collapsed. */
int dpf_merge_topk_dev(dpf_handle h, const int32_t* gathered_ids_dev, const double* gathered_scores_dev,
                       int32_t G, int64_t nq, int32_t topk, int32_t metric, int32_t* ids_out_dev,
                       double* score_out_dev);

/* ---- introspection --------------------------------------------------------------------------------------- */
/* canonical dump of one table's forest (parity hook for bucket membership): leaf buckets in (root, path) order,
 * ids ascending inside a bucket.  desc: nbuckets x 3 = (root = pid*SEG+seg, level, path of slots MAXL..level).
 * Pass NULLs to get counts. */
int dpf_dump_buckets(dpf_handle h, int32_t table, int64_t* nbuckets_out, int64_t* nids_out, int32_t* desc_out,
                     int64_t* off_out, int32_t* ids_out);

enum {
    DPF_STAT_SIZE = 0,             /* vectors indexed                                                        */
    DPF_STAT_NEAR_ZERO_FIXUPS = 1, /* projections within the DMMA error bound of 0, recomputed in reference
                                      order (cumulative over hash calls)                                    */
    DPF_STAT_SINGLETON_SPLITS = 2, /* quirk Q1 events (RandomDrawTreeMap.java:1733-1734), last build          */
    DPF_STAT_SPLITS = 3,           /* directory nodes created by bucket overflow, last build                 */
    DPF_STAT_DIR_NODES = 4,        /* all directory nodes incl. roots                                        */
    DPF_STAT_NLZ_GT28 = 5,         /* (query,table) pairs with nlz(h) > 28 (quirk Q4), last query            */
    DPF_STAT_LAST_CANDIDATES = 6,  /* unique candidates of the last query batch                              */
    DPF_STAT_LAST_CAND_WITH_DUPS = 7, /* bucket entries visited by the last query batch                      */
    DPF_STAT_KERNEL_LAUNCHES = 8,  /* kernels launched by this library in this process (cumulative)          */
    DPF_STAT_BM_PAIRS = 9,         /* bucket-major re-rank, last batch: (bucket, query) pairs                */
    DPF_STAT_BM_RUNS = 10,         /*   units scored (<= 16 queries sharing a bucket)                         */
    DPF_STAT_BM_ROWS_STAGED = 11,  /*   bucket rows staged in shared memory (each one row of the store)      */
    DPF_STAT_STORE_KIND = 12,      /* DPF_STORE_KIND_* of the compact store after the last dense fit         */
    DPF_STAT_STORE_ROW_BYTES = 13, /* bytes per row the re-rank kernels fetch                                */
    DPF_STAT_BM_SURVIVORS = 14,    /* bucket-major re-rank, last batch: scores that passed the threshold filter */
    DPF_STAT_BM_DIRECT = 15,       /*   queries answered by the exhaustive per-query kernel (no threshold could be
                                        guaranteed from the samples, or their survivors did not fit the pool)         */
    DPF_STAT_COUNT = 16
};
int dpf_stats(dpf_handle h, int64_t* stats_out /* DPF_STAT_COUNT */, double* occupancy_out /* 2^pb or NULL */);
/* introspection of the last bucket-major query chunk: for every leaf bucket of the forest, the start of its list in
 * the grouped (bucket, query) pair array (nleaves + 1 offsets) and its length in rows.  Pass NULLs to get the count. */
/* watchdog record of the tcgen05 scoring kernel: all zero unless some role of some CTA waited ~1 s for a barrier
 * ([0] = first barrier tag << 32 | CTA, [1..5] = time-outs per barrier: tile full, accumulator empty, accumulator
 * full, tile empty, operand empty) */
int dpf_debug_tc_diag(dpf_handle h, uint64_t* out24 /* 8 watchdog words, then 16 cycle counters: [8 + tag] = cycles
                                                          warps spent waiting on that barrier, [23] = kernel cycles x CTAs */);
int dpf_debug_leaf_pairs(dpf_handle h, int64_t* nleaves_out, uint32_t* pair_off_out, int32_t* leaf_len_out);

/* per-stage device times of the last fit / query call, measured with CUDA events on the handle's stream */
enum {
    DPF_T_HASH = 0, DPF_T_FIXUP = 1, DPF_T_PACK = 2, DPF_T_SORT = 3, DPF_T_SPLIT = 4,
    DPF_T_PROBE_COUNT = 5, DPF_T_EXPAND = 6, DPF_T_RERANK = 7, DPF_T_CAND_SORT = 8, DPF_T_SELECT = 9, DPF_T_NARROW = 10,
    DPF_T_COMM = 11, DPF_T_THRESHOLD = 12,   /* THRESHOLD runs on the handle's second stream, beside EXPAND */
    DPF_T_COUNT = 16
};
int dpf_set_profiling(dpf_handle h, int32_t enable);

/* ---- test / profiling hooks.  The product path reads no environment variable: which kernel variant runs is a
 * property of the handle, changed only through this call (defaults = the product path).  Tests use it to compare the
 * independent kernels with each other and with the oracle. */
enum {
    DPF_DBG_RERANK = 0,       /* 0 bucket-major when supported (default), 1 always the row-major kernel              */
    DPF_DBG_BM_KERNEL = 1,    /* 0 default, 1 the TMA-ring FP64 kernel (k_score_stream) even on byte rows             */
    DPF_DBG_U8_IMMA = 2,      /* 1 default: integer tensor pipe on byte rows x byte queries; 0: FP64 tensor pipe      */
    DPF_DBG_U8I_KERNEL = 3,   /* byte rows x byte queries: 0 default (mma.sync, cp.async ring), 1 lean register gather,
                                 2 = 0, 3 tcgen05 / TMEM kernel (dot product)                                         */
    DPF_DBG_TAU_TABLES = 4,   /* threshold samples per query; 0 = default                                             */
    DPF_DBG_TAU_KERNEL = 5,   /* 0 default (tensor pipe), 1 CUDA-core DP4A / FMA form                                 */
    DPF_DBG_HASH_EXACT = 6,   /* 1: angle keys from the reference-order CUDA-core kernel instead of DMMA + fix-up     */
                              /* 2: keys packed by the bit-by-bit kernel instead of the table-driven one             */
    DPF_DBG_CAND_BUDGET = 7,  /* candidate ids per chunk of a query batch (row-major / candidate-set paths); 0 default */
    DPF_DBG_TRACE = 8,        /* 1: host wall-clock per phase of a fit, to stderr                                     */
    DPF_DBG_STORE = 9,        /* 0 default, 1 keep FP64 rows only, 2 force a float copy (skips the byte check)        */
    DPF_DBG_POOL_RECORDS = 10,/* capacity of the survivor pool in records; 0 = default (tests force the overflow path) */
    DPF_DBG_APPEND = 11,      /* appending fit: 0 put small batches incrementally, rebuild for large ones (default);
                                 1 always rebuild; 2 always put incrementally (rebuild only when out of head-room)     */
    DPF_DBG_WIDE = 12,        /* d > 128: 0 bucket-major k_score_wide for dot / angular (default), 1 the row-major kernel */
    DPF_DBG_TAU_FORK = 13,    /* where the threshold stream forks: 0 / 1 right after the probe (default), 2 after the pair fill            */
    DPF_DBG_COUNT = 16
};
int dpf_set_debug_option(dpf_handle h, int32_t option, int64_t value);
int dpf_stage_times_ms(dpf_handle h, float* ms_out /* DPF_T_COUNT */);

#ifdef __cplusplus
}
#endif
#endif
