/*
 * dpf_oracle.h — C API of the CPU parity oracle for the Dynamic Partition Forest hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a from-scratch CPU restatement of the reference algorithm
 * (MacLLL/SimilaritySearchByRDF) used as the checker in tests/, in __graft_entry__.smoke() and as the
 * `cpu_baseline` / `--impl reference` leg of bench.py.  Nothing in the product path
 * (similaritysearchbyrdf_b200/, include/dpf.h, libdpf_b200.so) may include, link or call it.
 *
 * Parity pinning: the reference has no JVM-free build (Scala 2.10/sbt; no java/scalac in this image), so
 * oracle/_ref cannot be produced.  The oracle is pinned against every known-answer test the reference's own
 * test tree holds for this path (see tests/test_oracle_kat.py); end-to-end key / bucket / candidate outputs
 * on real data are NOT pinned by the reference itself (all its datasets are absent and its hash functions
 * are unseeded) — for those the oracle's literal sequential restatement is the definition.
 *
 * All file:line citations are relative to /root/reference/.
 */
#ifndef DPF_ORACLE_H
#define DPF_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct dpfo dpfo;

typedef struct {
    int32_t d;               /* vectorDim (dense) or feature-space size D (sparse) */
    int32_t L;               /* tableNum * permutationNum  (DensevectorRDFInit.scala:107)      */
    int32_t k;               /* mclab.lshTable.chainLength                                    */
    int32_t P;               /* number of distinct hash functions supplied in A               */
    int32_t pb;              /* mclab.lsh.partitionBits                                       */
    int32_t bucket_bits;     /* mclab.lshTable.bucketBits  (RandomDrawTreeMap.java:435-438)    */
    int32_t dir_node_size;   /* mclab.lshTable.dirNodeSize (RandomDrawTreeMap.java:446-465)    */
    int32_t bucket_overflow; /* mclab.lshTable.bufferOverflow = BUCKET_OVERFLOW               */
    int32_t family_kind;     /* 0 = angle (AngleHashFamily.scala), 1 = pStable                 */
    int32_t key_transform;   /* 0 original, 1 sampling, 2 continueBitsCount, 3 angleNewMethod  */
    int32_t self_exclude_small_ids; /* quirk Q3: RandomDrawTreeMap.java:982 (ids -128..127)    */
    int32_t rank, world;     /* emulate one shard of the content-based partition scheme: only the
                                sub-indexes p with p % world == rank are built and searched (0/1 = all) */
} dpfo_cfg;

enum { DPFO_METRIC_DOT = 0, DPFO_METRIC_ANGULAR = 1, DPFO_METRIC_L2 = 2 };
enum { DPFO_PROBE_NONE = 0, DPFO_PROBE_DENSE = 1 };

dpfo*  dpfo_create(const dpfo_cfg* cfg);
void   dpfo_destroy(dpfo* o);
/* A: P x d row-major; chain_idx: L x k (row of A); b,w: per-function pStable offset/width (may be NULL for angle) */
int    dpfo_set_family(dpfo* o, const double* A, const int32_t* chain_idx, const double* b, const int32_t* w);
/* Ap: L x pb x 32 row-major — per-table private partitioner functions (Partitioner.scala:27-64) */
int    dpfo_set_partitioners(dpfo* o, const double* Ap);
int    dpfo_set_partitioners_pstable(dpfo* o, const double* Ap /* L x pb x 32 */, const double* b /* L x pb */,
                                     const int32_t* w /* L x pb */);
int    dpfo_set_owned(dpfo* o, const uint8_t* owned /* 2^pb flags or NULL */);
int    dpfo_set_owned_cells(dpfo* o, const uint8_t* owned /* L x 2^pb flags or NULL */);

/* keys_out / pids_out: L x n (table-major).  nthreads<=0 => hardware_concurrency */
int    dpfo_hash_dense(dpfo* o, const double* X, int64_t n, int32_t* keys_out, int32_t* pids_out, int nthreads);
int    dpfo_hash_csr(dpfo* o, const int64_t* indptr, const int32_t* indices, const double* values, int64_t n,
                     int32_t* keys_out, int32_t* pids_out, int nthreads);

/* append n vectors; ids are the running counter (DensevectorRDFInit.scala:174-184).  The oracle keeps its own
 * copy of the data.  Insertion order per table = ascending id (newFastFit, DensevectorRDFInit.scala:127-151);
 * threads own table slices [i*L/T,(i+1)*L/T) (DensevectorRDFInit.scala:183-188). */
int    dpfo_fit_dense(dpfo* o, const double* X, int64_t n, int nthreads);
int    dpfo_fit_csr(dpfo* o, const int64_t* indptr, const int32_t* indices, const double* values, int64_t n, int nthreads);
int64_t dpfo_size(dpfo* o);
/* RandomDrawTreeMap.remove for every table (RandomDrawTreeMap.java:1817-1932); returns the (table, id) entries removed */
int64_t dpfo_remove(dpfo* o, const int32_t* ids, int64_t m);

/* candidate sets (sorted unique ids per query).  Result is held inside the oracle; returns total ids or <0. */
int64_t dpfo_query_candidates_dense(dpfo* o, const double* Q, int64_t nq, const int32_t* qids, int steps,
                                    int probe_mode, int nthreads);
int64_t dpfo_query_candidates_csr(dpfo* o, const int64_t* indptr, const int32_t* indices, const double* values,
                                  int64_t nq, const int32_t* qids, int steps, int nthreads);
/* id-based, probe-less query of already-indexed vectors (RandomDrawTreeMap.java:630-675) */
int64_t dpfo_query_candidates_by_id(dpfo* o, const int32_t* qids, int64_t nq, int steps, int nthreads);
int    dpfo_get_candidates(dpfo* o, int64_t* offsets_out /*nq+1*/, int32_t* ids_out);

/* re-rank given candidate sets: dense data only.  ids_out/score_out: nq x topk, padded with -1 / NaN. */
int    dpfo_rerank_dense(dpfo* o, const double* Q, int64_t nq, const int64_t* offsets, const int32_t* cand,
                         int topk, int metric, int32_t* ids_out, double* score_out, int nthreads);
int    dpfo_query_topk_dense(dpfo* o, const double* Q, int64_t nq, const int32_t* qids, int steps, int probe_mode,
                             int topk, int metric, int32_t* ids_out, double* score_out, int nthreads);

/* canonical dump of one table's forest: leaf buckets in (root, path) order, ids ascending inside a bucket.
 * desc: nbuckets x 3 int32 = (root = pid*SEG+seg, level of the bucket, path = slots MAXL..level packed nb bits each).
 * Call with NULL outputs to get the bucket count. */
int64_t dpfo_dump_buckets(dpfo* o, int table, int32_t* desc_out, int64_t* off_out, int32_t* ids_out);
int64_t dpfo_num_dir_nodes(dpfo* o, int table);

/* stats[0] singleton-split events (quirk Q1), [1] queries with nlz(h)>28 (quirk Q4), [2] splits, [3] MAXL,
 * [4] nb, [5] SEG, then 2^pb sub-index occupancy averaged over tables is in occ_out (may be NULL). */
int    dpfo_stats(dpfo* o, int64_t* stats_out /*8*/, double* occ_out /*2^pb*/);

/* ---- known-answer hooks (each restates one reference function) ---- */
double  dpfo_dot_dense(const double* a, const double* x, int d);                                   /* SimilarityCalculator.scala:29-49 */
double  dpfo_dot_sparse(const int32_t* ia, const double* va, int na, const int32_t* ib, const double* vb, int nb); /* :9-27 */
int32_t dpfo_angle_key_from_dots(const double* dots, int k);                                       /* AngleHashFamily.scala:184-195 */
int32_t dpfo_pstable_key_from_dots(const double* dots, const double* b, const int32_t* w, int k);  /* PStableHashFamily.scala:122-143 */
int32_t dpfo_sampling_key(int32_t key);                                                            /* Sampling.scala:6-39 */
void    dpfo_sampling_index(int32_t* sigma32);
int32_t dpfo_continue_bits_count(int32_t key);                                                     /* significantBits.scala:11-67, Array(6,4,2,1) */
int32_t dpfo_angle_new_method(int32_t key);                                                        /* significantBits.scala:100-127 */
int32_t dpfo_partition_id(int32_t h, const double* Ap_t /*pb x 32*/, int pb, int key_transform);  /* Partitioner.scala:40-64 */
int32_t dpfo_partition_id_pstable(int32_t h, const double* Ap_t, int pb, int key_transform, const double* b /* pb */,
                                  const int32_t* w /* pb */);   /* the same when the partitioner's family is pStable */
int32_t dpfo_default_hasher(int32_t key);                                                          /* Hasher.scala:18-37 */
/* bitmap-compressed directory arithmetic (RandomDrawTreeMap.java:1186-1267) — the layout the GPU replaces */
int32_t dpfo_dir_offset_from_slot(const int32_t* bitmap, int bitmap_words, int slot);
void    dpfo_tree_params(int bucket_bits, int dir_node_size, int chain_length, int32_t* out /*SEG,nb,mask,MAXL*/);

#ifdef __cplusplus
}
#endif
#endif
