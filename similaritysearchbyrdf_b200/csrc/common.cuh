// common.cuh — internal declarations shared by the translation units of libdpf_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/dpf.h"

namespace dpf {

// ---------------------------------------------------------------------------------------------------------
// error plumbing: device-side failures become DPF_ERR_* codes, never exceptions across the ABI
// ---------------------------------------------------------------------------------------------------------
struct Error {
    int code;
    std::string msg;
};

#define DPF_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            throw ::dpf::Error{e__ == cudaErrorMemoryAllocation ? DPF_ERR_NOMEM : DPF_ERR_CUDA,          \
                               std::string(#expr) + ": " + cudaGetErrorString(e__)};                     \
    } while (0)

#define DPF_REQUIRE(cond, code, text)                        \
    do {                                                     \
        if (!(cond)) throw ::dpf::Error{(code), (text)};     \
    } while (0)

// Device memory is stream-ordered: every buffer of a handle is carved out of the device's default memory pool with
// cudaMallocAsync / cudaFreeAsync on the stream the handle works on (dpf_create raises the pool's release threshold,
// so freed blocks stay cached in the pool).  A re-build or a new handle therefore re-uses warm blocks without a
// device-wide synchronisation, which is what cudaMalloc/cudaFree cost inside the build.  `tl_stream` is the stream
// of the API call running on this thread (set by every entry point before it touches a buffer).
extern thread_local cudaStream_t tl_stream;

// owning device buffer (grow-only); all allocations of a handle go through these
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFreeAsync(p, tl_stream);
        p = nullptr;
        cap = 0;
    }
    // ensure capacity >= n elements; contents are NOT preserved
    void reserve(size_t n) {
        if (n <= cap) return;
        release();
        DPF_CUDA(cudaMallocAsync((void**)&p, n * sizeof(T), tl_stream));
        cap = n;
    }
    // ensure capacity >= n elements, keeping the first `keep` elements
    void grow_keep(size_t n, size_t keep, cudaStream_t st) {
        if (n <= cap) return;
        size_t ncap = n + n / 4;
        T* q = nullptr;
        DPF_CUDA(cudaMallocAsync((void**)&q, ncap * sizeof(T), st));
        if (keep && p) DPF_CUDA(cudaMemcpyAsync(q, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st));
        if (p) cudaFreeAsync(p, st);
        p = q;
        cap = ncap;
    }
};

// tree geometry (RandomDrawTreeMap.java:435-465)
struct TreeParams {
    int SEG;          // 2^(32-bucket_bits)
    int seg_bits;     // 32-bucket_bits
    int nb;           // log2(dirNodeSize)
    int W;            // dirNodeSize = 2^nb
    int MAXL;         // (chainLength - seg_bits)/nb - 1
    int bucket_bits;
    int pb;           // partition bits
    int R;            // roots per table = 2^pb * SEG
    int T;            // BUCKET_OVERFLOW
};

// flat forest in HBM: node n has W children; child_cnt: 0 empty, >0 leaf bucket of that many ids starting at
// ids_sorted[table_base[t] + child_ptr], <0 directory whose node index is child_ptr.
struct ForestView {
    const int32_t* child_ptr;
    const int32_t* child_cnt;
    const int32_t* ids_sorted;
    const int64_t* table_base;  // L+1
    int32_t num_nodes;
    // leaf table (bucket-major re-rank): every leaf bucket has a dense number; child_leaf[slot] = that number for a
    // leaf slot, leaf_pos / leaf_len = its range in ids_sorted (global position, < 2^32 when the table is built)
    const uint32_t* child_leaf;
    const uint32_t* leaf_pos;
    const int32_t* leaf_len;
    int32_t num_leaves;
};

enum : int { kMaxTables = 256, kMaxChain = 32, kMaxPb = 8 };

// int32 slots of dpf_index::counters (device).  [0, 16) belong to whichever of hash / build is running; the rest are
// per-query-batch counters, cleared at the start of a batch by one memset of [CTR_BATCH_FIRST, CTR_BATCH_END).
enum : int {
    CTR_FIX_COUNT = 0,        // hash: length of the near-zero list of the current chunk
    CTR_BAD_ID = 16,          // by-id queries: some id is not in the index   (also k_expand's query cursor)
    CTR_NEXT_UNIT = 17,       // row-major re-rank: unit cursor
    CTR_STAT_NLZ = 20,        // u64: (query, table) pairs with nlz(h) > 28
    CTR_STAT_UNIQUE = 24,     // u64: unique candidates of the batch (candidate-set path)
    CTR_BM_STAT = 26,         // 3 x u64: units scored, rows staged, survivors selected from
    CTR_STORE_FLAGS = 40,     // fit: representability scan
    CTR_Q8_BAD = 41,          // != 0: some query value of the batch is not a byte
    CTR_POOL = 42,            // survivor pool cursor (records)
    CTR_NPAIRS = 44,          // bucket-major: (bucket, query) pairs / units of the current chunk
    CTR_NUNITS = 45,
    CTR_SCAN_TILE = 46,       // tile cursor of the leaf-count scan
    CTR_POOL_OVERFLOW = 47,   // != 0: the survivor pool was too small for some warp (its queries are answered directly)
    CTR_NDIRTY = 48,          // queries of the current chunk flagged for k_topk_direct
    CTR_DIRECT = 49,          // queries answered by k_topk_direct in this batch
    CTR_NPAIRS_TC = 54,       // the same two totals from the scan at the tcgen05 kernel's unit width
    CTR_NUNITS_TC = 55,
    CTR_SCAN_TILE_TC = 56,
    CTR_BM_PAIRS_TOTAL = 50,  // u64: pairs over all chunks of the batch
    CTR_ENTRIES = 52,         // u64: bucket entries visited by the batch (candidates with duplicates)
    CTR_BATCH_FIRST = 16,
    CTR_BATCH_END = 64,
    CTR_FIX_TOTAL = 64,       // u64, cumulative: near-zero projections recomputed in reference order
    CTR_COUNT = 128
};

// the (table, sub-index) cells this handle owns on a multi-GPU box: L rows of 256 bits in device memory (partition ids
// are < 2^pb <= 256).  The reference keeps one store per (table, partition) (RandomDrawTreeMap.java:1430-1459), so a
// cell is the unit a placement can move; the plain scheme gives rank g the cells with p mod G == g in every table, the
// balanced one deals cells out by their size.
constexpr int kOwnWords = 8;
struct OwnMask {
    const uint32_t* cells;
    __device__ bool has(int t, int pid) const { return (cells[t * kOwnWords + (pid >> 5)] >> (pid & 31)) & 1u; }
};

// process-wide count of kernel launches issued by this library (reported through dpf_stats)
extern unsigned long long g_launches;
#define DPF_LAUNCHED() (++::dpf::g_launches)

}  // namespace dpf

// ---------------------------------------------------------------------------------------------------------
// the handle
// ---------------------------------------------------------------------------------------------------------
struct dpf_index {
    dpf_config cfg{};
    dpf::TreeParams tp{};
    int P = 0;
    int PW = 0;  // sign words per vector = ceil(P/32)
    cudaStream_t stream = nullptr;      // stream all work of this handle runs on
    cudaStream_t own_stream = nullptr;  // the handle's own stream while a caller stream is installed
    cudaStream_t aux_stream = nullptr;  // second stream for work that can overlap the main one (thresholds ‖ pair sort)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::mutex mu;
    std::string last_error;
    bool family_set = false, part_set = false, dense = true, fitted = false;
    int num_sms = 148;
    dpf::OwnMask own{};                  // cells owned by this rank (all of them when world <= 1): view of own_dev
    std::vector<uint32_t> own_cells;     // host copy, L x kOwnWords
    dpf::DevBuf<uint32_t> own_dev;
    bool owns(int t, int p) const { return (own_cells[(size_t)t * dpf::kOwnWords + (p >> 5)] >> (p & 31)) & 1u; }
    bool balance_partition = false;      // dpf_set_balanced_partition: ownership by occupancy instead of p % world
    bool own_fixed = false;              // the balanced assignment is made once, at the first fit

    // hash functions
    dpf::DevBuf<double> A;        // P x d row-major
    dpf::DevBuf<double> At;       // d x Ppad feature-major copy for the CSR kernel (built on demand)
    dpf::DevBuf<double> Anorm;    // P  (||a_p||_2)
    dpf::DevBuf<int32_t> chain;   // L x k
    dpf::DevBuf<double> fb;       // P (pStable b)
    dpf::DevBuf<int32_t> fw;      // P (pStable w)
    dpf::DevBuf<double> Ap;       // L x pb x 32
    dpf::DevBuf<double> Apb;      // L x pb: b of the partitioner chains (pStable family; dpf_set_partitioners_pstable)
    dpf::DevBuf<int32_t> Apw;     // L x pb: w
    bool part_pstable = false;
    std::vector<double> hA;       // host copy (for At construction)
    int At_ld = 0;

    // data store
    int64_t n = 0;
    dpf::DevBuf<double> X;        // owned dense store (n x d) unless borrowed
    const double* Xdev = nullptr; // points at X.p or at the caller's buffer (dpf_fit_dense_dev)
    bool X_borrowed = false;
    // compact store: lossless narrow copy of the dense store for the re-rank kernels (store.cu)
    int store_mode = DPF_STORE_AUTO;
    int Xc_kind = DPF_STORE_KIND_F64;   // element type of Xc; F64 = no copy, the kernels read Xdev
    int64_t Xc_row_bytes = 0;           // row pitch of Xc in bytes (multiple of 16, zero padded)
    dpf::DevBuf<unsigned char> Xc;
    dpf::DevBuf<unsigned char> Q8;      // uint8 copy of the current query batch, when every query value is a byte
    dpf::DevBuf<double> qnorm8;         //   and the queries' norms
    dpf::DevBuf<int32_t> qsq8;          //   and squared norms (integers)
    bool Q8_valid = false;
    dpf::DevBuf<int64_t> sp_ptr;  // CSR store
    dpf::DevBuf<int32_t> sp_idx;
    dpf::DevBuf<double> sp_val;
    int64_t sp_nnz = 0;

    // per-vector keys / partition ids, table-major with leading dimension key_ld
    dpf::DevBuf<int32_t> keys;
    dpf::DevBuf<uint8_t> pids;
    int64_t key_ld = 0;

    // forest
    dpf::DevBuf<int32_t> child_ptr, child_cnt, ids_sorted;
    dpf::DevBuf<int32_t> node_table;            // table of every directory node
    dpf::DevBuf<uint32_t> child_leaf, leaf_pos; // leaf table (see ForestView)
    dpf::DevBuf<int32_t> leaf_len;
    dpf::DevBuf<uint32_t> leaf_cnt;             // per leaf: pairs of the current query chunk (all zero between batches)
    dpf::DevBuf<uint32_t> leaf_off, leaf_unit_off;   //           first pair / first unit
    dpf::DevBuf<uint32_t> leaf_unit_off_tc;          //           first unit at the tcgen05 kernel's unit width
    dpf::DevBuf<char> bm_descs;                      // TcRec records (the tcgen05 kernel's units)
    dpf::DevBuf<int32_t> bm_taui;                    // per query: integer score threshold
    int64_t max_leaf_len = 0, total_leaf_tiles = 0;  // largest leaf bucket; sum over leaves of ceil(len / 128)
    int64_t arena_used = 0;                     // entries of ids_sorted in use: the tables, then buckets moved by put
    dpf::DevBuf<uint8_t> removed;               // per id: removed (a rebuild skips it); empty until the first remove
    int64_t n_removed = 0;
    int32_t num_leaves = 0;
    bool leaf_table = false;                    // false when the forest has >= 2^32 entries (row-major re-rank only)
    dpf::DevBuf<int64_t> table_base;
    std::vector<int64_t> h_table_base;
    std::vector<double> occupancy;   // 2^pb: ids per sub-index averaged over tables (last build)
    int32_t num_nodes = 0, node_cap = 0;

    // scratch
    dpf::DevBuf<uint32_t> signs;      // n x PW
    dpf::DevBuf<int32_t> pq;          // n x P quantised pStable values
    dpf::DevBuf<int2> fix_list;
    dpf::DevBuf<int32_t> counters;    // small device counters
    dpf::DevBuf<uint32_t> sk0, sk1, sv0, sv1;  // sort ping-pong
    dpf::DevBuf<uint32_t> hist;
    dpf::DevBuf<int32_t> work0, work1;         // split worklists
    dpf::DevBuf<uint32_t> bitmap;              // candidate de-dup bitmaps
    dpf::DevBuf<int64_t> q_off;                // nq+1
    dpf::DevBuf<int32_t> q_cnt;                // nq
    dpf::DevBuf<int32_t> cand;                 // candidate ids
    dpf::DevBuf<unsigned long long> sk64a, sk64b;
    dpf::DevBuf<double> qbuf;                  // device copy of queries
    dpf::DevBuf<int32_t> qidbuf, qkeys;
    dpf::DevBuf<uint8_t> qpids;
    dpf::DevBuf<int32_t> out_ids;
    dpf::DevBuf<int32_t> ucnt, part_id;        // re-rank work units / partial top-k lists
    dpf::DevBuf<uint32_t> scan_scratch, pair_cnt, pair_base, pair_seg, pair_len;   // bucket-major re-rank
    dpf::DevBuf<int32_t> pair_q;
    dpf::DevBuf<uint32_t> probe_cache;         // leaf numbers of the distinct buckets per (query, table) (k_probe_leaves)
    dpf::DevBuf<unsigned long long> pair_key, pair_key_alt;
    dpf::DevBuf<double> scores;                // bucket-major re-rank: survivor scores / ids (Filter, rerank_units.cuh)
    dpf::DevBuf<int32_t> surv_id;
    dpf::DevBuf<char> surv_pool;               // SurvRec blocks written by the scoring warps
    dpf::DevBuf<double> bm_tau;                // per query: score threshold
    dpf::DevBuf<uint32_t> bm_scnt, bm_sbase;   //            survivors so far, start of its list
    dpf::DevBuf<uint32_t> bm_big;              // queries whose survivor list is long
    dpf::DevBuf<double> direct_keys;           // k_topk_direct: partial lists per (dirty query, table group)
    dpf::DevBuf<int32_t> direct_ids;
    dpf::DevBuf<double> bm_tl_keys;            // threshold sample lists: nq x NT x k
    dpf::DevBuf<int> bm_tl_ids, bm_tl_cnt;
    dpf::DevBuf<uint32_t> bm_flag, bm_run_start, bm_ucnt, bm_counts;   // runs / units of the sorted pairs
    dpf::DevBuf<char> bm_units;
    unsigned long long* bm_sorted = nullptr;   // pair keys sorted by bucket (points into pair_key_alt / sk64a)
    int64_t bm_npairs = 0;
    dpf::DevBuf<int64_t> unit_off;
    dpf::DevBuf<double> part_key;
    dpf::DevBuf<double> out_scores;
    dpf::DevBuf<char> stage;                   // generic staging

    // test / profiling hooks (dpf_set_debug_option); the product path never reads the environment
    int64_t dbg[DPF_DBG_COUNT] = {0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

    // multi-GPU plane (comm.cu): NCCL communicator of this rank, staging of the key all-gather, the top-k gather buffer
    void* comm = nullptr;
    dpf::DevBuf<char> comm_stage, comm_topk;

    // stats / profiling
    int64_t stats[DPF_STAT_COUNT] = {0};
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;          // event pairs; one pair per timed stage instance of the last call
    std::vector<int> ev_stage;                 // stage id of pair i
    size_t ev_used = 0;                        // pairs used by the current call
    float stage_ms[DPF_T_COUNT] = {0};
};

namespace dpf {

// RAII stage timer: records a CUDA event pair on the handle's stream around a stage when profiling is on; a
// stage that runs several times in one call (chunks) gets one pair per instance and the times are summed
struct StageTimer {
    dpf_index* h;
    size_t slot = 0;
    bool on;
    cudaStream_t st;
    StageTimer(dpf_index* h_, int id, cudaStream_t stream = nullptr) : h(h_), on(h_->profiling), st(stream ? stream : h_->stream) {
        if (!on) return;
        slot = h->ev_used++;
        if (h->ev_pool.size() < 2 * (slot + 1)) {
            cudaEvent_t a = nullptr, b = nullptr;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            h->ev_pool.push_back(a);
            h->ev_pool.push_back(b);
            h->ev_stage.push_back(id);
        }
        h->ev_stage[slot] = id;
        cudaEventRecord(h->ev_pool[2 * slot], st);
    }
    ~StageTimer() {
        if (on) cudaEventRecord(h->ev_pool[2 * slot + 1], st);
    }
};

// ---- hash.cu ------------------------------------------------------------------------------------------------
// keys/pids for n dense vectors already on the device; keys_out/pids_out table-major with leading dim ld
void hash_dense_device(dpf_index* h, const double* Xd, int64_t n, int32_t* keys_out, uint8_t* pids_out, int64_t ld);
void hash_dense_device_exact(dpf_index* h, const double* Xd, int64_t n, int32_t* keys_out, uint8_t* pids_out, int64_t ld);
void hash_csr_device(dpf_index* h, const int64_t* ptr, const int32_t* idx, const double* val, int64_t n,
                     int32_t* keys_out, uint8_t* pids_out, int64_t ld);
void prepare_family(dpf_index* h);

// ---- store.cu -----------------------------------------------------------------------------------------------
// (re)builds the compact store from Xdev[0 .. n); leaves Xc_kind = F64 when no narrower type is lossless
void build_compact_store(dpf_index* h);
bool append_compact_store(dpf_index* h, int64_t n_old, int64_t m);

// ---- sort.cu ------------------------------------------------------------------------------------------------
// stable LSD radix sort of the bit range [lo_bit, hi_bit) — result ends in (*keys_io, *vals_io) which may be
// swapped with the alternates.
void radix_sort_pairs_u32(dpf_index* h, uint32_t** keys, uint32_t** keys_alt, uint32_t** vals, uint32_t** vals_alt,
                          int64_t n, int lo_bit, int hi_bit);
void radix_sort_keys_u64(dpf_index* h, unsigned long long** keys, unsigned long long** keys_alt, int64_t n, int lo_bit,
                         int hi_bit);
void exclusive_scan_i64(dpf_index* h, const int32_t* in, int64_t* out, int64_t n);  // out has n+1 entries
void exclusive_scan_u32(dpf_index* h, uint32_t* data, int64_t n);                   // in place, multi-CTA

// ---- forest.cu ----------------------------------------------------------------------------------------------
void build_forest(dpf_index* h);
void rebuild_leaf_table(dpf_index* h);
// ---- incremental.cu -------------------------------------------------------------------------------------------
bool forest_insert_incremental(dpf_index* h, int64_t n_old, int64_t m);   // false: no room, rebuild
int64_t forest_remove(dpf_index* h, const int32_t* ids_host, int64_t m);  // (table, id) entries removed
void update_occupancy(dpf_index* h);                                      // sub-index occupancy after a put / remove
ForestView forest_view(const dpf_index* h);

// ---- query.cu -----------------------------------------------------------------------------------------------
struct QueryKeys {
    const int32_t* keys;   // L x ld
    int64_t ld;
    int64_t nq;
    const int32_t* qids;   // device, may be null
};
void probe_count_all(dpf_index* h, const QueryKeys& qk, int steps, int probe_mode, std::vector<int64_t>& off_host);
int64_t next_chunk_end(const std::vector<int64_t>& off, int64_t q0, int64_t budget);
void expand_range(dpf_index* h, const QueryKeys& qk, int steps, int probe_mode, int64_t q0, int64_t q1, int64_t base,
                  int64_t ub_chunk);
int64_t finalize_candidates_sorted(dpf_index* h, int64_t q0, int64_t q1, int64_t base, int64_t* off_dev);
void rerank_topk(dpf_index* h, const double* Qd, int64_t q0, int64_t q1, int64_t base, const int64_t* off,
                 const int32_t* cnt, const int32_t* cand, int64_t max_cnt, int64_t total_ub, int topk, int metric,
                 int32_t* ids_out, double* score_out);
// bucket-major re-rank of queries [q0, q1) (rerank_bm.cu): top-k straight from the probe result, no candidate lists
bool bucket_major_supported(const dpf_index* h, int metric, int topk);
bool score_u8_usable(const dpf_index* h);                                  // rerank_u8.cu: byte store present and enabled
void prepare_queries_u8(dpf_index* h, const double* Qd, int64_t nq, bool need_host_flag);   // byte copy of the batch + device flag
void topk_bucket_major(dpf_index* h, const double* Qd, const QueryKeys& qk, int steps, int probe_mode, int64_t q0, int64_t q1,
                       int cap, int topk, int metric, int32_t* ids_out, double* score_out);
void tc_diag_read(unsigned long long* out8);   // rerank_tc.cu
int64_t bm_chunk_queries(const dpf_index* h, int steps, int probe_mode, int* cap_out);   // bm_group.cu
constexpr int64_t kMaxPool = 1LL << 29;                  // survivor records of a chunk (28 bytes each with their lists)
int64_t bm_pool_per_query(const dpf_index* h, int topk, int steps);   // rerank_bm.cu
void gather_query_keys(dpf_index* h, const int32_t* qids_dev, int64_t nq);
void counts_from_offsets(dpf_index* h, const int64_t* off, int64_t nq, int32_t* cnt);
void merge_topk(dpf_index* h, const int32_t* gids, const double* gsc, int G, int64_t nq, int topk, int metric,
                int32_t* ids_out, double* score_out);
// list g of the gathered results starts at gids + g * stride_ids / gsc + g * stride_sc (elements)
void merge_topk_strided(dpf_index* h, const int32_t* gids, const double* gsc, int64_t stride_ids, int64_t stride_sc, int G, int64_t nq,
                        int topk, int metric, int32_t* ids_out, double* score_out);

// ---- comm.cu ------------------------------------------------------------------------------------------------
void comm_unique_id(uint8_t* out);
void comm_init(dpf_index* h, const uint8_t* id_bytes);
void comm_destroy(dpf_index* h);
// keys / sub-index ids of n new vectors (ids n_old ..): this rank hashes its slice, one all-gather, every rank has all
void hash_dense_sharded(dpf_index* h, const double* Xnew, int64_t n, int64_t n_old);
// this rank's segment of the top-k gather buffer; after the local top k was written there: all-gather + merge
void comm_topk_buffers(dpf_index* h, int64_t nq, int topk, int32_t** ids_seg, double** sc_seg);
void comm_gather_merge_topk(dpf_index* h, int64_t nq, int topk, int metric, int32_t* ids_out, double* score_out);

}  // namespace dpf
