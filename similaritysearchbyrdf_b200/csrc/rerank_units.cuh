// rerank_units.cuh — declarations shared by the bucket-major re-rank kernels (rerank_bm.cu, rerank_u8.cu): the unit
// record, the FP64 tensor-pipe instruction wrapper and the geometry constants.
#pragma once
#include <cstddef>

#include "query_common.cuh"

namespace dpf {

constexpr int BM_KC = 128;             // largest d of the staged kernels (k_score_stream, byte kernels); wider rows: rerank_wide.cu

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}


// ---------------------------------------------------------------------------------------------------------
// units: the (bucket, query) pairs of a batch grouped by bucket (counting sort over the leaf numbers, bm_group.cu),
// every bucket's list cut into pieces of <= SS_UQ queries
// ---------------------------------------------------------------------------------------------------------
constexpr int SS_UQ = 16;              // queries per unit = 2 DMMA n-blocks held in registers
constexpr int SS_WARPS = 8;            // warps per CTA, one CTA per SM
constexpr int SS_STAGES = 3;           // ring slots per warp
constexpr int SS_ROWS = 8;             // rows per slot = DMMA M extent (bucket rows) / N extent (query rows)
constexpr int SS_PITCH = BM_KC + 8;    // doubles; (pitch * 8 B) mod 128 == 64: conflict-free LDS.128 per quarter warp
constexpr int SS_WIN = 32;             // row ids per id window
constexpr int SS_WIN_COPY = SS_WIN + 4;   // ids copied per window: the copy starts at a 16-byte boundary <= the window

// Everything the scoring warp needs to know about a unit, in one 16-byte-aligned record it can pull into shared
// memory with a single bulk copy (no dependent global loads in the scoring loop).
struct __align__(16) UnitRec {
    uint32_t bstart;       // bucket start in ids_sorted
    uint32_t len;          // bucket length (rows)
    uint32_t pos0;         // position of the unit's first pair in the grouped pair list
    uint32_t m;            // queries in the unit (1..SS_UQ)
    int32_t q[SS_UQ];      // query (index inside the chunk) of pair j; slots >= m repeat the last query
    int32_t ids0[SS_WIN];  // ids of the bucket's first 32 rows (the other windows are copied from ids_sorted)
};
static_assert(sizeof(UnitRec) % 16 == 0, "bulk copies move multiples of 16 bytes");

// The tcgen05 kernel (rerank_tc.cu) takes other units: <= TC_TQ queries of one bucket x <= 128 of its rows (the N and M
// extents of its MMA), one self-contained record each, streamed into shared memory ahead of use.
constexpr int TC_TQ = 32;
struct __align__(16) TcRec {
    uint32_t bstart;       // bucket start in ids_sorted
    uint32_t nrows;        // rows of this unit (1..128)
    uint32_t m;            // queries in the unit (1..TC_TQ)
    uint32_t row0;         // first row of the unit inside the bucket (multiple of 128)
    int32_t q[TC_TQ];      // query (index inside the chunk) of column j; slots >= m repeat the last query
    int32_t ids[128];      // row ids; slots >= nrows repeat the last row
};
static_assert(sizeof(TcRec) == 16 + 4 * TC_TQ + 512, "record layout");

// Threshold filter.  A batch produces ~50k (query, candidate) scores per query of which k survive.  Writing them all
// and reading them back for the selection costs as many bytes as the byte rows themselves, so the scoring kernels
// only keep a score that can still be among the query's best k: before scoring, k_threshold scores the rows of the
// first buckets each query probes (its own bucket in the first tables) and takes the k-th best of them, minus a
// rounding allowance, as tau[q] — k distinct rows score >= tau, hence no row below tau can be in the result.  The
// kernels emit (query, row, score) of the rows with score >= tau (a few hundred per query, SurvivorSink);
// k_scatter_survivors sorts them into per-query lists and k_select_survivors finishes.  The result is exactly the top k of all candidates.
// Nothing here is sized by a number the host would have to read back: the pool has a fixed capacity, the per-query lists
// are laid out from the counts the scoring kernel leaves behind, and a query whose survivors did not fit (or whose
// samples held fewer than k rows: tau = +inf, nothing survives) is flagged `dirty` and answered by k_topk_direct, which
// walks the query's own buckets and needs no scratch — so a batch runs without a single host synchronisation.
struct __align__(16) SurvRec {
    int32_t q;             // query, or -1 for an unused slot
    uint32_t pos;          // position of the row in ids_sorted
    double score;
};
// the queries the filter cannot serve, answered by k_topk_direct: a flag per query, and a list so that the exhaustive
// kernel can spread every such query over many CTAs (and costs nothing when the list is empty)
struct DirtySet {
    uint32_t* flag;        // per query
    int32_t* list;         // dirty queries in the order they were found
    int* count;
    __device__ __forceinline__ void mark(int q) const {
        if (atomicExch(flag + q, 1u) == 0u) list[atomicAdd(count, 1)] = q;
    }
};

struct Filter {
    const double* tau;     // per query: score threshold (k_threshold*); +inf for a dirty query
    uint32_t* cnt;         // per query: survivors stored in the pool (k_count_survivors)
    uint32_t* base;        // per query: start of its survivor list = exclusive scan of cnt (k_survivor_offsets)
    uint32_t* fill;        // per query: list entries written so far (k_scatter_survivors)
    DirtySet dirty;        // queries answered by k_topk_direct
    double* s_score;
    int32_t* s_id;
    SurvRec* pool;         // survivors in the order the scoring warps found them, blocks of SURV_BLOCK per warp
    uint32_t* pool_cursor;
    uint32_t pool_cap;     // records, a multiple of SURV_BLOCK
    int* overflow;         // set when a warp could not get a block
};
constexpr int SURV_BLOCK = 256;

// Per-warp survivor output.  An append straight into the query's list needs the old value of a global atomic before
// it can store: the scoring warp sits out a full memory round trip per survivor (measured: 22 % of k_score_u8s).
// Instead the warp writes (query, row position, score) records into blocks it reserves from one pool — one atomic per
// 256 survivors — and k_scatter_survivors, a streaming kernel with parallelism to hide that latency, moves the
// records into the per-query lists.  push() must be called by all 32 lanes.
struct SurvivorSink {
    uint32_t wpos = 0;
    int wleft = 0;
    bool full = false;                                     // the pool ran out: everything further marks its query dirty
    __device__ __forceinline__ void push(const Filter& f, bool keep, int q, uint32_t pos, double score, int lane) {
        const uint32_t mask = __ballot_sync(0xffffffffu, keep);
        if (!mask) return;
        const int n = __popc(mask);
        if (n > wleft && !full) {                          // open a new block; the rest of the old one stays unused
            if (lane < wleft) f.pool[wpos + lane].q = -1;
            uint32_t b = 0;
            if (lane == 0) b = atomicAdd(f.pool_cursor, (uint32_t)SURV_BLOCK);
            wpos = __shfl_sync(0xffffffffu, b, 0);
            wleft = SURV_BLOCK;
            if (wpos > f.pool_cap - (uint32_t)SURV_BLOCK) {
                full = true;
                wleft = 0;
                if (lane == 0) *f.overflow = 1;
            }
        }
        if (full) {
            if (keep) f.dirty.mark(q);
            return;
        }
        if (keep) {
            SurvRec r;
            r.q = q; r.pos = pos; r.score = score;
            f.pool[wpos + __popc(mask & ((1u << lane) - 1u))] = r;
        }
        wpos += n;
        wleft -= n;
    }
    __device__ __forceinline__ void flush(const Filter& f, int lane) {
        if (full) return;
        for (int i = lane; i < wleft; i += 32) f.pool[wpos + i].q = -1;
    }
};

constexpr size_t SS_WARP_BYTES = (size_t)SS_STAGES * SS_ROWS * SS_PITCH * sizeof(double) + 2 * sizeof(UnitRec) + 2 * SS_WIN_COPY * 4 + 32;
constexpr size_t SS_SMEM = SS_WARPS * ((SS_WARP_BYTES + 15) / 16 * 16);


// rerank_u8.cu: scores of every unit from the uint8 compact store (register gather, DMMA or IMMA)
// one chunk of a query batch as the bucket-major kernels see it: every per-query array starts at the chunk's first query
struct ChunkView {
    const double* Q;             // nqc x d
    const unsigned char* Q8;     // byte copy of the queries (valid when *q8_bad == 0)
    const double* qnorm8;
    const int32_t* qsq8;
    const int32_t* qids;         // may be null
    int64_t nqc;
    const int* q8_bad;           // device flag: some query of the BATCH is not a byte vector
    // probe result: distinct leaves per (query, table)
    const uint32_t* pair_cnt;    // nqc x L
    const uint32_t* cache;       // nqc x L x cap leaf numbers
    int cap;
};

// rerank_u8.cu: scores of every unit from the uint8 compact store (register gather, DMMA or IMMA)
void launch_threshold_u8i(dpf_index* h, cudaStream_t st, int metric, const ChunkView& cv, int NT, int topk, size_t list_smem);
int u8_query_pitch();                                                      // row pitch of dpf_index::Q8
void launch_score_u8(dpf_index* h, const ChunkView& cv, const void* units, const uint32_t* nunits_p, int metric,
                     const Filter& flt, unsigned long long* bm_stat, bool int_kernel);
// bm_group.cu: probe -> pairs grouped by leaf -> unit records, all sized on the host without reading anything back
void probe_leaves(dpf_index* h, const QueryKeys& qk, int steps, int probe_mode, int64_t q0, int64_t nqc, int cap, uint32_t* q_entries);
void group_pairs(dpf_index* h, int64_t nqc, int cap, bool use_tc);
void emit_units(dpf_index* h, bool only_if_fp64_queries);
void emit_tc_recs(dpf_index* h, int64_t cap, const DirtySet& dirty);                                           // the tcgen05 kernel's units
// rerank_wide.cu: d > BM_KC (FP64 rows streamed from global memory, no staging)
bool wide_supported(const dpf_index* h, int metric);
void launch_threshold_wide(dpf_index* h, cudaStream_t st, int metric, const ChunkView& cv, int NT, int topk, size_t list_smem);
void launch_score_wide(dpf_index* h, const ChunkView& cv, const void* units, const uint32_t* nunits_p, int metric, const Filter& flt,
                       unsigned long long* bm_stat);
// rerank_tc.cu
bool score_u8t_usable(const dpf_index* h, int metric);
void launch_score_u8t(dpf_index* h, const ChunkView& cv, const TcRec* recs, const uint32_t* nunits_p, int64_t cap, const int32_t* taui,
                      const Filter& flt, unsigned long long* bm_stat);
void survivor_lists(dpf_index* h, const Filter& flt, int64_t nqc);         // offsets + scatter
int64_t bm_chunk_queries(const dpf_index* h, int steps, int probe_mode, int* cap_out);
// query.cu: exhaustive per-query top-k over the query's own buckets for the queries flagged dirty
void topk_direct(dpf_index* h, const ChunkView& cv, const DirtySet& dirty, int topk, int metric, int32_t* ids_out, double* score_out);

}  // namespace dpf
