// rerank_units.cuh — declarations shared by the bucket-major re-rank kernels (rerank_bm.cu, rerank_u8.cu): the unit
// record, the FP64 tensor-pipe instruction wrapper and the geometry constants.
#pragma once
#include "query_common.cuh"

namespace dpf {

constexpr int BM_KC = 128;             // largest supported d

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}


// ---------------------------------------------------------------------------------------------------------
// units: runs of sorted pairs that share a bucket, cut into pieces of <= SS_UQ queries
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
    uint32_t pos0;         // position of the unit's first pair in the sorted pair list
    uint32_t m;            // queries in the unit (1..SS_UQ)
    int32_t q[SS_UQ];      // query index of pair j
    uint32_t seg[SS_UQ];   // start of pair j's score segment (dense output: k_score_stream)
    double tau[SS_UQ];     // that query's score threshold (filtered output: k_score_u8*, see Filter)
    int32_t ids0[SS_WIN];  // ids of the bucket's first 32 rows (the other windows are copied from ids_sorted)
};
static_assert(sizeof(UnitRec) % 16 == 0, "bulk copies move multiples of 16 bytes");

// Threshold filter.  A batch produces ~50k (query, candidate) scores per query of which k survive.  Writing them all
// and reading them back for the selection costs as many bytes as the byte rows themselves, so the scoring kernels
// only keep a score that can still be among the query's best k: before scoring, k_threshold scores the rows of the
// first buckets each query probes (its own bucket in the first tables) and takes the k-th best of them, minus a
// rounding allowance, as tau[q] — k distinct rows score >= tau, hence no row below tau can be in the result.  The
// kernels append (score, id) of the rows with score >= tau to the query's survivor list (a few hundred entries);
// k_select_survivors finishes.  The result is exactly the top k of all candidates.
struct Filter {
    uint32_t* cnt;         // per query of the chunk: survivors appended so far
    const uint32_t* base;  // per query: start of its survivor list (capacity = all its bucket entries)
    double* s_score;
    int32_t* s_id;
};
__device__ __forceinline__ void keep_survivor(const Filter& f, int q, double score, int id) {
    const uint32_t at = f.base[q] + atomicAdd(f.cnt + q, 1u);
    f.s_score[at] = score;
    f.s_id[at] = id;
}

constexpr size_t SS_WARP_BYTES = (size_t)SS_STAGES * SS_ROWS * SS_PITCH * sizeof(double) + 2 * sizeof(UnitRec) + 2 * SS_WIN_COPY * 4 + 32;
constexpr size_t SS_SMEM = SS_WARPS * ((SS_WARP_BYTES + 15) / 16 * 16);


// rerank_u8.cu: scores of every unit from the uint8 compact store (register gather, DMMA or IMMA)
bool score_u8_usable(const dpf_index* h);
int u8_query_pitch();                                                      // row pitch of dpf_index::Q8
void prepare_queries_u8(dpf_index* h, const double* Qd, int64_t nq);       // queries -> uint8 copy if they are bytes
void launch_score_u8(dpf_index* h, const double* Qd, const void* units, const uint32_t* nunits_p, bool angular,
                     const Filter& flt, unsigned long long* bm_stat);

}  // namespace dpf
