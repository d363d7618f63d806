// rerank_bm.cu — K5b: bucket-major re-rank.
//
// Replaces topKAndPrecisionScore's gather + dgemv + argsort (src/main/scala/mclab/deploy/DensevectorRDFInit.scala:
// 472-507) for a whole query batch.  The row-major kernel in query.cu reads a candidate row once per (query,
// candidate): 8d bytes of HBM for 2d flops.  In a batch, many queries probe the same leaf buckets (a bucket is
// probed by every query whose key falls in it or one bit away), so the same rows are fetched again and again.
// Here the batch is regrouped by bucket:
//   k_pairs_from_cache / k_probe_pairs   one (bucket, query) pair per distinct bucket a query probes: the buckets the
//                    probe pass (K4, k_probe_count) found, or the same walk again when they were not cached
//   radix sort       pairs by bucket start
//   k_run_flags .. k_emit_units   runs of pairs that share a bucket, cut into units of <= 16 queries; one
//                    self-contained record per unit (bucket, query list, score segments, first row ids)
// and scored unit by unit, in one of two pipelines chosen by the store (store.cu):
//   FP64 (float) rows   k_score_stream: a warp per unit; the unit's queries and then the bucket's rows stream L2/HBM ->
//                    shared memory through a per-warp ring of TMA bulk copies (one copy per row, mbarrier completion, 2
//                    slots of 8 rows in flight per warp while a third is consumed); queries become DMMA B fragments in
//                    registers, rows are multiplied against them -> one score per (pair, row), dense;
//                    k_select_pairs: per query, top-k over its score segments
//   byte rows        k_threshold* (here / rerank_u8.cu): per query a score that k distinct candidates provably reach;
//                    k_score_u8s / k_score_u8d (rerank_u8.cu): integer / FP64 tensor pipe, only scores >= the threshold
//                    are emitted (SurvivorSink); k_scatter_survivors, k_select_survivors(_big): per-query lists, top-k
// Both de-duplicate ids reached through several tables (their scores are bit-identical: same row, same query, same k
// order).  Candidate *sets* are unchanged (same probe), so results equal the row-major path up to summation order
// (exactly, on integer data).  Why TMA rows for FP64: tools/gather_patterns.cu — a DMMA A-fragment gather straight from
// HBM touches 8 rows x 64 B per instruction and collapses to 2.9 TB/s at high occupancy; 1 KB bulk row copies hold
// 7.3 TB/s with 8 warps/SM.
#include <cstdlib>
#include <type_traits>

#include "rerank_units.cuh"

namespace dpf {


bool bucket_major_supported(const dpf_index* h, int metric, int topk) {
    if (h->dbg[DPF_DBG_RERANK] == 1) return false;             // test hook: force the row-major kernel
    if (!(h->leaf_table && h->dense && h->Xdev && (h->cfg.d % 2) == 0 && (reinterpret_cast<uintptr_t>(h->Xdev) & 15) == 0 &&
          topk <= RR_MAXK))                                    // d even: the queries are FP64 rows moved in 16-byte pieces
        return false;
    if (h->cfg.d > BM_KC) return wide_supported(h, metric);    // rerank_wide.cu
    if (metric == DPF_METRIC_DOT || metric == DPF_METRIC_ANGULAR) return true;
    // squared L2 = |q|^2 + |x|^2 - 2 q.x cancels in floating point; it is exact — and then identical to the reference's
    // sum of squared differences — only on the integer pipeline: byte rows and a batch of byte queries (the one case in
    // which the host has to read the batch's byte flag back before it can choose the path)
    return metric == DPF_METRIC_L2 && score_u8_usable(h) && h->Q8_valid;
}

// Threshold samples per query (one warp each = the query's own bucket in one table).  Byte rows: 6 on one GPU; a rank of G
// sees 1/G of every query's buckets, so its lists stay short with fewer samples — measured per rank at configs[1]
// (tools/nt_sweep.py --world G), 3 tables is the best or within 1% of it at G = 2, 4 and 8 (one table leaves 860 survivors per
// query on the fullest rank of 8 and costs 0.2 ms more than it saves).  FP64 / float rows cost 8 / 4 times the bytes of byte
// rows per sampled row: two samples there (measured on configs[1] with the byte copy off: 4.1 ms of sampling with six tables
// against 11.4 ms of scoring).  Wide rows with k > 32: the k-th best of two buckets' rows is hardly a bar: four.  Multi-step
// search on wide rows visits 3-4 times the entries: twice the samples keep the survivor lists in proportion.
static int bm_threshold_tables(const dpf_index* h, int topk, int steps) {
    const bool wide = h->cfg.d > BM_KC;
    const bool use_u8 = !wide && score_u8_usable(h);
    const int world = h->cfg.world > 1 ? h->cfg.world : 1;
    const int nt_default = (!use_u8 ? (wide && topk > 32 ? 4 : 2) : (world <= 1 ? 6 : 3)) * (wide && steps > 0 ? 2 : 1);
    return std::min(32, std::max(1, h->dbg[DPF_DBG_TAU_TABLES] > 0 ? (int)h->dbg[DPF_DBG_TAU_TABLES] : nt_default));
}

// Survivor records the pool holds per query of a chunk.  The survivors of a query are roughly k x (entries it visits) /
// (rows sampled for its threshold): ~60k entries at steps = 0 (3-4 times that with multi-step search), ~200 rows per sampled
// bucket — the pool takes four times that estimate, never less than 2048 (configs[1], 6 samples: 330 survivors per query
// measured; GIST shape, k = 100, 4 samples: 3.4k at steps = 0, 6-9k with 8 samples at steps >= 1; Deep shape, FP64 rows,
// 2 samples: the former fixed 2048 overflowed and sent 9967 of 10000 queries to the exhaustive kernel).  topk_device cuts
// the batch so that a chunk's share fits kMaxPool.
int64_t bm_pool_per_query(const dpf_index* h, int topk, int steps) {
    const int64_t est = 1200LL * topk * (steps > 0 ? 4 : 1) / bm_threshold_tables(h, topk, steps);
    return std::min<int64_t>(std::max<int64_t>(2048, est), 1 << 20);
}

// ---------------------------------------------------------------------------------------------------------
// mbarrier / TMA bulk copy (sm_90+ PTX; on sm_100a: SYNCS.* and UBLKCP.S.G)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)), "l"(policy)
                 : "memory");
}
// L2 residency: the bucket rows stream through once per unit (evict first); the query batch (nq x d, a few MB) is
// re-read by every unit and must not be washed out of the 126 MB L2 by the ~50 GB of rows (evict last)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// ---------------------------------------------------------------------------------------------------------
// k_score_stream
//
// Warp w of the grid walks units w, w + W, w + 2W, ... (W = warps in the grid; consecutive units — usually pieces of
// one run — land on the warps of one CTA at the same time, so a bucket staged twice is an L2 hit the second time).
// Each warp is its own producer and consumer of a 3-slot ring in shared memory.  The slot sequence of a unit is
//     [its queries 0..7] [its queries 8..15, if any] [bucket rows 0..7] [rows 8..15] ...
// every slot filled by up to 8 bulk copies (one row each, issued by 8 lanes) that complete on the slot's mbarrier.
// A query slot is unpacked into DMMA B fragments (registers); a row slot is multiplied against them: A fragments by
// LDS.128, 2 x 16 k-steps of DMMA.8x8x4 per n-block, one score per (pair, row) stored to the pair's segment.
// The warp runs the copies two slots ahead of the multiplication.  With ~128 KB of copies in flight per SM, any
// *synchronous* global load issued from this loop waits behind them for microseconds (measured: 4 us per unit for a
// demand load of the query rows), so nothing here loads from global memory: the unit record (descriptor, query
// list, score segments, first 32 row ids) and the later id windows also arrive by bulk copy, one unit / one
// window ahead.
// k permutation: DMMA k-step 2w takes columns 8w + 2t, k-step 2w + 1 columns 8w + 2t + 1, so that both fragment
// elements of a thread are adjacent in memory (one LDS.128 feeds two DMMAs); A and B use the same permutation, the
// sum over k is unchanged.
// ---------------------------------------------------------------------------------------------------------
// Store kinds: the vector store is kept in the narrowest type that represents every value exactly (store.cu):
// FP64 (E = 2 values per 16-byte chunk), FP32 (E = 4) or uint8 (E = 16).  Rows are widened to FP64 in registers
// (exactly) and multiplied on the FP64 tensor pipe, so a score is the same FP64 sum of the same products whatever
// the store kind; only the bytes a row costs in HBM change (8d, 4d or d).
// k permutation, all kinds: thread t of a DMMA row group takes the 16-byte chunks 4j + t of the row (j = 0, 1, ...);
// k-step j * E + e multiplies element e of those chunks, i.e. column (64j + 16t) / sizeof(T) + e.  A (rows) and B
// (queries) use the same permutation, so the sum over k is unchanged.  For FP64 this is the pairing "k-step 2w
// takes columns 8w + 2t, k-step 2w + 1 columns 8w + 2t + 1" (one LDS.128 feeds two DMMAs); for uint8 one LDS.128
// feeds 16.  Row pitch in a slot = 64 (mod 128) bytes: the LDS.128 of a quarter warp (2 rows x 4 chunks) is
// conflict-free.
template <int KIND> struct StoreKind;
template <> struct StoreKind<DPF_STORE_KIND_F64> { static constexpr int SZ = 8, E = 2, RP = SS_PITCH * 8; };
template <> struct StoreKind<DPF_STORE_KIND_F32> { static constexpr int SZ = 4, E = 4, RP = BM_KC * 4 + 64; };
template <> struct StoreKind<DPF_STORE_KIND_U8> { static constexpr int SZ = 1, E = 16, RP = BM_KC + 64; };

template <int KIND>
__device__ __forceinline__ void widen_chunk(const uint4& c, double (&a)[StoreKind<KIND>::E]) {
    if constexpr (KIND == DPF_STORE_KIND_F64) {
        a[0] = __hiloint2double((int)c.y, (int)c.x);
        a[1] = __hiloint2double((int)c.w, (int)c.z);
    } else if constexpr (KIND == DPF_STORE_KIND_F32) {
        a[0] = (double)__uint_as_float(c.x);
        a[1] = (double)__uint_as_float(c.y);
        a[2] = (double)__uint_as_float(c.z);
        a[3] = (double)__uint_as_float(c.w);
    } else {
        // byte v -> the double 2^52 + v (mantissa = v), minus 2^52: exact, one PRMT + one DADD per value
        const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                a[4 * i + b] = __hiloint2double(0x43300000, (int)__byte_perm(w[i], 0u, 0x4440u | (unsigned)b)) - 4503599627370496.0;
    }
}

template <bool ANGULAR, int KIND>
__global__ void __launch_bounds__(SS_WARPS * 32, 1)
k_score_stream(const unsigned char* __restrict__ X, unsigned row_bytes /* bytes per stored row, multiple of 16 */, int d,
               const double* __restrict__ Q, const UnitRec* __restrict__ units,
               const uint32_t* __restrict__ nunits_p, const int32_t* __restrict__ ids_sorted, Filter flt,
               unsigned long long* __restrict__ stat /* [0] units, [1] rows staged */) {
    using SK = StoreKind<KIND>;
    constexpr int E = SK::E;
    constexpr int NJ = BM_KC * SK::SZ / 64;         // 64-byte groups (4 chunks) per row
    constexpr int NS = NJ * E;                      // DMMA k-steps per row = BM_KC / 4
    constexpr int SLOT_DOUBLES = SS_ROWS * SS_PITCH;
    extern __shared__ __align__(128) unsigned char ssm_raw[];
    __shared__ uint64_t bars[SS_WARPS][SS_STAGES + 4];   // ring slots, 2 unit records, 2 id windows
    __shared__ int4 meta[SS_WARPS][SS_STAGES];           // per slot: {kind: 0/1 query block, 2 rows; first row; rows | m << 8; bucket start}
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    unsigned char* wbase = ssm_raw + (size_t)warp * ((SS_WARP_BYTES + 15) / 16 * 16);
    double* ring = reinterpret_cast<double*>(wbase);
    UnitRec* recs = reinterpret_cast<UnitRec*>(wbase + (size_t)SS_STAGES * SLOT_DOUBLES * sizeof(double));
    int32_t* wins = reinterpret_cast<int32_t*>(recs + 2);            // 2 x SS_WIN_COPY ids
    uint64_t* bar_slot = &bars[warp][0];
    uint64_t* bar_rec = &bars[warp][SS_STAGES];
    uint64_t* bar_win = &bars[warp][SS_STAGES + 2];
    for (int i = lane; i < SS_STAGES * SLOT_DOUBLES; i += 32) ring[i] = 0.0;
    if (lane == 0)
        for (int s = 0; s < SS_STAGES + 4; ++s) mbar_init(&bars[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();

    const int nj = (d * SK::SZ + 63) >> 6;              // 64-byte groups in use
    const bool ragged = (d * SK::SZ) & 63;              // the last group holds columns >= d: masked in registers
    const unsigned q_bytes = (unsigned)d * 8u;
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    const int64_t nunits = *nunits_p;
    const int64_t W = (int64_t)gridDim.x * SS_WARPS;
    const int64_t gw = (int64_t)blockIdx.x * SS_WARPS + warp;
    const int64_t nmine = nunits > gw ? (nunits - gw + W - 1) / W : 0;   // units gw + k * W, k = 0 .. nmine - 1
    if (nmine == 0) return;

    // ---- producer state (warp-uniform) ------------------------------------------------------------------------
    int64_t pk = 0;                    // unit being issued
    int p_phase = -1;                  // -1 unit not opened yet, 0 / 1 next query block, 2 rows
    uint32_t p_bstart = 0, p_len = 0, p_m = 0;
    int p_row = 0;
    unsigned win_uses0 = 0, win_uses1 = 0;   // completed waits on each id-window barrier (-> parity)
    int issued = 0, consumed = 0;
    unsigned long long rows_staged = 0;
    SurvivorSink sink;
    auto fetch_rec = [&](int64_t k) {
        if (k < nmine && lane == 0) {
            mbar_expect_tx(&bar_rec[k & 1], (unsigned)sizeof(UnitRec));
            bulk_g2s(&recs[k & 1], units + (gw + k * W), (unsigned)sizeof(UnitRec), &bar_rec[k & 1], pol_stream);
        }
    };
    // ids of rows [32j, 32j + 32) of the current bucket -> window buffer j & 1 (the copy starts at the 16-byte
    // boundary below and is 4 ids longer; ids_sorted is allocated with slack for the over-read)
    auto fetch_win = [&](int j) {
        if (lane == 0) {
            const uint32_t e0 = p_bstart + (uint32_t)(SS_WIN * j);
            mbar_expect_tx(&bar_win[j & 1], SS_WIN_COPY * 4);
            bulk_g2s(wins + (j & 1) * SS_WIN_COPY, ids_sorted + (e0 & ~3u), SS_WIN_COPY * 4, &bar_win[j & 1], pol_stream);
        }
    };
    fetch_rec(0);

    // issue one slot; false when the ring is full or the walk is over
    auto issue = [&]() -> bool {
        if (pk >= nmine || issued - consumed >= SS_STAGES) return false;
        if (p_phase < 0) {             // open unit pk: its record was requested one unit ago
            mbar_wait(&bar_rec[pk & 1], (unsigned)((pk >> 1) & 1));
            fetch_rec(pk + 1);
            const UnitRec* r = &recs[pk & 1];
            p_bstart = r->bstart; p_len = r->len; p_m = r->m;
            p_phase = 0; p_row = 0;
        }
        const UnitRec* r = &recs[pk & 1];
        const int s = issued % SS_STAGES;
        double* slot = ring + (size_t)s * SLOT_DOUBLES;
        if (p_phase < 2) {             // 8 query rows (always FP64)
            const int nrows = min(SS_ROWS, (int)p_m - SS_ROWS * p_phase);
            if (lane == 0) {
                meta[warp][s] = make_int4(p_phase, 0, nrows | ((int)p_m << 8), (int)p_bstart);
                mbar_expect_tx(&bar_slot[s], (unsigned)nrows * (q_bytes + 16u));
            }
            __syncwarp();
            if (lane < nrows) {
                const int j = SS_ROWS * p_phase + lane;
                const int qj = r->q[j];
                // the query's index and its threshold ride in the row's padding: the index by a plain store, the threshold
                // (with its 16-byte neighbour) by a bulk copy — no demand load from this loop
                reinterpret_cast<int32_t*>(slot + (size_t)lane * SS_PITCH + BM_KC)[0] = qj;
                bulk_g2s(slot + (size_t)lane * SS_PITCH + BM_KC + 2, flt.tau + (qj & ~1), 16u, &bar_slot[s], pol_keep);
                bulk_g2s(slot + (size_t)lane * SS_PITCH, Q + (int64_t)qj * d, q_bytes, &bar_slot[s], pol_keep);
            }
            p_phase = (p_phase == 0 && p_m > SS_ROWS) ? 1 : 2;
        } else {
            const int j = p_row / SS_WIN;                  // id window of this slot
            if (p_row % SS_WIN == 0) {
                if (j >= 1) {
                    if (j & 1) { mbar_wait(&bar_win[1], win_uses1 & 1); win_uses1++; }
                    else { mbar_wait(&bar_win[0], win_uses0 & 1); win_uses0++; }
                }
                if ((uint32_t)(SS_WIN * (j + 1)) < p_len) fetch_win(j + 1);   // one window (4 slots) ahead
            }
            const int nrows = min(SS_ROWS, (int)p_len - p_row);
            if (lane == 0) {
                meta[warp][s] = make_int4(2, p_row, nrows | ((int)p_m << 8), (int)p_bstart);
                mbar_expect_tx(&bar_slot[s], (unsigned)nrows * row_bytes);
            }
            __syncwarp();
            if (lane < nrows) {
                const int w = p_row % SS_WIN + lane;
                const int id = j == 0 ? r->ids0[w]
                                      : wins[(j & 1) * SS_WIN_COPY + (int)((p_bstart + (uint32_t)(SS_WIN * j)) & 3u) + w];
                bulk_g2s(reinterpret_cast<unsigned char*>(slot) + (size_t)lane * SK::RP, X + (int64_t)id * row_bytes, row_bytes,
                         &bar_slot[s], pol_stream);
            }
            p_row += nrows;
            rows_staged += nrows;
            if (p_row >= (int)p_len) { pk++; p_phase = -1; }
        }
        issued++;
        return true;
    };

    // ---- consumer state ---------------------------------------------------------------------------------------
    double B[2][NS];                    // queries g (n-block 0) and 8 + g (n-block 1), k-step order
    int c_q[2][2] = {{0, 0}, {0, 0}};
    double c_tau[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    bool c_ok[2][2] = {{false, false}, {false, false}};
    double c_qn[2][2] = {{1.0, 1.0}, {1.0, 1.0}};
    int c_nb = 0;                       // n-blocks in use
    const int col_t = 16 * t / SK::SZ;  // first column of this thread's chunk in group 0

    for (;;) {
        while (issue()) {}
        __syncwarp();                       // slot meta and score segments written by single lanes are visible
        if (issued == consumed) break;
        const int s = consumed % SS_STAGES;
        const int4 mt = meta[warp][s];
        const double* slot = ring + (size_t)s * SLOT_DOUBLES;
        mbar_wait(&bar_slot[s], (unsigned)((consumed / SS_STAGES) & 1));
        const int mt_rows = mt.z & 0xff, mt_m = mt.z >> 8;
        if (mt.x < 2) {
            // query block mt.x of a unit of mt_m queries -> B fragments (columns >= d are 0), query indices, thresholds, norms
            const double* br = slot + (size_t)g * SS_PITCH;
            if (mt.x == 0) c_nb = (mt_m + 7) >> 3;
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) {
                if (nb == mt.x) {
#pragma unroll
                    for (int j = 0; j < NJ; ++j)
#pragma unroll
                        for (int e = 0; e < E; e += 2) {
                            const int col = 64 * j / SK::SZ + col_t + e;          // even
                            double2 v = make_double2(0.0, 0.0);
                            if constexpr (KIND == DPF_STORE_KIND_F64) {
                                // FP64 rows and query rows share one slot layout: the k padding is never written, stays 0
                                if (j < nj) v = *reinterpret_cast<const double2*>(br + col);
                            } else {
                                if (j < nj && col < d) v = *reinterpret_cast<const double2*>(br + col);
                                if (col + 1 >= d) v.y = 0.0;
                            }
                            B[nb][j * E + e] = v.x;
                            B[nb][j * E + e + 1] = v.y;
                        }
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double* pad = slot + (size_t)(2 * t + e) * SS_PITCH + BM_KC;
                        c_ok[nb][e] = 8 * nb + 2 * t + e < mt_m;
                        c_q[nb][e] = c_ok[nb][e] ? reinterpret_cast<const int32_t*>(pad)[0] : 0;
                        c_tau[nb][e] = pad[2 + (c_q[nb][e] & 1)];
                    }
                    if (ANGULAR) {
                        double sq = 0.0;    // ||query 8nb + g||^2: this thread's columns, then the 4 threads of the group
#pragma unroll
                        for (int w = 0; w < NS; ++w) sq = fma(B[nb][w], B[nb][w], sq);
                        sq += __shfl_xor_sync(0xffffffffu, sq, 1);
                        sq += __shfl_xor_sync(0xffffffffu, sq, 2);
                        const double nrm = sqrt(sq);
#pragma unroll
                        for (int e = 0; e < 2; ++e) c_qn[nb][e] = __shfl_sync(0xffffffffu, nrm, (2 * t + e) * 4);
                    }
                }
            }
            __syncwarp();
            consumed++;
            continue;
        }
        const unsigned char* ar = reinterpret_cast<const unsigned char*>(slot) + (size_t)g * SK::RP + 16 * t;
        double acc[2][2][2];                // [n-block][even / odd k-step chain][column]
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
#pragma unroll
            for (int c = 0; c < 2; ++c) acc[nb][c][0] = acc[nb][c][1] = 0.0;
        double xn = 0.0;
        auto multiply = [&](auto two_blocks) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if (j < nj) {
                    double a[E];
                    widen_chunk<KIND>(*reinterpret_cast<const uint4*>(ar + 64 * j), a);
                    if constexpr (KIND != DPF_STORE_KIND_F64) {      // narrow rows: the slot padding holds query bytes
                        if (ragged && j == nj - 1) {
#pragma unroll
                            for (int e = 0; e < E; ++e)
                                if (64 * j / SK::SZ + col_t + e >= d) a[e] = 0.0;
                        }
                    }
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        dmma884(acc[0][e & 1][0], acc[0][e & 1][1], a[e], B[0][j * E + e]);
                        if (decltype(two_blocks)::value) dmma884(acc[1][e & 1][0], acc[1][e & 1][1], a[e], B[1][j * E + e]);
                        if (ANGULAR) xn = fma(a[e], a[e], xn);
                    }
                }
            }
        };
        if (c_nb == 2) multiply(std::true_type{});
        else multiply(std::false_type{});
        double xnr = 1.0;
        if (ANGULAR) {
            xn += __shfl_xor_sync(0xffffffffu, xn, 1);
            xn += __shfl_xor_sync(0xffffffffu, xn, 2);
            xnr = sqrt(xn);
        }
        // thread (g, t) holds (row mt.y + g, queries 8nb + 2t, 8nb + 2t + 1); only scores that reach the query's
        // threshold leave the kernel
        {
            const uint32_t pos = (uint32_t)mt.w + (uint32_t)(mt.y + g);
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double v = acc[nb][0][e] + acc[nb][1][e];
                    if (ANGULAR) v = v / (c_qn[nb][e] * xnr);
                    sink.push(flt, g < mt_rows && nb < c_nb && c_ok[nb][e] && v >= c_tau[nb][e], c_q[nb][e], pos, v, lane);
                }
        }
        __syncwarp();                       // every lane has read the slot before it is refilled
        consumed++;
    }
    sink.flush(flt, lane);
    if (lane == 0) { atomicAdd(&stat[0], (unsigned long long)nmine); atomicAdd(&stat[1], rows_staged); }
}

// ---------------------------------------------------------------------------------------------------------
// k_threshold — one CTA per query, one sampled table per warp: tau[q] = a score that k distinct candidates of the query are guaranteed to reach
// in the scoring kernel.  The warp scores the rows of the first bucket the query probes in each of its first NT
// tables (in practice its own bucket: the probe with the lowest bit flipped usually falls below the leaf's level) with
// plain FP64 FMAs, keeps the k best distinct rows by a LOWER BOUND of their score — score minus (more than) twice the rounding
// bound (d + 2) * 2^-53 * sum |x_c q_c| of a length-d dot product in any order, which covers the different summation
// order of the tensor-pipe kernels — and publishes the k-th.  Fewer than k rows seen: tau = -inf (keep everything).
// Also sets the query's survivor list: base = start of its score segment, cnt = 0.
// ---------------------------------------------------------------------------------------------------------
// 16 consecutive columns of a stored row: raw 16-byte vectors (loaded ahead) and their widening to doubles
template <int KIND> struct RawRow { static constexpr int N = KIND == DPF_STORE_KIND_U8 ? 1 : (KIND == DPF_STORE_KIND_F32 ? 4 : 8); };

template <int KIND>
__device__ __forceinline__ void raw_to_cols16(const uint4 (&raw)[RawRow<KIND>::N], double (&x)[16]) {
    if constexpr (KIND == DPF_STORE_KIND_U8) {
        const uint32_t w[4] = {raw[0].x, raw[0].y, raw[0].z, raw[0].w};
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = (double)((w[i >> 2] >> (8 * (i & 3))) & 0xffu);
    } else if constexpr (KIND == DPF_STORE_KIND_F32) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            x[4 * i] = (double)__uint_as_float(raw[i].x); x[4 * i + 1] = (double)__uint_as_float(raw[i].y);
            x[4 * i + 2] = (double)__uint_as_float(raw[i].z); x[4 * i + 3] = (double)__uint_as_float(raw[i].w);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[2 * i] = __hiloint2double((int)raw[i].y, (int)raw[i].x);
            x[2 * i + 1] = __hiloint2double((int)raw[i].w, (int)raw[i].z);
        }
    }
}

// INTQ: byte rows and byte queries (Q8, qnorm from k_quantise_queries): exact integer dot products with DP4A, the
// same value the integer tensor pipe produces, no rounding allowance needed for the dot product.
template <bool ANGULAR, int KIND, bool INTQ, bool REG /* K <= 32: the sample list lives in registers */>
__global__ void __launch_bounds__(RR_THREADS)
k_threshold(const unsigned char* __restrict__ X, unsigned row_bytes, int d, ChunkView cv, int q8_pitch, int gate,
            int L, int NT, const uint32_t* __restrict__ leaf_pos, const int32_t* __restrict__ leaf_len,
            const int32_t* __restrict__ ids_sorted, int self_exclude, int K, double* __restrict__ tl_keys,
            int* __restrict__ tl_ids, int* __restrict__ tl_cnt) {
    static_assert(!INTQ || KIND == DPF_STORE_KIND_U8, "integer path needs byte rows");
    // gate: 0 always; 1 only when the whole batch is byte vectors; 2 only when it is not (the host launches both forms
    // for a byte store and the batch's flag, still on the device, picks one)
    if (gate && ((*cv.q8_bad != 0) != (gate == 2))) return;
    const double* __restrict__ Q = cv.Q;
    const unsigned char* __restrict__ Q8 = cv.Q8;
    const double* __restrict__ qnorm8 = cv.qnorm8;
    const int32_t* __restrict__ qids = cv.qids;
    const int64_t nqc = cv.nqc;
    constexpr int RN = RawRow<KIND>::N;
    constexpr int SZ = KIND == DPF_STORE_KIND_U8 ? 1 : (KIND == DPF_STORE_KIND_F32 ? 4 : 8);
    constexpr int PFT = KIND == DPF_STORE_KIND_U8 ? 4 : 2;     // steps (of 4 rows) whose rows are in flight per warp
    extern __shared__ double rsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane >> 3, l8 = lane & 7;            // 4 rows per step, 8 lanes x 16 columns per row
    double* mykeys = rsm + (size_t)warp * K;
    int* myids = reinterpret_cast<int*>(rsm + (size_t)RR_WARPS * K) + (size_t)warp * K;
    // one warp per (query, sampled table): the dependent loads of a sample (pair -> bucket -> ids -> rows) overlap with
    // those of the other warps; the per-sample lists go to global memory and k_threshold_merge combines them
    const int64_t wid = (int64_t)blockIdx.x * RR_WARPS + warp;
    if (wid >= nqc * NT) return;
    const int64_t ql = wid / NT;
    const int sample = (int)(wid % NT);
    const int64_t q = ql;
    const int qid = qids ? qids[q] : INT32_MIN;
    const bool excl = self_exclude && qids && qid >= -128 && qid <= 127;
    const int c0 = 16 * l8;
    // a lane whose columns lie beyond the row reads the row's first bytes instead (a select on a loaded value would
    // make the load synchronous): its query values are 0
    const unsigned off = (unsigned)c0 * SZ < row_bytes ? (unsigned)c0 * SZ : 0u;
    double qv[INTQ ? 1 : 16];
    uint4 q8 = make_uint4(0, 0, 0, 0);
    double qn = 1.0;
    if constexpr (INTQ) {
        q8 = __ldg(reinterpret_cast<const uint4*>(Q8 + (size_t)q * q8_pitch + c0));     // zero beyond column d
        if (ANGULAR) qn = __ldg(qnorm8 + q);
    } else {
        double qq = 0.0;
#pragma unroll
        for (int i = 0; i < 16; i += 2) {                // d even, Q 16-byte aligned (bucket_major_supported)
            const double2 v = c0 + i < d ? __ldg(reinterpret_cast<const double2*>(Q + q * d + c0 + i)) : make_double2(0.0, 0.0);
            qv[i] = v.x; qv[i + 1] = v.y;
            qq = fma(v.x, v.x, qq); qq = fma(v.y, v.y, qq);
        }
        if (ANGULAR) {
            qq += __shfl_xor_sync(0xffffffffu, qq, 1);
            qq += __shfl_xor_sync(0xffffffffu, qq, 2);
            qq += __shfl_xor_sync(0xffffffffu, qq, 4);
            qn = sqrt(qq);
        }
    }
    const double slack = 4.0 * (double)(d + 8) * 1.1102230246251565e-16;     // >= twice the bound, with room for the norms
    // tables (among the first 32) in which the query probes something
    uint32_t nonempty = __ballot_sync(0xffffffffu, lane < L && cv.pair_cnt[ql * L + min(lane, L - 1)] > 0u);
    int count = 0;
    double kth = 0.0;                        // mykeys[K - 1] once the list is full
    RegList rl;
    for (int i = 0; i < sample && nonempty; ++i) nonempty &= nonempty - 1;
    // the sample-th table that has a pair: the first bucket the query probes there (in practice its own bucket) and, should
    // it hold fewer than K rows, the next ones until K rows have been seen
    const int t = nonempty ? __ffs(nonempty) - 1 : 0;
    const uint32_t nb_ = nonempty ? cv.pair_cnt[ql * L + t] : 0u;
    int seen = 0;
    for (uint32_t e_ = 0; e_ < nb_ && (e_ == 0 || seen < K); ++e_) {
        const uint32_t leaf = cv.cache[(ql * L + t) * cv.cap + e_];
        const uint32_t bstart = leaf_pos[leaf];
        const int len = leaf_len[leaf];
        seen += len;
        const int32_t* bids = ids_sorted + bstart;
        const int nmine = (len + 3) >> 2;    // steps of 4 rows; ids two groups of PFT steps ahead, rows one group ahead
        uint4 raw[PFT][RN];
        int id_row[PFT], id_ahead[PFT];
        auto load_id = [&](int k) { return __ldg(bids + min(4 * k + sub, len - 1)); };
        auto issue = [&](uint4 (&dst)[RN], int id) {
            const unsigned char* xr = X + (size_t)id * row_bytes + off;
#pragma unroll
            for (int v = 0; v < RN; ++v) {
                // vectors past the end of the row (d < 128) re-read the lane's first vector
                const unsigned o = off + 16u * v < row_bytes ? 16u * v : 0u;
                dst[v] = __ldg(reinterpret_cast<const uint4*>(xr + o));
            }
        };
#pragma unroll
        for (int s0 = 0; s0 < PFT; ++s0) id_row[s0] = s0 < nmine ? load_id(s0) : 0;
#pragma unroll
        for (int s0 = 0; s0 < PFT; ++s0) id_ahead[s0] = PFT + s0 < nmine ? load_id(PFT + s0) : 0;
#pragma unroll
        for (int s0 = 0; s0 < PFT; ++s0)
            if (s0 < nmine) issue(raw[s0], id_row[s0]);
        for (int base = 0; base < nmine; base += PFT) {
#pragma unroll
            for (int s0 = 0; s0 < PFT; ++s0) {
                const int k = base + s0;
                if (k >= nmine) break;                           // warp-uniform
                const int id = id_row[s0];
                const int row = 4 * k + sub;
                double lb;                   // lower bound of the score the scoring kernel will compute for this row
                if constexpr (INTQ) {
                    const uint4 x = raw[s0][0];
                    unsigned dot = 0, xx = 0;
                    dot = __dp4a(x.x, q8.x, dot); dot = __dp4a(x.y, q8.y, dot); dot = __dp4a(x.z, q8.z, dot); dot = __dp4a(x.w, q8.w, dot);
                    if (ANGULAR && (unsigned)c0 < row_bytes) { xx = __dp4a(x.x, x.x, xx); xx = __dp4a(x.y, x.y, xx); xx = __dp4a(x.z, x.z, xx); xx = __dp4a(x.w, x.w, xx); }
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        dot += __shfl_xor_sync(0xffffffffu, dot, o);
                        if (ANGULAR) xx += __shfl_xor_sync(0xffffffffu, xx, o);
                    }
                    lb = (double)dot;
                    if (ANGULAR) { lb = lb / (qn * sqrt((double)xx)); lb -= 8.0 * 1.1102230246251565e-16 * fabs(lb); }
                } else {
                    double x[16];
                    raw_to_cols16<KIND>(raw[s0], x);
                    double dot = 0.0, ab = 0.0, xx = 0.0;
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        dot = fma(x[c], qv[c], dot);
                        ab = fma(fabs(x[c]), fabs(qv[c]), ab);
                        if (ANGULAR && c0 + c < d) xx = fma(x[c], x[c], xx);
                    }
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        dot += __shfl_xor_sync(0xffffffffu, dot, o);
                        ab += __shfl_xor_sync(0xffffffffu, ab, o);
                        if (ANGULAR) xx += __shfl_xor_sync(0xffffffffu, xx, o);
                    }
                    if (ANGULAR) {
                        const double den = qn * sqrt(xx);
                        lb = dot / den - slack * (ab / den);
                    } else {
                        lb = dot - slack * ab;
                    }
                }
                // next rows / ids of this ring slot
                if (k + PFT < nmine) {
                    id_row[s0] = id_ahead[s0];
                    issue(raw[s0], id_row[s0]);
                    if (k + 2 * PFT < nmine) id_ahead[s0] = load_id(k + 2 * PFT);
                }
                // rows that can enter the list (rare once it is full): one after the other, all lanes take part
                bool ok = l8 == 0 && row < len && lb == lb && !(excl && id == qid);
                if (ok && count == K) ok = lb >= kth;
                uint32_t todo = __ballot_sync(0xffffffffu, ok);
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const double lb_s = __shfl_sync(0xffffffffu, lb, src);
                    const int id_s = __shfl_sync(0xffffffffu, id, src);
                    if (REG) {
                        if (!rl.contains(id_s, lane)) rl.insert(K, lb_s, id_s, lane);
                        count = rl.count;
                        if (count == K) kth = __shfl_sync(0xffffffffu, rl.key, K - 1);
                        continue;
                    }
                    if (count == K && !better(lb_s, id_s, mykeys[K - 1], myids[K - 1])) continue;
                    bool dup = false;        // the same row reached through another table: keep it once
                    for (int b2 = 0; b2 < count; b2 += 32) dup |= __any_sync(0xffffffffu, b2 + lane < count && myids[b2 + lane] == id_s);
                    if (!dup) warp_insert(mykeys, myids, count, K, lb_s, id_s, lane);
                    if (count == K) kth = mykeys[K - 1];
                }
            }
        }
    }
    __syncwarp();
    if (REG) {
        if (lane < count) { tl_keys[wid * K + lane] = rl.key; tl_ids[wid * K + lane] = rl.id; }
    } else {
        for (int r = lane; r < count; r += 32) {
            tl_keys[wid * K + r] = mykeys[r];
            tl_ids[wid * K + r] = myids[r];
        }
    }
    if (lane == 0) tl_cnt[wid] = count;
}

// one warp per query: tau = k-th best distinct row over its NT sample lists (a row sampled through two tables has the
// same bound in both lists, so duplicates are adjacent in the merged order).  Fewer than k distinct rows in the samples:
// no threshold can be guaranteed.  A query that visits few entries anyway (<= SMALL_QUERY_ENTRIES on this rank: common on
// a multi-GPU shard, where a rank holds one of eight sub-indexes) simply keeps them all (tau = -inf); a heavy one is
// flagged for k_topk_direct and tau = +inf keeps it out of the pool.
constexpr uint32_t SMALL_QUERY_ENTRIES = 4096;
__global__ void __launch_bounds__(RR_THREADS)
k_threshold_merge(int64_t nqc, int NT, int K, const double* __restrict__ tl_keys, const int* __restrict__ tl_ids,
                  const int* __restrict__ tl_cnt, const uint32_t* __restrict__ q_entries, double* __restrict__ tau,
                  int32_t* __restrict__ taui, DirtySet dirty) {
    const int lane = threadIdx.x & 31;
    const int64_t ql = (int64_t)blockIdx.x * RR_WARPS + (threadIdx.x >> 5);
    if (blockIdx.x == 0 && threadIdx.x == 0) taui[nqc] = 0x7fffffff;      // sentinel: "no query" for the tcgen05 kernel's padding
    if (ql >= nqc) return;
    const int64_t q = ql;
    const double* lkeys = tl_keys + (ql * NT + lane) * K;
    const int* lids = tl_ids + (ql * NT + lane) * K;
    const int mycount = lane < NT ? tl_cnt[ql * NT + lane] : 0;
    int head = 0, last = -1, found = 0;
    double kth = 0.0;
    while (found < K) {
        double bk = 0.0;
        int bi = 0x7fffffff, bl = -1;
        if (head < mycount) { bk = lkeys[head]; bi = lids[head]; bl = lane; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ok_ = __shfl_xor_sync(0xffffffffu, bk, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
            if (ol >= 0 && (bl < 0 || better(ok_, oi, bk, bi) || (ok_ == bk && oi == bi && ol < bl))) { bk = ok_; bi = oi; bl = ol; }
        }
        if (bl < 0) break;                   // lists exhausted
        if (lane == bl) head++;
        if (bi != last) { found++; kth = bk; last = bi; }
    }
    if (lane == 0) {
        const bool small = q_entries[q] <= SMALL_QUERY_ENTRIES;
        const double tv = found == K ? kth : (small ? -__longlong_as_double(0x7ff0000000000000LL) : __longlong_as_double(0x7ff0000000000000LL));
        tau[q] = tv;
        // the same threshold for integer scores (a dot product of bytes is below 2^31 - 1: INT_MAX masks the query)
        taui[q] = tv >= 2147483647.0 ? 0x7fffffff : (tv <= -2147483648.0 ? (int)0x80000000 : (int)ceil(tv));
        if (found < K && !small) dirty.mark((int)q);
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_select_survivors — one warp per query: top k of its survivor list, larger score first, ties by smaller id; an id
// reached through several tables appears several times with bit-identical scores and is kept once.
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t SEL_BIG = 1024;

template <bool REG /* K <= 32: the list lives in registers */>
__global__ void __launch_bounds__(RR_THREADS)
k_select_survivors(int64_t q0, int64_t nqc, Filter flt, const int32_t* __restrict__ qids, int self_exclude, int K, bool negate,
                   int32_t* __restrict__ ids_out, double* __restrict__ score_out, unsigned long long* __restrict__ stat,
                   uint32_t* __restrict__ big_list, uint32_t* __restrict__ big_count) {
    extern __shared__ double rsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* mykeys = rsm + (size_t)warp * K;
    int* myids = reinterpret_cast<int*>(rsm + (size_t)RR_WARPS * K) + (size_t)warp * K;
    const int64_t ql = (int64_t)blockIdx.x * RR_WARPS + warp;
    if (ql >= nqc) return;
    const int64_t q = q0 + ql;
    if (flt.dirty.flag[q]) return;           // answered by k_topk_direct
    const uint32_t n = flt.cnt[q];
    if (lane == 0) atomicAdd(&stat[2], (unsigned long long)n);
    if (n > SEL_BIG) {                       // a weak threshold left a long list: k_select_survivors_big (a CTA per query)
        if (lane == 0) big_list[atomicAdd(big_count, 1u)] = (uint32_t)ql;
        return;
    }
    const double* sc = flt.s_score + flt.base[q];
    const int32_t* si = flt.s_id + flt.base[q];
    const int qid = qids ? qids[q] : INT32_MIN;
    const bool excl = self_exclude && qids && qid >= -128 && qid <= 127;
    int count = 0;
    RegList rl;
    double kth_k = 0.0;
    int kth_i = 0;
    for (uint32_t j0 = 0; j0 < n; j0 += 32) {
        const uint32_t j = j0 + lane;
        const double key = j < n ? sc[j] : 0.0;
        const int id = j < n ? si[j] : -1;
        bool cand = j < n && !(excl && id == qid);
        if (REG) { if (cand && count == K) cand = better(key, id, kth_k, kth_i); }
        else if (cand && count == K) cand = better(key, id, mykeys[K - 1], myids[K - 1]);
        uint32_t todo = __ballot_sync(0xffffffffu, cand);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const double kk = __shfl_sync(0xffffffffu, key, src);
            const int ii = __shfl_sync(0xffffffffu, id, src);
            if (REG) {
                if (!rl.contains(ii, lane)) rl.insert(K, kk, ii, lane);
                count = rl.count;
                if (count == K) rl.kth(K, kth_k, kth_i);
            } else {
                bool dup = false;
                for (int base = 0; base < count; base += 32) dup |= __any_sync(0xffffffffu, base + lane < count && myids[base + lane] == ii);
                if (!dup) warp_insert(mykeys, myids, count, K, kk, ii, lane);
            }
        }
    }
    if (REG) {
        if (lane < K) {
            ids_out[q * K + lane] = lane < count ? rl.id : -1;
            score_out[q * K + lane] = lane < count ? (negate ? -rl.key : rl.key) : __longlong_as_double(0x7ff8000000000000LL);
        }
    } else {
        for (int r = lane; r < K; r += 32) {
            ids_out[q * K + r] = r < count ? myids[r] : -1;
            score_out[q * K + r] = r < count ? (negate ? -mykeys[r] : mykeys[r]) : __longlong_as_double(0x7ff8000000000000LL);
        }
    }
}

// the same for the queries whose list is long (few: those whose sampled buckets held fewer than k rows, or rows far
// from the query): one CTA per query, the warps take interleaved chunks of the list, then the per-warp lists are merged
template <bool REG>
__global__ void __launch_bounds__(RR_THREADS)
k_select_survivors_big(int64_t q0, Filter flt, const int32_t* __restrict__ qids, int self_exclude, int K, bool negate,
                       int32_t* __restrict__ ids_out, double* __restrict__ score_out, const uint32_t* __restrict__ big_list,
                       const uint32_t* __restrict__ big_count) {
    extern __shared__ double rsm[];
    __shared__ int s_counts[RR_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nbig = *big_count;
    for (uint32_t bi_ = blockIdx.x; bi_ < nbig; bi_ += gridDim.x) {  // the long lists k_select_survivors set aside
    const int64_t q = q0 + big_list[bi_];
    const uint32_t n = flt.cnt[q];
    __syncthreads();                         // the lists in shared memory are reused
    double* lkeys = rsm;
    int* lids = reinterpret_cast<int*>(rsm + (size_t)RR_WARPS * K);
    double* mykeys = lkeys + (size_t)warp * K;
    int* myids = lids + (size_t)warp * K;
    const double* sc = flt.s_score + flt.base[q];
    const int32_t* si = flt.s_id + flt.base[q];
    const int qid = qids ? qids[q] : INT32_MIN;
    const bool excl = self_exclude && qids && qid >= -128 && qid <= 127;
    int count = 0;
    RegList rl;
    double kth_k = 0.0;
    int kth_i = 0;
    for (uint32_t j0 = 32u * warp; j0 < n; j0 += 32u * RR_WARPS) {
        const uint32_t j = j0 + lane;
        const double key = j < n ? sc[j] : 0.0;
        const int id = j < n ? si[j] : -1;
        bool cand = j < n && !(excl && id == qid);
        if (REG) { if (cand && count == K) cand = better(key, id, kth_k, kth_i); }
        else if (cand && count == K) cand = better(key, id, mykeys[K - 1], myids[K - 1]);
        uint32_t todo = __ballot_sync(0xffffffffu, cand);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const double kk = __shfl_sync(0xffffffffu, key, src);
            const int ii = __shfl_sync(0xffffffffu, id, src);
            if (REG) {
                if (!rl.contains(ii, lane)) rl.insert(K, kk, ii, lane);
                count = rl.count;
                if (count == K) rl.kth(K, kth_k, kth_i);
            } else {
                bool dup = false;
                for (int base = 0; base < count; base += 32) dup |= __any_sync(0xffffffffu, base + lane < count && myids[base + lane] == ii);
                if (!dup) warp_insert(mykeys, myids, count, K, kk, ii, lane);
            }
        }
    }
    if (REG && lane < count) { mykeys[lane] = rl.key; myids[lane] = rl.id; }   // the merge below reads the lists from shared memory
    if (lane == 0) s_counts[warp] = count;
    __syncthreads();
    if (warp == 0) {                         // merge; the same id in two lists carries the same score: adjacent, kept once
        int head = 0, last = -1;
        const int mycount = lane < RR_WARPS ? s_counts[lane] : 0;
        for (int r = 0; r < K; ++r) {
            double bk;
            int bi, bl;
            for (;;) {
                bk = 0; bi = 0x7fffffff; bl = -1;
                if (lane < RR_WARPS && head < mycount) { bk = lkeys[lane * K + head]; bi = lids[lane * K + head]; bl = lane; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ok_ = __shfl_xor_sync(0xffffffffu, bk, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
                    if (ol >= 0 && (bl < 0 || better(ok_, oi, bk, bi) || (ok_ == bk && oi == bi && ol < bl))) { bk = ok_; bi = oi; bl = ol; }
                }
                if (bl < 0) break;
                if (lane == bl) head++;
                if (bi != last) break;
            }
            if (lane == 0) {
                ids_out[q * K + r] = bl >= 0 ? bi : -1;
                score_out[q * K + r] = bl >= 0 ? (negate ? -bk : bk) : __longlong_as_double(0x7ff8000000000000LL);
            }
            if (bl >= 0) last = bi;
        }
    }
    }
}

// the compact store's rows when the build found a narrower lossless type (store.cu), else the FP64 rows
static void store_rows(const dpf_index* h, int& kind, const unsigned char*& rows, unsigned& row_bytes) {
    kind = h->Xc_kind;
    rows = kind == DPF_STORE_KIND_F64 ? reinterpret_cast<const unsigned char*>(h->Xdev) : h->Xc.p;
    row_bytes = kind == DPF_STORE_KIND_F64 ? (unsigned)h->cfg.d * 8u : (unsigned)h->Xc_row_bytes;
}

template <class F>
static void dispatch_kind(int kind, bool ang, F&& f) {
    auto k = [&](auto kc) {
        if (ang) f(std::true_type{}, kc); else f(std::false_type{}, kc);
    };
    if (kind == DPF_STORE_KIND_U8) k(std::integral_constant<int, DPF_STORE_KIND_U8>{});
    else if (kind == DPF_STORE_KIND_F32) k(std::integral_constant<int, DPF_STORE_KIND_F32>{});
    else k(std::integral_constant<int, DPF_STORE_KIND_F64>{});
}


// One chunk of a query batch, bucket-major: probe -> pairs grouped by bucket -> units -> thresholds -> scoring (only
// scores that can still be among a query's best k leave the kernel) -> per-query lists -> top k.  Nothing is read back
// by the host: scratch is sized by worst cases known up front, counts stay on the device, and the few queries the
// filter cannot serve are answered by k_topk_direct.  Every per-query array below is indexed inside the chunk.
void topk_bucket_major(dpf_index* h, const double* Qd, const QueryKeys& qk, int steps, int probe_mode, int64_t q0, int64_t q1,
                       int cap, int topk, int metric, int32_t* ids_out, double* score_out) {
    cudaStream_t st = h->stream;
    const int64_t nqc = q1 - q0;
    const int L = h->cfg.L, d = h->cfg.d;
    if (nqc <= 0) return;
    const bool ang = metric == DPF_METRIC_ANGULAR;
    const bool l2 = metric == DPF_METRIC_L2;                    // byte pipeline only; keys are -distance, negated on output
    const bool wide = d > BM_KC;                                // k_score_wide / k_threshold_wide on the FP64 rows
    const bool use_u8 = !wide && score_u8_usable(h);
    int kind;
    const unsigned char* rows;
    unsigned row_bytes;
    store_rows(h, kind, rows, row_bytes);
    int32_t* ctr = h->counters.p;

    // ---- scratch, all worst case ------------------------------------------------------------------------------------
    const int64_t pairs_ub = nqc * L * cap;
    const int64_t units_ub = std::min<int64_t>(pairs_ub, (int64_t)h->num_leaves + pairs_ub / SS_UQ) + 1;
    h->bm_units.reserve((size_t)units_ub * sizeof(UnitRec));
    // byte rows x byte queries, dot product: the tcgen05 kernel with its own, wider units
    const bool try_int = use_u8 && h->dbg[DPF_DBG_U8_IMMA] != 0 && (int64_t)d * 255 * 255 < (1LL << 31);
    // its units: (<= 32 pairs of a leaf) x (<= 128 of its rows); worst case from the forest's largest leaf and tile count
    // Every probed leaf contributes one unit per 128 of its rows, plus one per further group of 32 pairs: the record
    // array holds that with room to spare; the true worst case (every group of every leaf multiplied by the largest
    // leaf's tile count) is never approached, and a unit that does not fit hands its queries to k_topk_direct.
    const int64_t max_tiles = std::max<int64_t>(1, (h->max_leaf_len + 127) / 128);
    const int64_t tc_cap = std::min(pairs_ub / TC_TQ * max_tiles + h->total_leaf_tiles, h->total_leaf_tiles + pairs_ub / 16 + 1024) + 1;
    const bool use_tc = try_int && score_u8t_usable(h, metric) && tc_cap < (1LL << 31);
    if (use_tc) h->bm_descs.reserve((size_t)tc_cap * sizeof(TcRec));
    h->bm_taui.reserve((size_t)nqc + 2);
    int64_t pool_cap = h->dbg[DPF_DBG_POOL_RECORDS] > 0 ? h->dbg[DPF_DBG_POOL_RECORDS]
                                                        : std::min<int64_t>(std::max<int64_t>(nqc * bm_pool_per_query(h, topk, steps), 1 << 20), kMaxPool);
    pool_cap = std::max<int64_t>(SURV_BLOCK, pool_cap / SURV_BLOCK * SURV_BLOCK);
    h->surv_pool.reserve((size_t)pool_cap * sizeof(SurvRec));
    h->scores.reserve((size_t)pool_cap);
    h->surv_id.reserve((size_t)pool_cap);
    h->bm_tau.reserve((size_t)nqc + 2);
    // one block cleared per chunk: survivor counts, list cursors, dirty flags
    h->bm_scnt.reserve((size_t)nqc * 5 + 4);
    uint32_t* s_cnt = h->bm_scnt.p;
    uint32_t* s_fill = s_cnt + nqc;
    uint32_t* s_dirty = s_fill + nqc;
    uint32_t* q_entries = s_dirty + nqc;
    const DirtySet dirty{s_dirty, reinterpret_cast<int32_t*>(q_entries + nqc), ctr + CTR_NDIRTY};   // (the list needs no clearing)
    h->bm_sbase.reserve((size_t)nqc + 1);
    DPF_CUDA(cudaMemsetAsync(s_cnt, 0, (size_t)nqc * 4 * sizeof(uint32_t), st));
    DPF_CUDA(cudaMemsetAsync(ctr + CTR_POOL, 0, 7 * sizeof(int32_t), st));      // pool cursor, chunk counts, overflow flag, dirty count

    probe_leaves(h, qk, steps, probe_mode, q0, nqc, cap, q_entries);

    ChunkView cv;
    cv.Q = Qd + q0 * d;
    cv.Q8 = h->Q8.p ? h->Q8.p + q0 * u8_query_pitch() : nullptr;
    cv.qnorm8 = h->qnorm8.p ? h->qnorm8.p + q0 : nullptr;
    cv.qsq8 = h->qsq8.p ? h->qsq8.p + q0 : nullptr;
    cv.qids = qk.qids ? qk.qids + q0 : nullptr;
    cv.nqc = nqc;
    cv.q8_bad = ctr + CTR_Q8_BAD;
    cv.pair_cnt = h->pair_cnt.p;
    cv.cache = h->probe_cache.p;
    cv.cap = cap;
    const Filter flt{h->bm_tau.p, s_cnt, h->bm_sbase.p, s_fill, dirty, h->scores.p, h->surv_id.p,
                     reinterpret_cast<SurvRec*>(h->surv_pool.p), reinterpret_cast<uint32_t*>(ctr + CTR_POOL), (uint32_t)pool_cap,
                     ctr + CTR_POOL_OVERFLOW};
    const size_t list_smem = (size_t)RR_WARPS * topk * (sizeof(double) + sizeof(int));
    const unsigned qgrid = (unsigned)((nqc + RR_WARPS - 1) / RR_WARPS);
    // where the threshold stream forks: right after the probe (thresholds beside scan / fill / emit; default) or after the pair
    // fill (beside the unit records only).  Measured with tools/rank_emul.py --fork: the same 2.19-2.23 ms on one GPU; on a
    // shard of 4 / 8 GPUs the early fork gives 0.95-0.97 / 0.72-0.76 ms per local step against 0.99-1.04 / 0.77-0.78.
    const bool fork_early = h->dbg[DPF_DBG_TAU_FORK] != 2;

    // ---- thresholds: a row gather, on the handle's second stream beside the grouping chain (scan, pair fill, unit records)
    //      on the main one ------------------------------------------------------------------------------------------------
    {
        const int NT = bm_threshold_tables(h, topk, steps);
        h->bm_tl_keys.reserve((size_t)nqc * NT * topk);
        h->bm_tl_ids.reserve((size_t)nqc * NT * topk);
        h->bm_tl_cnt.reserve((size_t)nqc * NT);
        cudaStream_t st2 = h->aux_stream;
        if (!fork_early) group_pairs(h, nqc, cap, use_tc);
        DPF_CUDA(cudaEventRecord(h->ev_fork, st));
        DPF_CUDA(cudaStreamWaitEvent(st2, h->ev_fork, 0));
        StageTimer tmt(h, DPF_T_THRESHOLD, st2);
        const unsigned tgrid = (unsigned)((nqc * NT + RR_WARPS - 1) / RR_WARPS);
        auto go = [&](auto kern, int gate) {
            kern<<<tgrid, RR_THREADS, list_smem, st2>>>(rows, row_bytes, d, cv, u8_query_pitch(), gate, L, NT, h->leaf_pos.p,
                                                         h->leaf_len.p, h->ids_sorted.p, h->cfg.self_exclude_small_ids, topk,
                                                         h->bm_tl_keys.p, h->bm_tl_ids.p, h->bm_tl_cnt.p);
            DPF_LAUNCHED();
        };
        if (wide) {
            launch_threshold_wide(h, st2, metric, cv, NT, topk, list_smem);
        } else if (use_u8) {
            // byte rows: the integer form when the whole batch is bytes, the FP64 form otherwise; which one applies is a
            // flag on the device, so both are launched and one of them returns at once
            if (try_int) {
                if (h->dbg[DPF_DBG_TAU_KERNEL] != 1 || l2) launch_threshold_u8i(h, st2, metric, cv, NT, topk, list_smem);
                else if (ang) { if (topk <= 32) go(k_threshold<true, DPF_STORE_KIND_U8, true, true>, 1); else go(k_threshold<true, DPF_STORE_KIND_U8, true, false>, 1); }
                else { if (topk <= 32) go(k_threshold<false, DPF_STORE_KIND_U8, true, true>, 1); else go(k_threshold<false, DPF_STORE_KIND_U8, true, false>, 1); }
            }
            if (!l2) {
                const int gate = try_int ? 2 : 0;
                if (ang) { if (topk <= 32) go(k_threshold<true, DPF_STORE_KIND_U8, false, true>, gate); else go(k_threshold<true, DPF_STORE_KIND_U8, false, false>, gate); }
                else { if (topk <= 32) go(k_threshold<false, DPF_STORE_KIND_U8, false, true>, gate); else go(k_threshold<false, DPF_STORE_KIND_U8, false, false>, gate); }
            }
        } else {
            dispatch_kind(kind, ang, [&](auto a, auto kc) {
                if (topk <= 32) go(k_threshold<decltype(a)::value, decltype(kc)::value, false, true>, 0);
                else go(k_threshold<decltype(a)::value, decltype(kc)::value, false, false>, 0);
            });
        }
        k_threshold_merge<<<qgrid, RR_THREADS, 0, st2>>>(nqc, NT, topk, h->bm_tl_keys.p, h->bm_tl_ids.p, h->bm_tl_cnt.p, q_entries,
                                                         h->bm_tau.p, h->bm_taui.p, dirty); DPF_LAUNCHED();
        DPF_CUDA(cudaGetLastError());
    }
    DPF_CUDA(cudaEventRecord(h->ev_join, h->aux_stream));
    if (fork_early) group_pairs(h, nqc, cap, use_tc);    // the main stream meanwhile: pairs grouped by leaf, unit records
    if (use_tc) emit_tc_recs(h, tc_cap, dirty);
    emit_units(h, use_tc);               // with the tcgen05 kernel the records only serve a batch that is not byte vectors

    unsigned long long* bm_stat = reinterpret_cast<unsigned long long*>(ctr + CTR_BM_STAT);
    const UnitRec* units = reinterpret_cast<const UnitRec*>(h->bm_units.p);
    const uint32_t* nunits_p = reinterpret_cast<const uint32_t*>(ctr + CTR_NUNITS);
    DPF_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));            // thresholds from the second stream
    {
        StageTimer tm(h, DPF_T_RERANK);
        if (wide) {
            launch_score_wide(h, cv, units, nunits_p, metric, flt, bm_stat);
        } else if (use_u8) {
            if (use_tc)
                launch_score_u8t(h, cv, reinterpret_cast<const TcRec*>(h->bm_descs.p),
                                 reinterpret_cast<const uint32_t*>(ctr + CTR_NUNITS_TC), tc_cap, h->bm_taui.p, flt, bm_stat);
            launch_score_u8(h, cv, units, nunits_p, metric, flt, bm_stat, try_int && !use_tc);
        } else {
            dispatch_kind(kind, ang, [&](auto a, auto kc) {
                auto kern = k_score_stream<decltype(a)::value, decltype(kc)::value>;
                DPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SS_SMEM));
                kern<<<h->num_sms, SS_WARPS * 32, SS_SMEM, st>>>(rows, row_bytes, d, cv.Q, units, nunits_p, h->ids_sorted.p, flt, bm_stat);
                DPF_LAUNCHED();
            });
        }
        DPF_CUDA(cudaGetLastError());
    }
    {
        StageTimer tm(h, DPF_T_SELECT);
        survivor_lists(h, flt, nqc);
        h->bm_big.reserve((size_t)nqc + 1);                      // [0] = count, then the queries with long lists
        DPF_CUDA(cudaMemsetAsync(h->bm_big.p, 0, sizeof(uint32_t), st));
        int32_t* io = ids_out + q0 * topk;
        double* so = score_out + q0 * topk;
        if (topk <= 32)
            k_select_survivors<true><<<qgrid, RR_THREADS, list_smem, st>>>(0, nqc, flt, cv.qids, h->cfg.self_exclude_small_ids, topk, l2, io,
                                                                           so, bm_stat, h->bm_big.p + 1, h->bm_big.p);
        else
            k_select_survivors<false><<<qgrid, RR_THREADS, list_smem, st>>>(0, nqc, flt, cv.qids, h->cfg.self_exclude_small_ids, topk, l2, io,
                                                                            so, bm_stat, h->bm_big.p + 1, h->bm_big.p);
        DPF_LAUNCHED();
        if (topk <= 32)
            k_select_survivors_big<true><<<(unsigned)std::min<int64_t>(nqc, 1024), RR_THREADS, list_smem, st>>>(
                0, flt, cv.qids, h->cfg.self_exclude_small_ids, topk, l2, io, so, h->bm_big.p + 1, h->bm_big.p);
        else
            k_select_survivors_big<false><<<(unsigned)std::min<int64_t>(nqc, 1024), RR_THREADS, list_smem, st>>>(
                0, flt, cv.qids, h->cfg.self_exclude_small_ids, topk, l2, io, so, h->bm_big.p + 1, h->bm_big.p);
        DPF_LAUNCHED();
        topk_direct(h, cv, dirty, topk, metric, io, so);
        DPF_CUDA(cudaGetLastError());
    }
}

}  // namespace dpf
