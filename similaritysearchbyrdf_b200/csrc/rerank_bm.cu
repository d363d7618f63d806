// rerank_bm.cu — K5b: bucket-major re-rank on the FP64 tensor pipe.
//
// Replaces topKAndPrecisionScore's gather + dgemv + argsort (src/main/scala/mclab/deploy/DensevectorRDFInit.scala:
// 472-507) for a whole query batch.  The row-major kernel in query.cu reads a candidate row once per (query,
// candidate): 8d bytes of HBM for 2d flops.  In a batch, many queries probe the same leaf buckets (a bucket is
// probed by every query whose key falls in it or one bit away), so the same rows are fetched again and again.
// Here the batch is regrouped by bucket:
//   k_probe_pairs    one (bucket, query) pair per distinct bucket a query probes (same walk as K4)
//   radix sort       pairs by bucket start
//   k_run_flags .. k_emit_units   runs of pairs that share a bucket, cut into units of <= 16 queries; one
//                    self-contained record per unit (bucket, query list, score segments, first row ids)
//   k_score_stream   a warp per unit: the unit's queries and then the bucket's rows stream L2/HBM -> shared memory
//                    through a per-warp ring of TMA bulk copies (one 8d-byte copy per row, mbarrier completion, 2
//                    slots of 8 rows in flight per warp while a third is consumed); queries become DMMA B fragments
//                    in registers, rows are multiplied against them -> one score per (pair, row)
//   k_select_pairs   per query: top-k over its score segments, de-duplicating ids reached through several tables
//                    (their scores are bit-identical: same row, same query, same k order)
// Per step this replaces nC_q * 8d bytes per query by ~(bucket rows * 8d) per <= 16 queries plus 16 B per (query,
// candidate).  Candidate *sets* are unchanged (same probe), so results equal the row-major path up to summation
// order.  Why TMA rows: tools/gather_patterns.cu — a DMMA A-fragment gather straight from HBM touches 8 rows x 64 B
// per instruction and collapses to 2.9 TB/s at high occupancy; 1 KB bulk row copies hold 7.3 TB/s with 8 warps/SM.
#include <cstdlib>
#include <type_traits>

#include "rerank_units.cuh"

namespace dpf {

constexpr int BM_QT = 32;              // pairs per group of the register-gather kernel (k_score_warps)

bool bucket_major_supported(const dpf_index* h, int metric, int topk) {
    const char* e = getenv("DPF_RERANK");
    if (e && e[0] == 'r') return false;                        // DPF_RERANK=rowmajor forces the row-major kernel
    return h->dense && h->Xdev && h->cfg.d <= BM_KC && (h->cfg.d % 2) == 0 &&
           (reinterpret_cast<uintptr_t>(h->Xdev) & 15) == 0 &&
           (metric == DPF_METRIC_DOT || metric == DPF_METRIC_ANGULAR) && topk <= RR_MAXK;   // d even: the queries are FP64 rows
}

// fill pass: pair i of (query q, table t) in (q, t, lane) order
__global__ void __launch_bounds__(256)
k_probe_pairs(ProbeCtx c, const int32_t* __restrict__ qkeys, const uint8_t* __restrict__ qpids, int64_t ld, int64_t q0,
              int64_t nqc, const uint32_t* __restrict__ pair_base /* (q - q0) * L + t */,
              unsigned long long* __restrict__ pair_key, int32_t* __restrict__ pair_q, uint32_t* __restrict__ pair_len) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= nqc * c.L) return;
    const int64_t q = q0 + wid / c.L;
    const int t = (int)(wid % c.L);
    const uint32_t h = (uint32_t)qkeys[(int64_t)t * ld + q];
    const int pid = qpids[(int64_t)t * ld + q];
    const int seg = c.tp.seg_bits ? (int)(h >> c.tp.bucket_bits) : 0;
    const int nprobes = probe_count(h, c.probe_mode);
    if (nprobes < 0) return;
    uint32_t at = pair_base[wid];
    const long long tbase = c.f.table_base[t];
    const int np = 1 << c.tp.pb;
    for (int sub = 0; sub < np; ++sub) {
        if (__popc(sub ^ pid) > c.steps) continue;
        if (c.world > 1 && (sub % c.world) != c.rank) continue;
        bool leader;
        int ptr, cnt;
        warp_lookup(c, t, sub, seg, h, nprobes, lane, leader, ptr, cnt);
        const uint32_t m = __ballot_sync(0xffffffffu, leader);
        if (leader) {
            const uint32_t i = at + __popc(m & ((1u << lane) - 1u));
            pair_key[i] = ((unsigned long long)(tbase + ptr) << 32) | i;   // sort key: bucket start; payload: pair index
            pair_q[i] = (int32_t)q;
            pair_len[i] = (uint32_t)cnt;
        }
        at += __popc(m);
    }
}

// flag[p] = 1 where a run starts
__global__ void __launch_bounds__(256)
k_run_flags(const unsigned long long* __restrict__ sorted, int64_t npairs, uint32_t* __restrict__ flag) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    flag[p] = (p == 0 || (uint32_t)(sorted[p - 1] >> 32) != (uint32_t)(sorted[p] >> 32)) ? 1u : 0u;
}

// run_idx = exclusive scan of flag: the flagged position p starts run run_idx[p]
__global__ void __launch_bounds__(256)
k_run_starts(const unsigned long long* __restrict__ sorted, int64_t npairs, const uint32_t* __restrict__ run_idx,
             uint32_t* __restrict__ run_start, uint32_t* __restrict__ nruns_out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    const bool start = p == 0 || (uint32_t)(sorted[p - 1] >> 32) != (uint32_t)(sorted[p] >> 32);
    if (start) run_start[run_idx[p]] = (uint32_t)p;
    if (p == npairs - 1) {
        const uint32_t nruns = run_idx[p] + (start ? 1u : 0u);
        run_start[nruns] = (uint32_t)npairs;
        *nruns_out = nruns;
    }
}

__global__ void __launch_bounds__(256)
k_run_unit_counts(const uint32_t* __restrict__ run_start, const uint32_t* __restrict__ nruns_p, int64_t cap,
                  uint32_t* __restrict__ ucnt) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= cap) return;
    const uint32_t nruns = *nruns_p;
    ucnt[r] = r < nruns ? (run_start[r + 1] - run_start[r] + SS_UQ - 1) / SS_UQ : 0u;
}

// uoff = exclusive scan of ucnt.  One thread per sorted position; the positions that open a unit (every SS_UQ-th
// of a run) write its record.
__global__ void __launch_bounds__(256)
k_emit_units(const unsigned long long* __restrict__ sorted, int64_t npairs, const uint32_t* __restrict__ run_idx,
             const uint32_t* __restrict__ run_start, const uint32_t* __restrict__ uoff, const int32_t* __restrict__ pair_q,
             const uint32_t* __restrict__ pair_len, const uint32_t* __restrict__ pair_seg, const int32_t* __restrict__ ids_sorted,
             UnitRec* __restrict__ units) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    const unsigned long long k = sorted[p];
    const uint32_t bstart = (uint32_t)(k >> 32);
    const bool start = p == 0 || (uint32_t)(sorted[p - 1] >> 32) != bstart;
    const uint32_t r = run_idx[p] + (start ? 1u : 0u) - 1u;
    const uint32_t p0 = run_start[r], p1 = run_start[r + 1];
    if ((p - p0) % SS_UQ) return;
    UnitRec rec;
    rec.bstart = bstart;
    rec.len = pair_len[(uint32_t)k];
    rec.pos0 = (uint32_t)p;
    rec.m = min((uint32_t)SS_UQ, p1 - (uint32_t)p);
#pragma unroll
    for (int j = 0; j < SS_UQ; ++j) {
        const uint32_t pi = (uint32_t)sorted[min((uint32_t)p + j, p1 - 1)];
        rec.q[j] = pair_q[pi];
        rec.seg[j] = pair_seg[pi];
    }
#pragma unroll
    for (int j = 0; j < SS_WIN; ++j) rec.ids0[j] = ids_sorted[bstart + min((uint32_t)j, rec.len - 1)];
    units[uoff[r] + (uint32_t)(p - p0) / SS_UQ] = rec;
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
               const uint32_t* __restrict__ nunits_p, const int32_t* __restrict__ ids_sorted, double* __restrict__ scores,
               unsigned long long* __restrict__ stat /* [0] units, [1] rows staged */) {
    using SK = StoreKind<KIND>;
    constexpr int E = SK::E;
    constexpr int NJ = BM_KC * SK::SZ / 64;         // 64-byte groups (4 chunks) per row
    constexpr int NS = NJ * E;                      // DMMA k-steps per row = BM_KC / 4
    constexpr int SLOT_DOUBLES = SS_ROWS * SS_PITCH;
    extern __shared__ __align__(128) unsigned char ssm_raw[];
    __shared__ uint64_t bars[SS_WARPS][SS_STAGES + 4];   // ring slots, 2 unit records, 2 id windows
    __shared__ int4 meta[SS_WARPS][SS_STAGES];           // per slot: {kind: 0/1 query block, 2 rows; first row; rows; m}
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
        if (p_phase < 2) {             // 8 query rows (always FP64); each row's score segment rides in the row's padding
            const int nrows = min(SS_ROWS, (int)p_m - SS_ROWS * p_phase);
            if (lane == 0) {
                meta[warp][s] = make_int4(p_phase, 0, nrows, (int)p_m);
                mbar_expect_tx(&bar_slot[s], (unsigned)nrows * q_bytes);
            }
            __syncwarp();
            if (lane < nrows) {
                const int j = SS_ROWS * p_phase + lane;
                reinterpret_cast<uint32_t*>(slot + (size_t)lane * SS_PITCH + BM_KC)[0] = r->seg[j];
                bulk_g2s(slot + (size_t)lane * SS_PITCH, Q + (int64_t)r->q[j] * d, q_bytes, &bar_slot[s], pol_keep);
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
                meta[warp][s] = make_int4(2, p_row, nrows, (int)p_m);
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
    uint32_t c_seg[2][2] = {{0, 0}, {0, 0}};
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
        if (mt.x < 2) {
            // query block mt.x of a unit of mt.w queries -> B fragments (columns >= d are 0), score segments, norms
            const double* br = slot + (size_t)g * SS_PITCH;
            if (mt.x == 0) c_nb = (mt.w + 7) >> 3;
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) {
                if (nb == mt.x) {
#pragma unroll
                    for (int j = 0; j < NJ; ++j)
#pragma unroll
                        for (int e = 0; e < E; e += 2) {
                            const int col = 64 * j / SK::SZ + col_t + e;          // even
                            double2 v = make_double2(0.0, 0.0);
                            if (j < nj && col < d) v = *reinterpret_cast<const double2*>(br + col);
                            B[nb][j * E + e] = v.x;
                            B[nb][j * E + e + 1] = col + 1 < d ? v.y : 0.0;
                        }
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        c_seg[nb][e] = reinterpret_cast<const uint32_t*>(slot + (size_t)(2 * t + e) * SS_PITCH + BM_KC)[0];
                        c_ok[nb][e] = 8 * nb + 2 * t + e < mt.w;
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
                    if (ragged && j == nj - 1) {
#pragma unroll
                        for (int e = 0; e < E; ++e)
                            if (64 * j / SK::SZ + col_t + e >= d) a[e] = 0.0;
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
        // thread (g, t) holds (row mt.y + g, queries 8nb + 2t, 8nb + 2t + 1)
        if (g < mt.z) {
            const uint32_t row = (uint32_t)(mt.y + g);
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (nb < c_nb && c_ok[nb][e]) {
                        const double v = acc[nb][0][e] + acc[nb][1][e];
                        scores[(size_t)c_seg[nb][e] + row] = ANGULAR ? v / (c_qn[nb][e] * xnr) : v;
                    }
        }
        __syncwarp();                       // every lane has read the slot before it is refilled
        consumed++;
    }
    if (lane == 0) { atomicAdd(&stat[0], (unsigned long long)nmine); atomicAdd(&stat[1], rows_staged); }
}

// Warp-autonomous variant (d even): every warp is an independent worker that pulls groups of BM_QT sorted pairs
// from a counter.  A-operand fragments (bucket rows) go straight from global memory to registers with LDG.128 —
// 16 independent 128-bit loads per 8-row block, two blocks in flight per warp (register double buffering), no
// shared-memory staging of rows and no CTA barrier — using a k permutation (k-step 2w takes columns 8w+2t, k-step
// 2w+1 takes 8w+2t+1) so that each thread's two fragment elements are adjacent in memory; the B operand (up to
// 8*NB queries of the run) sits in the warp's private slice of shared memory with the same permutation.
constexpr int WQ_PITCH = BM_KC + 8;           // (pitch * 8) mod 128 == 64: conflict-free LDS.128 per quarter warp

template <int NB>
struct WarpCfg {
    static constexpr int WQ = 8 * NB;                                               // queries per pass
    static constexpr int WARPS = (NB == 1) ? 8 : (NB == 2 ? 8 : 6);                 // shared memory bound
    static constexpr size_t SMEM = (size_t)WARPS * WQ * WQ_PITCH * sizeof(double);
};

template <bool ANGULAR, int NB>
__global__ void __launch_bounds__(WarpCfg<NB>::WARPS * 32, 1)
k_score_warps(const double* __restrict__ X, int d, const double* __restrict__ Q,
              const unsigned long long* __restrict__ pair_key /* sorted by bucket */, int64_t npairs,
              const int32_t* __restrict__ pair_q, const uint32_t* __restrict__ pair_len,
              const uint32_t* __restrict__ pair_seg, const int32_t* __restrict__ ids_sorted, double* __restrict__ scores,
              int* __restrict__ next_group, unsigned long long* __restrict__ stat /* [0] runs, [1] rows staged */) {
    constexpr int WQ = WarpCfg<NB>::WQ;
    constexpr int NW = BM_KC / 8;
    extern __shared__ double wsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    double* Qw = wsm + warp * WQ * WQ_PITCH;
    for (int i = lane; i < WQ * WQ_PITCH; i += 32) Qw[i] = 0.0;     // k padding stays zero: only columns < d are written
    __syncwarp();
    const int nw8 = (d + 7) >> 3;                                    // 8-column windows in use
    const int64_t ngroups = (npairs + BM_QT - 1) / BM_QT;
    unsigned long long runs = 0, rows_staged = 0;
    for (;;) {
        long long grp = 0;
        if (lane == 0) grp = atomicAdd(next_group, 1);
        grp = __shfl_sync(0xffffffffu, grp, 0);
        if (grp >= ngroups) break;
        const int64_t p0 = grp * BM_QT;
        const int np = (int)min((int64_t)BM_QT, npairs - p0);
        const unsigned long long mykey = lane < np ? pair_key[p0 + lane] : ~0ULL;
        const uint32_t mybucket = (uint32_t)(mykey >> 32);
        int j0 = 0;
        while (j0 < np) {
            const uint32_t bstart = __shfl_sync(0xffffffffu, mybucket, j0);
            const uint32_t same = __ballot_sync(0xffffffffu, lane >= j0 && lane < np && mybucket == bstart);
            const int m = __popc(same);                              // sorted => the run is lanes j0 .. j0+m-1
            const uint32_t first_pair = (uint32_t)__shfl_sync(0xffffffffu, mykey, j0);
            const int blen = (int)__ldg(pair_len + first_pair);
            const int32_t* bids = ids_sorted + bstart;
            runs++;
            for (int c0 = j0; c0 < j0 + m; c0 += WQ) {
                const int mc = min(WQ, j0 + m - c0);
                const int nbu = (mc + 7) >> 3;                       // n-blocks in use this pass
                rows_staged += blen;
                __syncwarp();
                {   // stage this pass's queries (row r of Qw = pair c0 + r): lane r fetches its query index, then the
                    // rows are copied with all loads of 4 queries in flight at a time
                    const uint32_t mypi = (uint32_t)mykey;
                    const int myq = (lane >= c0 && lane < c0 + mc) ? __ldg(pair_q + mypi) : 0;
                    for (int r0 = 0; r0 < mc; r0 += 4) {
                        double2 v[4][BM_KC / 64];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int qidx = __shfl_sync(0xffffffffu, myq, min(c0 + r0 + u, c0 + mc - 1));
                            const double* qsrc = Q + (int64_t)qidx * d;
#pragma unroll
                            for (int i = 0; i < BM_KC / 64; ++i) {
                                const int cc = 2 * lane + 64 * i;
                                v[u][i] = (cc < d) ? __ldg(reinterpret_cast<const double2*>(qsrc + cc)) : make_double2(0.0, 0.0);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (r0 + u < mc) {
#pragma unroll
                                for (int i = 0; i < BM_KC / 64; ++i) {
                                    const int cc = 2 * lane + 64 * i;
                                    if (cc < d) *reinterpret_cast<double2*>(Qw + (r0 + u) * WQ_PITCH + cc) = v[u][i];
                                }
                            }
                    }
                }
                // score segments of the queries this thread's accumulators belong to (columns 8nb+2t, 8nb+2t+1)
                int64_t seg[NB][2];
                bool qok[NB][2];
#pragma unroll
                for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int qi = 8 * nb + 2 * t + e;
                        const uint32_t pi = (uint32_t)__shfl_sync(0xffffffffu, mykey, min(c0 + qi, np - 1));
                        qok[nb][e] = qi < mc;
                        seg[nb][e] = qok[nb][e] ? (int64_t)__ldg(pair_seg + pi) : 0;
                    }
                __syncwarp();
                double qn[NB][2];
                if (ANGULAR) {
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            double s = 0;
                            const double* qq = Qw + (8 * nb + 2 * t + e) * WQ_PITCH;
                            for (int cc = 0; cc < d; ++cc) s = fma(qq[cc], qq[cc], s);
                            qn[nb][e] = sqrt(s);
                        }
                }
                const double* bq = Qw + g * WQ_PITCH + 2 * t;

                // row ids: a 32-row window per coalesced load, the next window prefetched one window ahead
                int idwin = __ldg(bids + min(lane, blen - 1));
                int idwin_next = __ldg(bids + min(32 + lane, blen - 1));
                int win_base = 0;
                auto load_block = [&](double2 (&a)[NW], int rb) {
                    if (rb >= win_base + 32) {            // warp-uniform
                        idwin = idwin_next;
                        win_base += 32;
                        idwin_next = __ldg(bids + min(win_base + 32 + lane, blen - 1));
                    }
                    const int id = __shfl_sync(0xffffffffu, idwin, (rb - win_base) + g);
                    const double* xr = X + (int64_t)id * d + 2 * t;
#pragma unroll
                    for (int w = 0; w < NW; ++w)
                        if (w < nw8)
                            a[w] = (8 * w + 2 * t < d) ? __ldg(reinterpret_cast<const double2*>(xr + 8 * w)) : make_double2(0.0, 0.0);
                };
                auto compute_block = [&](const double2 (&a)[NW], int rb) {
                    double acc[NB][2];
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) acc[nb][0] = acc[nb][1] = 0.0;
                    double xn = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        if (w < nw8) {
#pragma unroll
                            for (int nb = 0; nb < NB; ++nb) {
                                if (nb < nbu) {
                                    const double2 b = *reinterpret_cast<const double2*>(bq + nb * 8 * WQ_PITCH + 8 * w);
                                    dmma884(acc[nb][0], acc[nb][1], a[w].x, b.x);
                                    dmma884(acc[nb][0], acc[nb][1], a[w].y, b.y);
                                }
                            }
                            if (ANGULAR) { xn = fma(a[w].x, a[w].x, xn); xn = fma(a[w].y, a[w].y, xn); }
                        }
                    }
                    double xnr = 1.0;
                    if (ANGULAR) {
                        xn += __shfl_xor_sync(0xffffffffu, xn, 1);
                        xn += __shfl_xor_sync(0xffffffffu, xn, 2);
                        xnr = sqrt(xn);
                    }
                    const int row = rb + g;
                    if (row < blen) {
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                            for (int e = 0; e < 2; ++e)
                                if (qok[nb][e]) scores[seg[nb][e] + row] = ANGULAR ? acc[nb][e] / (qn[nb][e] * xnr) : acc[nb][e];
                    }
                };

                double2 A0[NW], A1[NW];
                load_block(A0, 0);
                for (int rb = 0; rb < blen; rb += 16) {
                    const bool has1 = rb + 8 < blen;
                    if (has1) load_block(A1, rb + 8);
                    compute_block(A0, rb);
                    if (rb + 16 < blen) load_block(A0, rb + 16);
                    if (has1) compute_block(A1, rb + 8);
                }
            }
            j0 += m;
        }
    }
    if (lane == 0) { atomicAdd(&stat[0], runs); atomicAdd(&stat[1], rows_staged); }
}

template <bool ANGULAR, int NB>
static void launch_score_warps(dpf_index* h, const double* Qd, int64_t npairs, int metric, unsigned long long* bm_stat) {
    (void)metric;
    static bool attr = false;
    if (!attr) {
        DPF_CUDA(cudaFuncSetAttribute(k_score_warps<ANGULAR, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)WarpCfg<NB>::SMEM));
        attr = true;
    }
    int* next_group = h->counters.p + 18;
    DPF_CUDA(cudaMemsetAsync(next_group, 0, sizeof(int), h->stream));
    const int64_t groups = (npairs + BM_QT - 1) / BM_QT;
    const unsigned grid = (unsigned)std::min<int64_t>((groups + WarpCfg<NB>::WARPS - 1) / WarpCfg<NB>::WARPS, (int64_t)h->num_sms);
    k_score_warps<ANGULAR, NB><<<grid, WarpCfg<NB>::WARPS * 32, WarpCfg<NB>::SMEM, h->stream>>>(
        h->Xdev, h->cfg.d, Qd, h->bm_sorted, npairs, h->pair_q.p, h->pair_len.p, h->pair_seg.p, h->ids_sorted.p, h->scores.p,
        next_group, bm_stat);
}

// per-query selection from its score segments: one CTA per query, a warp walks whole pairs
__global__ void __launch_bounds__(RR_THREADS)
k_select_pairs(int64_t q0, int L, const uint32_t* __restrict__ pair_base, const unsigned long long* __restrict__ pair_key_unsorted,
               const uint32_t* __restrict__ pair_len, const uint32_t* __restrict__ pair_seg,
               const int32_t* __restrict__ ids_sorted, const double* __restrict__ scores, const int32_t* __restrict__ qids,
               int self_exclude, int K, int32_t* __restrict__ ids_out, double* __restrict__ score_out) {
    extern __shared__ double rsm[];
    double* lkeys = rsm;                                 // RR_WARPS x K
    int* lids = reinterpret_cast<int*>(lkeys + RR_WARPS * K);
    __shared__ int s_counts[RR_WARPS];
    const int64_t ql = blockIdx.x, q = q0 + ql;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* mykeys = lkeys + warp * K;
    int* myids = lids + warp * K;
    const uint32_t pbeg = pair_base[ql * L], pend = pair_base[(ql + 1) * L];
    const int qid = qids ? qids[q] : INT32_MIN;
    const bool excl = self_exclude && qids && qid >= -128 && qid <= 127;
    int count = 0;
    // a warp walks whole pairs; SEL_U x 32 entries of a segment are loaded before any is examined (the loop is
    // otherwise one dependent load per iteration), and the next pair's descriptor is fetched one pair ahead
    constexpr int SEL_U = 4;
    uint32_t p = pbeg + warp;
    uint32_t n_bstart = 0, n_seg = 0;
    int n_len = 0;
    if (p < pend) { n_bstart = (uint32_t)(pair_key_unsorted[p] >> 32); n_len = (int)pair_len[p]; n_seg = pair_seg[p]; }
    for (; p < pend; p += RR_WARPS) {
        const uint32_t bstart = n_bstart;
        const int len = n_len;
        const double* sc = scores + n_seg;
        const int32_t* ids = ids_sorted + bstart;
        if (p + RR_WARPS < pend) {
            n_bstart = (uint32_t)(pair_key_unsorted[p + RR_WARPS] >> 32);
            n_len = (int)pair_len[p + RR_WARPS];
            n_seg = pair_seg[p + RR_WARPS];
        }
        for (int j0 = 0; j0 < len; j0 += 32 * SEL_U) {
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
            const int j = j0 + lane;
            const double k0 = j < len ? sc[j] : nan, k1 = j + 32 < len ? sc[j + 32] : nan;
            const double k2 = j + 64 < len ? sc[j + 64] : nan, k3 = j + 96 < len ? sc[j + 96] : nan;
            const int i0 = j < len ? __ldg(ids + j) : -1, i1 = j + 32 < len ? __ldg(ids + j + 32) : -1;
            const int i2 = j + 64 < len ? __ldg(ids + j + 64) : -1, i3 = j + 96 < len ? __ldg(ids + j + 96) : -1;
            auto examine = [&](double key, int id) {
                bool cand = (key == key) && !(excl && id == qid);
                if (cand && count == K) cand = better(key, id, mykeys[K - 1], myids[K - 1]);
                uint32_t todo = __ballot_sync(0xffffffffu, cand);
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const double kk = __shfl_sync(0xffffffffu, key, src);
                    const int ii = __shfl_sync(0xffffffffu, id, src);
                    // the same id reached through another table carries a bit-identical score: keep it once
                    bool dup = false;
                    for (int base = 0; base < count; base += 32) {
                        const int i = base + lane;
                        dup |= __any_sync(0xffffffffu, i < count && myids[i] == ii);
                    }
                    if (!dup) warp_insert(mykeys, myids, count, K, kk, ii, lane);
                }
            };
            static_assert(SEL_U == 4, "examine() calls below are written out");
            examine(k0, i0);
            examine(k1, i1);
            examine(k2, i2);
            examine(k3, i3);
        }
    }
    if (lane == 0) s_counts[warp] = count;
    __syncthreads();
    if (warp == 0) {
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
                if (bi != last) break;                   // duplicate across warps: skip
            }
            if (lane == 0) {
                ids_out[q * K + r] = bl >= 0 ? bi : -1;
                score_out[q * K + r] = bl >= 0 ? bk : __longlong_as_double(0x7ff8000000000000LL);
            }
            if (bl >= 0) last = bi;
        }
    }
}

__global__ void k_copy_u32(const uint32_t* __restrict__ a, uint32_t* __restrict__ b, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = a[i];
}


// no bucket probed by any query of the chunk: every result row is padding (-1, NaN)
__global__ void k_topk_select_empty(int64_t q0, int64_t nqc, int K, int32_t* __restrict__ ids_out, double* __restrict__ score_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nqc * K) return;
    ids_out[q0 * K + i] = -1;
    score_out[q0 * K + i] = __longlong_as_double(0x7ff8000000000000LL);
}


// runs of the sorted pair list -> h->units (device), number of units in h->bm_counts[1]
// runs of the sorted pair list -> unit records in h->bm_units (device); the number of units is also left in
// h->bm_counts[1]
static void build_units(dpf_index* h, int64_t npairs) {
    cudaStream_t st = h->stream;
    const unsigned gp = (unsigned)((npairs + 255) / 256);
    h->bm_flag.reserve(npairs + 1);
    h->bm_run_start.reserve(npairs + 2);
    h->bm_ucnt.reserve(npairs + 2);
    h->bm_counts.reserve(4);
    k_run_flags<<<gp, 256, 0, st>>>(h->bm_sorted, npairs, h->bm_flag.p); DPF_LAUNCHED();
    exclusive_scan_u32(h, h->bm_flag.p, npairs);
    k_run_starts<<<gp, 256, 0, st>>>(h->bm_sorted, npairs, h->bm_flag.p, h->bm_run_start.p, h->bm_counts.p); DPF_LAUNCHED();
    k_run_unit_counts<<<(unsigned)((npairs + 1 + 255) / 256), 256, 0, st>>>(h->bm_run_start.p, h->bm_counts.p, npairs + 1,
                                                                            h->bm_ucnt.p); DPF_LAUNCHED();
    exclusive_scan_u32(h, h->bm_ucnt.p, npairs + 1);      // bm_ucnt[r] = first unit of run r; [npairs] = number of units
    DPF_CUDA(cudaMemcpyAsync(h->bm_counts.p + 1, h->bm_ucnt.p + npairs, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    uint32_t nunits = 0;
    DPF_CUDA(cudaMemcpyAsync(&nunits, h->bm_ucnt.p + npairs, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    DPF_CUDA(cudaStreamSynchronize(st));
    h->bm_units.reserve((size_t)std::max<uint32_t>(nunits, 1) * sizeof(UnitRec));
    k_emit_units<<<gp, 256, 0, st>>>(h->bm_sorted, npairs, h->bm_flag.p, h->bm_run_start.p, h->bm_ucnt.p, h->pair_q.p, h->pair_len.p,
                                     h->pair_seg.p, h->ids_sorted.p, reinterpret_cast<UnitRec*>(h->bm_units.p)); DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
}

void topk_bucket_major(dpf_index* h, const double* Qd, const QueryKeys& qk, int steps, int probe_mode, int64_t q0, int64_t q1,
                       int64_t entries_ub, int topk, int metric, int32_t* ids_out, double* score_out) {
    const ProbeCtx c = make_ctx(h, steps, probe_mode);
    cudaStream_t st = h->stream;
    const int64_t nqc = q1 - q0;
    const int L = c.L, d = h->cfg.d;
    if (nqc <= 0) return;
    DPF_REQUIRE(h->h_table_base[L] < (1LL << 32), DPF_ERR_INVALID, "bucket-major re-rank: more than 2^32 forest entries");
    DPF_REQUIRE(entries_ub < (1LL << 32), DPF_ERR_INVALID, "bucket-major re-rank: chunk too large");
    const char* ev = getenv("DPF_BM_KERNEL");
    const bool use_stream = !(ev && ev[0] == 'w') && (reinterpret_cast<uintptr_t>(Qd) & 15) == 0;   // =warps: register-gather kernel
    // pair offsets of this chunk = exclusive scan of the per-(query, table) bucket counts from the probe pass
    const int64_t nslots = nqc * L + 1;
    h->pair_base.reserve(nslots);
    const bool use_u8 = use_stream && score_u8_usable(h);      // byte store: register-gather kernel (rerank_u8.cu)
    {
        StageTimer tm(h, DPF_T_EXPAND);
        if (use_u8 && q0 == 0) prepare_queries_u8(h, Qd, qk.nq);
        k_copy_u32<<<(unsigned)((nslots + 255) / 256), 256, 0, st>>>(h->pair_cnt.p + q0 * L, h->pair_base.p, nslots - 1); DPF_LAUNCHED();
        DPF_CUDA(cudaMemsetAsync(h->pair_base.p + nslots - 1, 0, sizeof(uint32_t), st));
        exclusive_scan_u32(h, h->pair_base.p, nslots);
        uint32_t npairs32 = 0;
        DPF_CUDA(cudaMemcpyAsync(&npairs32, h->pair_base.p + nslots - 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        DPF_CUDA(cudaStreamSynchronize(st));
        const int64_t npairs = npairs32;
        if (npairs == 0) {
            // nothing probed: all rows padded
            h->ucnt.reserve(nqc + 1);
            h->unit_off.reserve(nqc + 2);
            DPF_CUDA(cudaMemsetAsync(h->unit_off.p, 0, (nqc + 2) * sizeof(int64_t), st));
            k_topk_select_empty<<<(unsigned)((nqc * topk + 255) / 256), 256, 0, st>>>(q0, nqc, topk, ids_out, score_out); DPF_LAUNCHED();
            DPF_CUDA(cudaGetLastError());
            return;
        }
        h->pair_key.reserve(npairs);
        h->pair_key_alt.reserve(npairs);
        h->pair_q.reserve(npairs);
        h->pair_len.reserve(npairs + 1);
        h->pair_seg.reserve(npairs + 1);
        h->scores.reserve((size_t)std::max<int64_t>(entries_ub, 1));
        const int64_t warps = nqc * L;
        k_probe_pairs<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(c, qk.keys, h->qpids.p, qk.ld, q0, nqc, h->pair_base.p,
                                                                    h->pair_key.p, h->pair_q.p, h->pair_len.p); DPF_LAUNCHED();
        DPF_CUDA(cudaGetLastError());
        k_copy_u32<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(h->pair_len.p, h->pair_seg.p, npairs); DPF_LAUNCHED();
        exclusive_scan_u32(h, h->pair_seg.p, npairs);
        // sort a copy of the keys by bucket start (bits 32..); the unsorted array stays for the selection pass
        DPF_CUDA(cudaMemcpyAsync(h->pair_key_alt.p, h->pair_key.p, npairs * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
        int ebits = 1;
        while ((1LL << ebits) < h->h_table_base[L]) ebits++;
        h->sk64a.reserve(npairs);
        unsigned long long *a = h->pair_key_alt.p, *b = h->sk64a.p;
        radix_sort_keys_u64(h, &a, &b, npairs, 32, 32 + ebits);
        h->bm_sorted = a;
        h->bm_npairs = npairs;
        if (use_stream) build_units(h, npairs);
    }
    {
        StageTimer tm(h, DPF_T_RERANK);
        const int64_t npairs = h->bm_npairs;
        unsigned long long* bm_stat = reinterpret_cast<unsigned long long*>(h->counters.p + 26);   // cleared by probe_count_all
        h->stats[DPF_STAT_BM_PAIRS] += npairs;
        const bool ang = metric == DPF_METRIC_ANGULAR;
        if (use_u8) {
            launch_score_u8(h, Qd, h->bm_units.p, h->bm_counts.p + 1, ang, bm_stat);
        } else if (use_stream) {
            const UnitRec* units = reinterpret_cast<const UnitRec*>(h->bm_units.p);
            // rows come from the compact store when the build found a narrower lossless type (store.cu)
            const int kind = h->Xc_kind;
            const unsigned char* rows = kind == DPF_STORE_KIND_F64 ? reinterpret_cast<const unsigned char*>(h->Xdev) : h->Xc.p;
            const unsigned row_bytes = kind == DPF_STORE_KIND_F64 ? (unsigned)d * 8u : (unsigned)h->Xc_row_bytes;
            auto launch = [&](auto kern) {
                DPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SS_SMEM));
                kern<<<h->num_sms, SS_WARPS * 32, SS_SMEM, st>>>(rows, row_bytes, d, Qd, units, h->bm_counts.p + 1, h->ids_sorted.p,
                                                                 h->scores.p, bm_stat);
            };
            if (kind == DPF_STORE_KIND_U8) { if (ang) launch(k_score_stream<true, DPF_STORE_KIND_U8>); else launch(k_score_stream<false, DPF_STORE_KIND_U8>); }
            else if (kind == DPF_STORE_KIND_F32) { if (ang) launch(k_score_stream<true, DPF_STORE_KIND_F32>); else launch(k_score_stream<false, DPF_STORE_KIND_F32>); }
            else { if (ang) launch(k_score_stream<true, DPF_STORE_KIND_F64>); else launch(k_score_stream<false, DPF_STORE_KIND_F64>); }
        } else {
            const char* nbv = getenv("DPF_BM_NB");
            const int nb = nbv ? atoi(nbv) : 2;
            if (nb <= 1) { if (ang) launch_score_warps<true, 1>(h, Qd, npairs, metric, bm_stat); else launch_score_warps<false, 1>(h, Qd, npairs, metric, bm_stat); }
            else if (nb == 2) { if (ang) launch_score_warps<true, 2>(h, Qd, npairs, metric, bm_stat); else launch_score_warps<false, 2>(h, Qd, npairs, metric, bm_stat); }
            else { if (ang) launch_score_warps<true, 4>(h, Qd, npairs, metric, bm_stat); else launch_score_warps<false, 4>(h, Qd, npairs, metric, bm_stat); }
        }
        DPF_LAUNCHED();
        DPF_CUDA(cudaGetLastError());
    }
    {
        StageTimer tm(h, DPF_T_SELECT);
        const size_t smem = (size_t)RR_WARPS * topk * (sizeof(double) + sizeof(int));
        k_select_pairs<<<(unsigned)nqc, RR_THREADS, smem, st>>>(q0, L, h->pair_base.p, h->pair_key.p, h->pair_len.p, h->pair_seg.p,
                                                                 h->ids_sorted.p, h->scores.p, qk.qids,
                                                                 h->cfg.self_exclude_small_ids, topk, ids_out, score_out); DPF_LAUNCHED();
        DPF_CUDA(cudaGetLastError());
    }
}

}  // namespace dpf
