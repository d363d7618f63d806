// bm_group.cu — bucket-major re-rank, bookkeeping: which (bucket, query) pairs a chunk of queries produces and how
// they are grouped by bucket, without the host reading a single size back.
//
// Replaces, for a batch, the per-query walk of RandomDrawTreeMap.getSimilarWithStepWiseFaster
// (src/main/java/mclab/mapdb/RandomDrawTreeMap.java:742-797) + the union over tables of QueryTask
// (src/main/scala/mclab/deploy/DensevectorRDFInit.scala:414-432) by
//   k_probe_leaves     warp per (query, table): the distinct leaf buckets its probe keys reach (dense leaf numbers from
//                      the build's leaf table), cached, and one atomic per bucket into a histogram over the leaves
//   k_scan_leaves      one single-pass (decoupled look-back) scan of that histogram, carrying two sums at once: pairs
//                      before a leaf and units (<= SS_UQ queries of one bucket) before it; leaves the totals on the device
//   k_fill_pairs       the cached leaves again: every pair takes a slot of its leaf's range (atomic decrement of the
//                      histogram, which is therefore all zero again afterwards — no clearing between batches)
//   k_emit_units       warp per leaf: the self-contained unit records the scoring kernels stream
// A counting sort over <= a few hundred thousand leaves instead of the 3-pass 64-bit radix sort + run detection of
// round 1 (20 launches, 3 host synchronisations): 4 launches, 0 synchronisations.  Every scratch array has a worst-case
// size the host knows: pairs <= queries x tables x (sub-indexes within `steps`) x 28 probe keys.
// The order of the queries inside a bucket's list depends on the order the atomics land; results do not (a query's
// scores do not depend on which other queries share its unit, and the final selection orders by (score, id)).
#include "rerank_units.cuh"

namespace dpf {

// A CTA takes 256 consecutive (query, table) pairs.  Phase 1, a thread per pair: does this rank own a sub-index the pair
// searches (on G GPUs 1 - 1/G of the pairs do not: they cost one thread a load, not a warp a launch)?  The others go to a
// list in shared memory.  Phase 2, a warp per listed pair: the pair's probe keys, one per lane.
__global__ void __launch_bounds__(256)
k_probe_leaves(ProbeCtx c, const int32_t* __restrict__ qkeys, const uint8_t* __restrict__ qpids, int64_t ld, int64_t q0,
               int64_t nqc, uint32_t* __restrict__ leaf_cnt, uint32_t* __restrict__ pair_cnt, uint32_t* __restrict__ cache,
               int cap, uint32_t* __restrict__ q_entries /* per query: bucket entries it visits on this rank */,
               unsigned long long* __restrict__ stat_nlz, unsigned long long* __restrict__ stat_entries) {
    __shared__ unsigned long long s_entries;
    __shared__ unsigned int s_nlz, s_n;
    __shared__ uint16_t s_list[256];
    if (threadIdx.x == 0) { s_entries = 0ULL; s_nlz = 0u; s_n = 0u; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t wid0 = (int64_t)blockIdx.x * 256;
    const int np = 1 << c.tp.pb;
    {
        const int64_t wid = wid0 + threadIdx.x;
        if (wid < nqc * c.L) {
            const int pid = qpids[(int64_t)(wid % c.L) * ld + q0 + wid / c.L];
            bool any = false;
            for (int sub = 0; sub < np; ++sub) any |= __popc(sub ^ pid) <= c.steps && c.own.has((int)(wid % c.L), sub);
            if (any) s_list[atomicAdd(&s_n, 1u)] = (uint16_t)threadIdx.x;
            else pair_cnt[wid] = 0u;
        }
    }
    __syncthreads();
    const unsigned n_listed = s_n;
    for (unsigned li = warp; li < n_listed; li += 8) {
        const int64_t wid = wid0 + s_list[li];
        const int64_t q = q0 + wid / c.L;
        const int t = (int)(wid % c.L);
        const uint32_t h = (uint32_t)qkeys[(int64_t)t * ld + q];
        const int pid = qpids[(int64_t)t * ld + q];
        const int seg = c.tp.seg_bits ? (int)(h >> c.tp.bucket_bits) : 0;
        const int nprobes = probe_count(h, c.probe_mode);
        int total = 0, nbuckets = 0;
        if (nprobes < 0) {
            if (lane == 0) atomicAdd(&s_nlz, 1u);
        } else {
            for (int sub = 0; sub < np; ++sub) {       // findStepWiseSubIndexIDs (RandomDrawTreeMap.java:613-621)
                if (__popc(sub ^ pid) > c.steps) continue;
                if (!c.own.has(t, sub)) continue;      // this GPU's sub-forest only
                bool leader;
                uint32_t leaf;
                int cnt;
                warp_lookup_leaf(c, t, sub, seg, h, nprobes, lane, leader, leaf, cnt);
                const uint32_t m = __ballot_sync(0xffffffffu, leader);
                if (leader) {
                    cache[wid * cap + nbuckets + __popc(m & ((1u << lane) - 1u))] = leaf;
                    atomicAdd(leaf_cnt + leaf, 1u);
                    total += cnt;
                }
                nbuckets += __popc(m);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        }
        if (lane == 0) {
            pair_cnt[wid] = (uint32_t)nbuckets;
            if (total > 0) {
                atomicAdd(&s_entries, (unsigned long long)total);
                atomicAdd(q_entries + wid / c.L, (uint32_t)total);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_entries) atomicAdd(stat_entries, s_entries);
        if (s_nlz) atomicAdd(stat_nlz, (unsigned long long)s_nlz);
    }
}

// exclusive scan of (pairs, units) over the leaves.  Status word of a tile: flag (2 bits: 1 aggregate, 2 inclusive
// prefix) | units (30 bits) | pairs (32 bits).
constexpr int LS_THREADS = 256, LS_ITEMS = 16, LS_TILE = LS_THREADS * LS_ITEMS;
constexpr unsigned long long LS_VALUE_MASK = (1ULL << 62) - 1ULL;

__global__ void __launch_bounds__(LS_THREADS)
k_scan_leaves(const uint32_t* __restrict__ leaf_cnt, int64_t nleaves, uint32_t uq /* queries per unit */,
              const int32_t* __restrict__ leaf_len, uint32_t rows_per_unit /* 0: a unit takes the whole bucket */,
              uint32_t* __restrict__ leaf_off, uint32_t* __restrict__ unit_off,
              unsigned int* __restrict__ tile_counter, volatile unsigned long long* __restrict__ status,
              uint32_t* __restrict__ totals /* [0] pairs, [1] units */, unsigned long long* __restrict__ pairs_total /* or null */) {
    __shared__ unsigned long long wsum[LS_THREADS / 32];
    __shared__ unsigned long long s_prefix;
    __shared__ uint32_t s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t base = (int64_t)tile * LS_TILE + (int64_t)threadIdx.x * LS_ITEMS;
    unsigned long long v[LS_ITEMS], s = 0;
#pragma unroll
    for (int i = 0; i < LS_ITEMS; ++i) {
        const uint32_t c = base + i < nleaves ? leaf_cnt[base + i] : 0u;
        uint32_t units = (c + uq - 1) / uq;
        if (rows_per_unit && c) units *= ((uint32_t)leaf_len[base + i] + rows_per_unit - 1) / rows_per_unit;
        v[i] = (unsigned long long)c | ((unsigned long long)units << 32);
        s += v[i];
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        unsigned long long ws = lane < LS_THREADS / 32 ? wsum[lane] : 0ULL, wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += y;
        }
        if (lane < LS_THREADS / 32) wsum[lane] = wi - ws;
        const unsigned long long total = __shfl_sync(0xffffffffu, wi, LS_THREADS / 32 - 1);
        unsigned long long prefix = 0;
        if (tile == 0) {
            if (lane == 0) status[0] = (2ULL << 62) | total;
        } else {
            if (lane == 0) status[tile] = (1ULL << 62) | total;
            long long p = (long long)tile - 1;                        // look back: lane l reads tile p - l
            for (;;) {
                const long long mine = p - lane;
                unsigned long long st = mine >= 0 ? 0ULL : (2ULL << 62);   // before tile 0: an empty prefix
                if (mine >= 0)
                    do { st = status[mine]; } while ((st >> 62) == 0ULL);
                const uint32_t has_prefix = __ballot_sync(0xffffffffu, (st >> 62) == 2ULL);
                const int first = has_prefix ? __ffs(has_prefix) - 1 : 32;   // nearest predecessor with a full prefix
                unsigned long long part = lane <= first ? (st & LS_VALUE_MASK) : 0ULL;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                prefix += part;
                if (has_prefix) break;
                p -= 32;
            }
            if (lane == 0) status[tile] = (2ULL << 62) | (prefix + total);
        }
        if (lane == 0) {
            s_prefix = prefix;
            if ((int64_t)(tile + 1) * LS_TILE >= nleaves) {              // last tile: totals, and the end of the last leaf
                const unsigned long long all = prefix + total;
                totals[0] = (uint32_t)all;
                totals[1] = (uint32_t)(all >> 32);
                leaf_off[nleaves] = (uint32_t)all;
                unit_off[nleaves] = (uint32_t)(all >> 32);
                if (pairs_total) atomicAdd(pairs_total, all & 0xffffffffULL);
            }
        }
    }
    __syncthreads();
    unsigned long long run = s_prefix + wsum[w] + inc - s;
#pragma unroll
    for (int i = 0; i < LS_ITEMS; ++i) {
        if (base + i < nleaves) {
            leaf_off[base + i] = (uint32_t)run;
            unit_off[base + i] = (uint32_t)(run >> 32);
        }
        run += v[i];
    }
}

__global__ void __launch_bounds__(256)
k_fill_pairs(const uint32_t* __restrict__ pair_cnt, const uint32_t* __restrict__ cache, int cap, int64_t nwarps,
             uint32_t* __restrict__ leaf_cnt, const uint32_t* __restrict__ leaf_off, int L, int32_t* __restrict__ pair_q) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= nwarps) return;
    const uint32_t nb = pair_cnt[wid];
    const int32_t ql = (int32_t)(wid / L);
    for (uint32_t i = lane; i < nb; i += 32) {
        const uint32_t leaf = cache[wid * cap + i];
        const uint32_t old = atomicSub(leaf_cnt + leaf, 1u);     // counts down to zero: the histogram cleans itself
        pair_q[leaf_off[leaf] + old - 1u] = ql;
    }
}

// warp per leaf: the tcgen05 kernel's units — one record per (<= TC_TQ pairs of the leaf's list) x (<= 128 of its rows)
__global__ void __launch_bounds__(256)
k_emit_tc_recs(const uint32_t* __restrict__ leaf_off, const uint32_t* __restrict__ unit_off, const uint32_t* __restrict__ leaf_pos,
               const int32_t* __restrict__ leaf_len, int64_t nleaves, const int32_t* __restrict__ pair_q,
               const int32_t* __restrict__ ids_sorted, TcRec* __restrict__ recs, const int* __restrict__ q8_bad,
               uint32_t cap /* records that fit */, DirtySet dirty) {
    if (*q8_bad != 0) return;                       // a batch that is not byte vectors is scored from the UnitRecs
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t leaf = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); leaf < nleaves; leaf += nwarps) {
        const uint32_t p0 = leaf_off[leaf], p1 = leaf_off[leaf + 1];
        if (p0 == p1) continue;
        const uint32_t bstart = leaf_pos[leaf], len = (uint32_t)leaf_len[leaf];
        uint32_t u = unit_off[leaf];
        TcRec* rec = recs + u;
        for (uint32_t p = p0; p < p1; p += TC_TQ) {
            const uint32_t m = min((uint32_t)TC_TQ, p1 - p);
            const int32_t qv = pair_q[p + min((uint32_t)lane, m - 1u)];
            for (uint32_t r0 = 0; r0 < len; r0 += 128, ++rec, ++u) {
                if (u >= cap) {                      // the record array is sized for any realistic batch, not for the worst
                    if ((uint32_t)lane < m) dirty.mark(qv);   // case: the queries of a unit that does not fit are answered exhaustively
                    continue;
                }
                const uint32_t nrows = min(128u, len - r0);
                if (lane == 0) *reinterpret_cast<uint4*>(rec) = make_uint4(bstart, nrows, m, r0);
                rec->q[lane] = qv;
#pragma unroll
                for (int i = 0; i < 4; ++i) rec->ids[lane + 32 * i] = ids_sorted[bstart + r0 + min((uint32_t)(lane + 32 * i), nrows - 1u)];
            }
        }
    }
}

// warp per leaf: one record per SS_UQ pairs of its list.  gate != 0: only when some query of the batch is not a byte
// vector (the records are then what k_score_u8d reads; a batch of byte vectors is scored from the descriptors above)
__global__ void __launch_bounds__(256)
k_emit_units(const uint32_t* __restrict__ leaf_off, const uint32_t* __restrict__ unit_off, const uint32_t* __restrict__ leaf_pos,
             const int32_t* __restrict__ leaf_len, int64_t nleaves, const int32_t* __restrict__ pair_q,
             const int32_t* __restrict__ ids_sorted, UnitRec* __restrict__ units, const int* __restrict__ q8_bad, int gate) {
    if (gate && *q8_bad == 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t leaf = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); leaf < nleaves; leaf += nwarps) {
        const uint32_t p0 = leaf_off[leaf], p1 = leaf_off[leaf + 1];
        if (p0 == p1) continue;
        const uint32_t bstart = leaf_pos[leaf];
        const int len = leaf_len[leaf];
        const int32_t id0 = ids_sorted[bstart + min(lane, len - 1)];
        UnitRec* rec = units + unit_off[leaf];
        for (uint32_t p = p0; p < p1; p += SS_UQ, ++rec) {
            const uint32_t m = min((uint32_t)SS_UQ, p1 - p);
            if (lane == 0) *reinterpret_cast<uint4*>(rec) = make_uint4(bstart, (uint32_t)len, p, m);
            if (lane < SS_UQ) rec->q[lane] = pair_q[p + min((uint32_t)lane, m - 1u)];
            rec->ids0[lane] = id0;
        }
    }
}

// ---- survivors: per-query list offsets from the counts the scoring kernel left, then the scatter ------------------
__global__ void __launch_bounds__(1024)
k_survivor_offsets(const uint32_t* __restrict__ cnt, int64_t n, uint32_t* __restrict__ base) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t b0 = 0; b0 < n; b0 += 1024) {
        const int64_t i = b0 + threadIdx.x;
        const uint32_t v = i < n ? cnt[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        if (w == 0) {
            uint32_t ws = wsum[lane], wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += y;
            }
            wsum[lane] = wi - ws;
        }
        __syncthreads();
        const uint32_t c = carry;
        if (i < n) base[i] = c + wsum[w] + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + wsum[w] + inc;
        __syncthreads();
    }
}

// survivors per query: one pass over the pool (the scoring kernels do not count — an atomic per survivor inside their
// loop cost them ~10 %); lanes of a warp that hold the same query add once
__global__ void __launch_bounds__(256)
k_count_survivors(Filter flt) {
    const uint32_t n = min(*flt.pool_cursor, flt.pool_cap);
    const int lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x - lane; i0 < n; i0 += stride) {   // warp-uniform trip count
        const uint32_t i = i0 + lane;
        const int q = i < n ? flt.pool[i].q : -1;
        const uint32_t peers = __match_any_sync(0xffffffffu, q);
        if (q >= 0 && lane == __ffs(peers) - 1) atomicAdd(flt.cnt + q, (uint32_t)__popc(peers));
    }
}

// survivor records (SurvivorSink) -> per-query lists: at = base[q] + (entries of q written so far); the row id is looked
// up here.  Survivors of one query come in bursts (a bucket near the query yields many), so the lanes of a warp that
// hold the same query reserve their slots with one atomic.
__global__ void __launch_bounds__(256)
k_scatter_survivors(Filter flt, const int32_t* __restrict__ ids_sorted) {
    const uint32_t n = min(*flt.pool_cursor, flt.pool_cap);
    const int lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x - lane; i0 < n; i0 += stride) {   // warp-uniform trip count
        const uint32_t i = i0 + lane;
        SurvRec r;
        r.q = -1;
        if (i < n) r = flt.pool[i];
        const uint32_t peers = __match_any_sync(0xffffffffu, r.q);
        if (r.q < 0) continue;
        const int leader = __ffs(peers) - 1;
        uint32_t first = 0;
        if (lane == leader) first = atomicAdd(flt.fill + r.q, (uint32_t)__popc(peers));
        first = __shfl_sync(peers, first, leader);
        const uint32_t at = flt.base[r.q] + first + __popc(peers & ((1u << lane) - 1u));
        flt.s_score[at] = r.score;
        flt.s_id[at] = __ldg(ids_sorted + r.pos);
    }
}

// ---- host side --------------------------------------------------------------------------------------------------------
// queries per chunk and cache slots per (query, table): pairs of a chunk <= nqc * L * cap <= kMaxPairs
constexpr int64_t kMaxPairs = 96LL << 20;

int64_t bm_chunk_queries(const dpf_index* h, int steps, int probe_mode, int* cap_out) {
    int nsub = 0;
    for (int sub = 0; sub < (1 << h->tp.pb); ++sub) nsub += __builtin_popcount((unsigned)sub) <= steps ? 1 : 0;
    const int cap = nsub * (probe_mode == DPF_PROBE_NONE ? 1 : 28);
    *cap_out = cap;
    return std::max<int64_t>(1, kMaxPairs / ((int64_t)h->cfg.L * cap));
}

// probe: the distinct leaves of every (query, table) pair into the probe cache + the leaf histogram.  The threshold
// samples only need this part, so the caller forks its second stream right after it.
void probe_leaves(dpf_index* h, const QueryKeys& qk, int steps, int probe_mode, int64_t q0, int64_t nqc, int cap, uint32_t* q_entries) {
    const ProbeCtx c = make_ctx(h, steps, probe_mode);
    cudaStream_t st = h->stream;
    const int64_t warps = nqc * c.L;
    const int64_t pairs_ub = warps * cap;
    h->pair_cnt.reserve((size_t)warps + 1);
    h->probe_cache.reserve((size_t)pairs_ub);
    h->pair_q.reserve((size_t)pairs_ub);
    const int64_t ntiles = (h->num_leaves + LS_TILE - 1) / LS_TILE;
    h->scan_scratch.reserve((size_t)(4 * ntiles + 16));
    int32_t* ctr = h->counters.p;
    StageTimer tm(h, DPF_T_PROBE_COUNT);
    k_probe_leaves<<<(unsigned)((warps + 255) / 256), 256, 0, st>>>(
        c, qk.keys, h->qpids.p, qk.ld, q0, nqc, h->leaf_cnt.p, h->pair_cnt.p, h->probe_cache.p, cap, q_entries,
        reinterpret_cast<unsigned long long*>(ctr + CTR_STAT_NLZ), reinterpret_cast<unsigned long long*>(ctr + CTR_ENTRIES)); DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
}

// counting sort of the cached (leaf, query) pairs by leaf: scan of the histogram, then every pair takes its slot
void group_pairs(dpf_index* h, int64_t nqc, int cap, bool use_tc) {
    cudaStream_t st = h->stream;
    const int L = h->cfg.L;
    const int64_t warps = nqc * L;
    const int64_t nleaves = h->num_leaves;
    const int64_t ntiles = (nleaves + LS_TILE - 1) / LS_TILE;
    int32_t* ctr = h->counters.p;
    StageTimer tm(h, DPF_T_EXPAND);
    DPF_CUDA(cudaMemsetAsync(h->scan_scratch.p, 0, (size_t)(4 * ntiles + 8) * sizeof(uint32_t), st));
    DPF_CUDA(cudaMemsetAsync(ctr + CTR_NPAIRS_TC, 0, 3 * sizeof(int32_t), st));
    if (ntiles > 0) {
        k_scan_leaves<<<(unsigned)ntiles, LS_THREADS, 0, st>>>(
            h->leaf_cnt.p, nleaves, (uint32_t)SS_UQ, h->leaf_len.p, 0u, h->leaf_off.p, h->leaf_unit_off.p,
            reinterpret_cast<unsigned int*>(ctr + CTR_SCAN_TILE),
            reinterpret_cast<unsigned long long*>(h->scan_scratch.p + 2), reinterpret_cast<uint32_t*>(ctr + CTR_NPAIRS),
            reinterpret_cast<unsigned long long*>(ctr + CTR_BM_PAIRS_TOTAL)); DPF_LAUNCHED();
    }
    if (ntiles > 0 && use_tc) {            // the same histogram once more, at the tcgen05 kernel's unit width
        uint32_t* status2 = h->scan_scratch.p + 2 * ntiles + 4;
        k_scan_leaves<<<(unsigned)ntiles, LS_THREADS, 0, st>>>(
            h->leaf_cnt.p, nleaves, (uint32_t)TC_TQ, h->leaf_len.p, 128u, h->leaf_off.p, h->leaf_unit_off_tc.p,
            reinterpret_cast<unsigned int*>(ctr + CTR_SCAN_TILE_TC), reinterpret_cast<unsigned long long*>(status2),
            reinterpret_cast<uint32_t*>(ctr + CTR_NPAIRS_TC), nullptr); DPF_LAUNCHED();
    }
    k_fill_pairs<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(h->pair_cnt.p, h->probe_cache.p, cap, warps, h->leaf_cnt.p,
                                                              h->leaf_off.p, L, h->pair_q.p); DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
}

// unit records of the grouped pairs: at most one partial unit per leaf plus one per SS_UQ pairs
void emit_units(dpf_index* h, bool only_if_fp64_queries) {
    StageTimer tm(h, DPF_T_EXPAND);
    const int64_t nleaves = h->num_leaves;
    if (nleaves > 0) {
        const unsigned grid = (unsigned)std::min<int64_t>((nleaves + 7) / 8, (int64_t)h->num_sms * 16);   // grid-stride: a gated launch costs ~nothing
        k_emit_units<<<grid, 256, 0, h->stream>>>(h->leaf_off.p, h->leaf_unit_off.p, h->leaf_pos.p, h->leaf_len.p,
                                                                          nleaves, h->pair_q.p, h->ids_sorted.p,
                                                                          reinterpret_cast<UnitRec*>(h->bm_units.p),
                                                                          h->counters.p + CTR_Q8_BAD, only_if_fp64_queries ? 1 : 0); DPF_LAUNCHED();
    }
    DPF_CUDA(cudaGetLastError());
}

void emit_tc_recs(dpf_index* h, int64_t cap, const DirtySet& dirty) {
    StageTimer tm(h, DPF_T_EXPAND);
    const int64_t nleaves = h->num_leaves;
    if (nleaves > 0) {
        const unsigned grid = (unsigned)std::min<int64_t>((nleaves + 7) / 8, (int64_t)h->num_sms * 16);
        k_emit_tc_recs<<<grid, 256, 0, h->stream>>>(h->leaf_off.p, h->leaf_unit_off_tc.p, h->leaf_pos.p, h->leaf_len.p, nleaves,
                                                    h->pair_q.p, h->ids_sorted.p, reinterpret_cast<TcRec*>(h->bm_descs.p),
                                                    h->counters.p + CTR_Q8_BAD, (uint32_t)cap, dirty); DPF_LAUNCHED();
    }
    DPF_CUDA(cudaGetLastError());
}

void survivor_lists(dpf_index* h, const Filter& flt, int64_t nqc) {
    cudaStream_t st = h->stream;
    k_count_survivors<<<h->num_sms * 8, 256, 0, st>>>(flt); DPF_LAUNCHED();
    k_survivor_offsets<<<1, 1024, 0, st>>>(flt.cnt, nqc, flt.base); DPF_LAUNCHED();
    k_scatter_survivors<<<h->num_sms * 8, 256, 0, st>>>(flt, h->ids_sorted.p); DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
}

}  // namespace dpf
