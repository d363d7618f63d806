// query.cu — K4 (tree descent, multi-step sub-index search, multi-probe, candidate de-dup) and K5 (candidate
// gather + FP64 re-rank + top-k), plus the multi-GPU top-k merge.
//
// Replaces, per query batch (reference file:line relative to /root/reference):
//   RandomDrawTreeMap.getSimilarWithStepWiseFaster / getSimilarWithStepWise   src/main/java/mclab/mapdb/RandomDrawTreeMap.java:630-797
//   findStepWiseSubIndexIDs :613-621, getInnerWithSimilarity :1106-1121, searchWithSimilarity :940-994
//   DensevectorRDFInit.QueryTask (union over tables)                           src/main/scala/mclab/deploy/DensevectorRDFInit.scala:414-432
//   topKAndPrecisionScore's gather + dgemv + argsort                          DensevectorRDFInit.scala:472-507
//
// K4: one warp per (query, table); lane i evaluates probe key h ^ (1 << i) (the reference's probe list has at
// most 28 entries, so a warp covers it exactly); the warp walks the flat W-way nodes, removes duplicate buckets
// with match.any, then streams each distinct bucket's ids (coalesced) through a per-query bitmap (shared memory
// when the index fits, else a per-CTA slice in HBM) so that every id is emitted once.
// K5: one warp per candidate row: 128-bit coalesced loads of the FP64 row, fused dot / cosine / squared-L2
// against the query held in shared memory, warp-shuffle reduction, per-warp top-k lists merged per query.
#include <cstdlib>

#include "common.cuh"

namespace dpf {

// ---------------------------------------------------------------------------------------------------------
// descent
// ---------------------------------------------------------------------------------------------------------
struct ProbeCtx {
    ForestView f;
    TreeParams tp;
    int L, steps, probe_mode, rank, world, self_exclude;
};

// bucket lookup for one probe key (RandomDrawTreeMap.java:940-994): empty slot -> nothing; leaf -> (ptr,cnt);
// directory -> descend; falling off level 0 -> nothing
__device__ __forceinline__ bool descend(const ForestView& f, const TreeParams& tp, int root_node, uint32_t probe,
                                        int& ptr, int& cnt) {
    int node = root_node;
    for (int level = tp.MAXL; level >= 0; --level) {
        const int slot = (int)((probe >> (tp.nb * level)) & (uint32_t)(tp.W - 1));
        const int64_t idx = (int64_t)node * tp.W + slot;
        const int c = __ldg(f.child_cnt + idx);
        const int p = __ldg(f.child_ptr + idx);
        if (c == 0) return false;
        if (c > 0) { ptr = p; cnt = c; return true; }
        node = p;
    }
    return false;
}

// Probe list of one (query, table): dense = { h ^ (1<<i) : 0 <= i < 28 - nlz(h) } — h itself is NOT in the list
// and its length depends on nlz(h) (RandomDrawTreeMap.java:753-756, quirk Q4); none = { h }.
// Returns the number of probes, or -1 for the reference's NegativeArraySizeException case.
__device__ __forceinline__ int probe_count(uint32_t h, int probe_mode) {
    if (probe_mode == DPF_PROBE_NONE) return 1;
    return 32 - __clz((int)h) - 4;
}

// For one sub-index: every lane looks up its probe, duplicates are folded; on return `leader` marks the lanes
// that hold a distinct, non-empty bucket.
__device__ __forceinline__ void warp_lookup(const ProbeCtx& c, int t, int sub, int seg, uint32_t h, int nprobes,
                                            int lane, bool& leader, int& ptr, int& cnt) {
    ptr = 0;
    cnt = 0;
    bool ok = false;
    if (lane < nprobes) {
        const uint32_t probe = (c.probe_mode == DPF_PROBE_NONE) ? h : (h ^ (1u << lane));
        ok = descend(c.f, c.tp, t * c.tp.R + sub * c.tp.SEG + seg, probe, ptr, cnt);
    }
    const int key = ok ? ptr : (-1 - lane);
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    leader = ok && ((__ffs(peers) - 1) == lane);
}

// pass A: upper bound of the candidate count per query (sum of distinct bucket sizes over tables)
__global__ void __launch_bounds__(256)
k_probe_count(ProbeCtx c, const int32_t* __restrict__ qkeys, const uint8_t* __restrict__ qpids, int64_t ld, int64_t nq,
              int32_t* __restrict__ q_ub, unsigned long long* __restrict__ stat /* [0] nlz>28, [1] with-dups */,
              uint32_t* __restrict__ pair_cnt /* nq x L distinct buckets per (query, table), may be null */) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= nq * c.L) return;
    const int64_t q = wid / c.L;
    const int t = (int)(wid % c.L);
    const uint32_t h = (uint32_t)qkeys[(int64_t)t * ld + q];
    const int pid = qpids[(int64_t)t * ld + q];
    const int seg = c.tp.seg_bits ? (int)(h >> c.tp.bucket_bits) : 0;
    const int nprobes = probe_count(h, c.probe_mode);
    if (nprobes < 0) {
        if (lane == 0) atomicAdd(&stat[0], 1ULL);
        return;
    }
    int total = 0, nbuckets = 0;
    const int np = 1 << c.tp.pb;
    for (int sub = 0; sub < np; ++sub) {       // findStepWiseSubIndexIDs (RandomDrawTreeMap.java:613-621)
        if (__popc(sub ^ pid) > c.steps) continue;
        if (c.world > 1 && (sub % c.world) != c.rank) continue;   // this GPU's sub-forest only
        bool leader;
        int ptr, cnt;
        warp_lookup(c, t, sub, seg, h, nprobes, lane, leader, ptr, cnt);
        total += leader ? cnt : 0;
        nbuckets += __popc(__ballot_sync(0xffffffffu, leader));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if (lane == 0 && pair_cnt) pair_cnt[wid] = (uint32_t)nbuckets;
    if (lane == 0 && total > 0) {
        atomicAdd(&q_ub[q], total);
        atomicAdd(&stat[1], (unsigned long long)total);
    }
}

// pass B: expansion + de-dup.  Persistent CTAs pull queries from a counter; bitmap in shared memory (SMEM_BM)
// or in this CTA's slice of a global scratch.
constexpr int EXP_THREADS = 512;
template <bool SMEM_BM>
__global__ void __launch_bounds__(EXP_THREADS)
k_expand(ProbeCtx c, const int32_t* __restrict__ qkeys, const uint8_t* __restrict__ qpids, int64_t ld, int64_t q0,
         int64_t q1, int64_t base, const int32_t* __restrict__ qids, const int64_t* __restrict__ q_off, int32_t* __restrict__ q_cnt,
         int32_t* __restrict__ cand, uint32_t* __restrict__ gbitmap, int64_t bm_words, int* __restrict__ next_query,
         unsigned long long* __restrict__ stat_unique) {
    extern __shared__ uint32_t sbm[];
    __shared__ int s_q, s_count;
    uint32_t* bm = SMEM_BM ? sbm : gbitmap + (int64_t)blockIdx.x * bm_words;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = EXP_THREADS / 32;
    if (SMEM_BM) {
        for (int64_t i = tid; i < bm_words; i += EXP_THREADS) bm[i] = 0u;
    }
    __syncthreads();
    const int np = 1 << c.tp.pb;
    for (;;) {
        if (tid == 0) { s_q = atomicAdd(next_query, 1); s_count = 0; }
        __syncthreads();
        const int64_t q = q0 + s_q;
        if (q >= q1) break;
        const int64_t off = q_off[q] - base;
        const int qid = qids ? qids[q] : INT32_MIN;
        // quirk Q3 (RandomDrawTreeMap.java:982): `ln.key != key` compares boxed Integers by reference, so the
        // query's own id is dropped only inside the Integer cache
        const bool excl = c.self_exclude && qids && qid >= -128 && qid <= 127;
        for (int t = warp; t < c.L; t += nwarps) {
            const uint32_t h = (uint32_t)qkeys[(int64_t)t * ld + q];
            const int pid = qpids[(int64_t)t * ld + q];
            const int seg = c.tp.seg_bits ? (int)(h >> c.tp.bucket_bits) : 0;
            const int nprobes = probe_count(h, c.probe_mode);
            if (nprobes < 0) continue;
            const int64_t tbase = c.f.table_base[t];
            for (int sub = 0; sub < np; ++sub) {
                if (__popc(sub ^ pid) > c.steps) continue;
                if (c.world > 1 && (sub % c.world) != c.rank) continue;
                bool leader;
                int ptr, cnt;
                warp_lookup(c, t, sub, seg, h, nprobes, lane, leader, ptr, cnt);
                uint32_t todo = __ballot_sync(0xffffffffu, leader);
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int bptr = __shfl_sync(0xffffffffu, ptr, src);
                    const int bcnt = __shfl_sync(0xffffffffu, cnt, src);
                    const int32_t* ids = c.f.ids_sorted + tbase + bptr;
                    for (int j0 = 0; j0 < bcnt; j0 += 32) {
                        const int j = j0 + lane;
                        bool fresh = false;
                        int id = 0;
                        if (j < bcnt) {
                            id = __ldg(ids + j);
                            if (!(excl && id == qid)) {
                                const uint32_t bit = 1u << (id & 31);
                                const uint32_t old = atomicOr(&bm[id >> 5], bit);
                                fresh = !(old & bit);
                            }
                        }
                        const uint32_t m = __ballot_sync(0xffffffffu, fresh);
                        if (m) {
                            int base = 0;
                            if (lane == 0) base = atomicAdd(&s_count, __popc(m));
                            base = __shfl_sync(0xffffffffu, base, 0);
                            if (fresh) cand[off + base + __popc(m & ((1u << lane) - 1u))] = id;
                        }
                    }
                }
            }
        }
        __syncthreads();
        const int count = s_count;
        if (tid == 0) { q_cnt[q] = count; atomicAdd(stat_unique, (unsigned long long)count); }
        for (int i = tid; i < count; i += EXP_THREADS) bm[cand[off + i] >> 5] = 0u;   // reset only what was touched
        __syncthreads();
    }
}

__global__ void k_gather_query_keys(const int32_t* __restrict__ keys, const uint8_t* __restrict__ pids, int64_t ld,
                                    const int32_t* __restrict__ qids, int64_t nq, int L, int64_t n,
                                    int32_t* __restrict__ qkeys, uint8_t* __restrict__ qpids, int* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * L) return;
    const int t = (int)(i / nq);
    const int64_t q = i % nq;
    const int id = qids[q];
    if (id < 0 || id >= n) { *bad = 1; qkeys[i] = 0; qpids[i] = 0; return; }
    qkeys[i] = keys[(int64_t)t * ld + id];
    qpids[i] = pids[(int64_t)t * ld + id];
}

void gather_query_keys(dpf_index* h, const int32_t* qids_dev, int64_t nq) {
    h->counters.reserve(64);
    DPF_CUDA(cudaMemsetAsync(h->counters.p + 16, 0, sizeof(int32_t), h->stream));
    const int64_t tot = nq * h->cfg.L;
    k_gather_query_keys<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(
        h->keys.p, h->pids.p, h->key_ld, qids_dev, nq, h->cfg.L, h->n, h->qkeys.p, h->qpids.p, h->counters.p + 16); DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
    int32_t bad = 0;
    DPF_CUDA(cudaMemcpyAsync(&bad, h->counters.p + 16, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    DPF_CUDA(cudaStreamSynchronize(h->stream));
    // the reference prints "fetch vector ... but got NULL" and calls System.exit(1) (RandomDrawTreeMap.java:1508-1511)
    DPF_REQUIRE(!bad, DPF_ERR_INVALID, "query id is not in the index");
}

static ProbeCtx make_ctx(dpf_index* h, int steps, int probe_mode) {
    ProbeCtx c;
    c.f = forest_view(h);
    c.tp = h->tp;
    c.L = h->cfg.L;
    c.steps = steps;
    c.probe_mode = probe_mode;
    c.world = h->cfg.world > 1 ? h->cfg.world : 1;
    c.rank = c.world > 1 ? h->cfg.rank : 0;
    c.self_exclude = h->cfg.self_exclude_small_ids;
    return c;
}

// pass A over the whole batch: h->q_cnt = per-query upper bound, h->q_off = its exclusive scan (device) and
// off_host = the same offsets on the host, from which the caller cuts the batch into memory-bounded chunks
void probe_count_all(dpf_index* h, const QueryKeys& qk, int steps, int probe_mode, std::vector<int64_t>& off_host) {
    const int64_t nq = qk.nq;
    const ProbeCtx c = make_ctx(h, steps, probe_mode);
    cudaStream_t st = h->stream;
    h->q_cnt.reserve(nq + 1);
    h->q_off.reserve(nq + 1);
    h->counters.reserve(64);
    unsigned long long* stat = reinterpret_cast<unsigned long long*>(h->counters.p + 20);
    DPF_CUDA(cudaMemsetAsync(h->counters.p + 16, 0, 16 * sizeof(int32_t), st));
    h->stats[DPF_STAT_BM_PAIRS] = 0;
    DPF_CUDA(cudaMemsetAsync(h->q_cnt.p, 0, (nq + 1) * sizeof(int32_t), st));
    h->pair_cnt.reserve((size_t)nq * c.L + 1);
    DPF_CUDA(cudaMemsetAsync(h->pair_cnt.p, 0, ((size_t)nq * c.L + 1) * sizeof(uint32_t), st));
    {
        StageTimer tm(h, DPF_T_PROBE_COUNT);
        const int64_t warps = nq * c.L;
        k_probe_count<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(c, qk.keys, h->qpids.p, qk.ld, nq, h->q_cnt.p, stat,
                                                                    h->pair_cnt.p); DPF_LAUNCHED();
        DPF_CUDA(cudaGetLastError());
        exclusive_scan_i64(h, h->q_cnt.p, h->q_off.p, nq);
    }
    off_host.resize((size_t)nq + 1);
    unsigned long long hstat[2];
    DPF_CUDA(cudaMemcpyAsync(off_host.data(), h->q_off.p, (nq + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    DPF_CUDA(cudaMemcpyAsync(hstat, stat, sizeof(hstat), cudaMemcpyDeviceToHost, st));
    DPF_CUDA(cudaStreamSynchronize(st));
    h->stats[DPF_STAT_NLZ_GT28] = (int64_t)hstat[0];
    h->stats[DPF_STAT_LAST_CAND_WITH_DUPS] = (int64_t)hstat[1];
}

// end of the chunk starting at q0 whose candidate upper bound fits `budget` ids (always at least one query)
int64_t next_chunk_end(const std::vector<int64_t>& off, int64_t q0, int64_t budget) {
    const int64_t nq = (int64_t)off.size() - 1;
    int64_t q1 = q0 + 1;
    while (q1 < nq && off[(size_t)q1 + 1] - off[(size_t)q0] <= budget) q1++;
    return q1;
}

// pass B for queries [q0, q1): unique (unsorted) candidates of query q at h->cand[q_off[q] - base ...], count in q_cnt[q]
void expand_range(dpf_index* h, const QueryKeys& qk, int steps, int probe_mode, int64_t q0, int64_t q1, int64_t base,
                  int64_t ub_chunk) {
    const ProbeCtx c = make_ctx(h, steps, probe_mode);
    cudaStream_t st = h->stream;
    const int64_t nqc = q1 - q0;
    unsigned long long* stat = reinterpret_cast<unsigned long long*>(h->counters.p + 20);
    h->cand.reserve((size_t)std::max<int64_t>(ub_chunk, 1));
    StageTimer tm(h, DPF_T_EXPAND);
    const int64_t bm_words = (h->n + 31) / 32 + 1;
    const size_t smem_need = (size_t)bm_words * sizeof(uint32_t);
    int* next_query = h->counters.p + 16;
    DPF_CUDA(cudaMemsetAsync(next_query, 0, sizeof(int), st));
    const bool use_smem = smem_need <= 200 * 1024;
    if (use_smem) {
        static size_t attr = 0;
        if (smem_need > attr) {
            DPF_CUDA(cudaFuncSetAttribute(k_expand<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_need));
            attr = smem_need;
        }
        int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / std::max<size_t>(smem_need, 1)));
        const int grid = (int)std::min<int64_t>(nqc, (int64_t)h->num_sms * per_sm);
        k_expand<true><<<grid, EXP_THREADS, smem_need, st>>>(c, qk.keys, h->qpids.p, qk.ld, q0, q1, base, qk.qids, h->q_off.p,
                                                              h->q_cnt.p, h->cand.p, nullptr, bm_words, next_query, stat + 2); DPF_LAUNCHED();
    } else {
        const int grid = (int)std::min<int64_t>(nqc, (int64_t)h->num_sms * 2);
        const size_t need = (size_t)h->num_sms * 2 * bm_words;
        if (h->bitmap.cap < need) {
            h->bitmap.reserve(need);
            DPF_CUDA(cudaMemsetAsync(h->bitmap.p, 0, need * sizeof(uint32_t), st));   // kept all-zero between calls
        }
        k_expand<false><<<grid, EXP_THREADS, 0, st>>>(c, qk.keys, h->qpids.p, qk.ld, q0, q1, base, qk.qids, h->q_off.p,
                                                      h->q_cnt.p, h->cand.p, h->bitmap.p, bm_words, next_query, stat + 2); DPF_LAUNCHED();
    }
    DPF_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------
// sorted unique CSR output
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_compose_keys(const int64_t* __restrict__ q_off, int64_t base, const int64_t* __restrict__ out_off,
               const int32_t* __restrict__ q_cnt, const int32_t* __restrict__ cand, unsigned long long* __restrict__ keys) {
    const int64_t q = blockIdx.x;
    const int cnt = q_cnt[q];
    const int64_t src = q_off[q] - base, dst = out_off[q];
    for (int i = threadIdx.x; i < cnt; i += blockDim.x)
        keys[dst + i] = ((unsigned long long)q << 32) | (uint32_t)cand[src + i];
}
__global__ void k_low32(const unsigned long long* __restrict__ keys, int64_t n, int32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int32_t)(uint32_t)keys[i];
}

// sorted unique ids of queries [q0, q1), concatenated, left at h->cand[0 .. total); off_dev (q1-q0+1 entries) gets
// the chunk-local CSR offsets
int64_t finalize_candidates_sorted(dpf_index* h, int64_t q0, int64_t q1, int64_t base, int64_t* off_dev) {
    StageTimer tm(h, DPF_T_CAND_SORT);
    cudaStream_t st = h->stream;
    const int64_t nqc = q1 - q0;
    exclusive_scan_i64(h, h->q_cnt.p + q0, off_dev, nqc);
    int64_t total = 0;
    DPF_CUDA(cudaMemcpyAsync(&total, off_dev + nqc, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    DPF_CUDA(cudaStreamSynchronize(st));
    if (total == 0) return 0;
    h->sk64a.reserve(total);
    h->sk64b.reserve(total);
    k_compose_keys<<<(unsigned)nqc, 256, 0, st>>>(h->q_off.p + q0, base, off_dev, h->q_cnt.p + q0, h->cand.p, h->sk64a.p); DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
    int idbits = 1, qbits = 1;
    while ((1LL << idbits) < h->n) idbits++;
    while ((1LL << qbits) < nqc) qbits++;
    unsigned long long *a = h->sk64a.p, *b = h->sk64b.p;
    radix_sort_keys_u64(h, &a, &b, total, 0, idbits);
    radix_sort_keys_u64(h, &a, &b, total, 32, 32 + qbits);
    k_low32<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, total, h->cand.p); DPF_LAUNCHED();   // total <= upper bound: fits
    DPF_CUDA(cudaGetLastError());
    return total;
}

// ---------------------------------------------------------------------------------------------------------
// K5: gather + re-rank + top-k
// ---------------------------------------------------------------------------------------------------------
// Work decomposition: a query's candidate list is cut into units of `seg` candidates; persistent CTAs pull units
// from a counter, so the grid stays full whether the batch is many light queries or a few heavy ones.  Each unit
// produces a sorted partial top-k; k_topk_select merges a query's partial lists.
constexpr int RR_THREADS = 256;
constexpr int RR_WARPS = RR_THREADS / 32;
constexpr int RR_MAXK = 256;
constexpr int RR_ROWS = 4;          // candidate rows in flight per warp
constexpr int RR_MIN_SEG = 2048;    // candidates per unit (lower bound)
constexpr int RR_MAX_UNITS = 128;   // units per query (upper bound; = threads of k_topk_select)

// total order of results: larger key first, ties by smaller id (key = score, or -distance for L2)
__device__ __forceinline__ bool better(double ka, int ia, double kb, int ib) {
    return ka > kb || (ka == kb && ia < ib);
}

// warp-cooperative insertion into a descending list of length <= K held in shared memory
__device__ __forceinline__ void warp_insert(double* keys, int* ids, int& count, int K, double key, int id, int lane) {
    if (count == K && !better(key, id, keys[K - 1], ids[K - 1])) return;
    int pos = 0;   // number of entries that are better than the new one
    for (int base = 0; base < count; base += 32) {
        const int i = base + lane;
        const bool b = i < count && better(keys[i], ids[i], key, id);
        pos += __popc(__ballot_sync(0xffffffffu, b));
    }
    const int newcount = min(count + 1, K);
    for (int hi = newcount - 1; hi > pos; hi -= 32) {   // shift [pos, newcount-1) right by one, from the back
        const int i = hi - lane;
        double kv = 0;
        int iv = 0;
        const bool mv = i > pos;
        if (mv) { kv = keys[i - 1]; iv = ids[i - 1]; }
        __syncwarp();
        if (mv) { keys[i] = kv; ids[i] = iv; }
        __syncwarp();
    }
    if (lane == 0) { keys[pos] = key; ids[pos] = id; }
    __syncwarp();
    count = newcount;
}

__global__ void k_unit_counts(const int32_t* __restrict__ cnt, int64_t q0, int64_t nqc, int seg, int32_t* __restrict__ ucnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nqc) ucnt[i] = (cnt[q0 + i] + seg - 1) / seg;
}

template <bool VEC2, int METRIC>
__device__ __forceinline__ void score_rows(const double* __restrict__ X, int d, const double* qs, const int (&id)[RR_ROWS],
                                           int lane, double (&s)[RR_ROWS], double (&xn)[RR_ROWS]) {
#pragma unroll
    for (int r = 0; r < RR_ROWS; ++r) { s[r] = 0.0; xn[r] = 0.0; }
    if (VEC2) {
        const double2* q2 = reinterpret_cast<const double2*>(qs);
        for (int j = lane; j < (d >> 1); j += 32) {
            double2 x[RR_ROWS];
#pragma unroll
            for (int r = 0; r < RR_ROWS; ++r)    // all loads of the iteration are issued before any use
                x[r] = __ldg(reinterpret_cast<const double2*>(X + (int64_t)id[r] * d) + j);
            const double2 qq = q2[j];
#pragma unroll
            for (int r = 0; r < RR_ROWS; ++r) {
                if (METRIC == DPF_METRIC_L2) {
                    const double a = qq.x - x[r].x, b = qq.y - x[r].y;
                    s[r] = fma(a, a, s[r]);
                    s[r] = fma(b, b, s[r]);
                } else {
                    s[r] = fma(x[r].x, qq.x, s[r]);
                    s[r] = fma(x[r].y, qq.y, s[r]);
                    if (METRIC == DPF_METRIC_ANGULAR) { xn[r] = fma(x[r].x, x[r].x, xn[r]); xn[r] = fma(x[r].y, x[r].y, xn[r]); }
                }
            }
        }
    } else {
        for (int j = lane; j < d; j += 32) {
            double x[RR_ROWS];
#pragma unroll
            for (int r = 0; r < RR_ROWS; ++r) x[r] = __ldg(X + (int64_t)id[r] * d + j);
            const double qq = qs[j];
#pragma unroll
            for (int r = 0; r < RR_ROWS; ++r) {
                if (METRIC == DPF_METRIC_L2) { const double a = qq - x[r]; s[r] = fma(a, a, s[r]); }
                else {
                    s[r] = fma(x[r], qq, s[r]);
                    if (METRIC == DPF_METRIC_ANGULAR) xn[r] = fma(x[r], x[r], xn[r]);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int r = 0; r < RR_ROWS; ++r) {
            s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
            if (METRIC == DPF_METRIC_ANGULAR) xn[r] += __shfl_xor_sync(0xffffffffu, xn[r], o);
        }
}

template <bool VEC2, int METRIC>
__global__ void __launch_bounds__(RR_THREADS)
k_rerank_units(const double* __restrict__ X, int d, const double* __restrict__ Q, int64_t q0, int64_t nqc, int64_t base,
               const int64_t* __restrict__ off, const int32_t* __restrict__ cnt, const int32_t* __restrict__ cand,
               const int64_t* __restrict__ unit_off, int seg, int K, double* __restrict__ part_key,
               int32_t* __restrict__ part_id, int* __restrict__ next_unit) {
    extern __shared__ double rsm[];
    double* qs = rsm;                                    // d (padded to even)
    double* lkeys = rsm + ((d + 1) & ~1);                // RR_WARPS x K
    int* lids = reinterpret_cast<int*>(lkeys + RR_WARPS * K);
    __shared__ int s_u;
    __shared__ int s_counts[RR_WARPS];
    __shared__ double s_qn;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* mykeys = lkeys + warp * K;
    int* myids = lids + warp * K;
    const int64_t total_units = unit_off[nqc];
    int64_t cur_q = -1;
    for (;;) {
        if (tid == 0) s_u = atomicAdd(next_unit, 1);
        __syncthreads();
        const int64_t u = s_u;
        if (u >= total_units) break;
        // unit -> (query, segment): last query whose first unit is <= u
        int64_t lo = 0, hi = nqc - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (unit_off[mid] <= u) lo = mid; else hi = mid - 1;
        }
        const int64_t q = q0 + lo;
        const int sgm = (int)(u - unit_off[lo]);
        if (q != cur_q) {                                // units of one query are consecutive: q is usually unchanged
            for (int j = tid; j < d; j += RR_THREADS) qs[j] = Q[q * d + j];
            cur_q = q;
            __syncthreads();
            if (METRIC == DPF_METRIC_ANGULAR) {
                if (warp == 0) {
                    double s = 0;
                    for (int j = lane; j < d; j += 32) s = fma(qs[j], qs[j], s);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    if (lane == 0) s_qn = sqrt(s);
                }
                __syncthreads();
            }
        }
        const int n_c = cnt[q];
        const int c_begin = sgm * seg, c_end = min(n_c, c_begin + seg);
        const int32_t* cq = cand + (off[q] - base);
        const double qn = (METRIC == DPF_METRIC_ANGULAR) ? s_qn : 1.0;
        int count = 0;
        // a warp takes 32 consecutive candidates per round (one coalesced id load), RR_ROWS rows in flight
        for (int c0 = c_begin + warp * 32; c0 < c_end; c0 += RR_WARPS * 32) {
            const int nloc = min(32, c_end - c0);
            const int myid = (lane < nloc) ? __ldg(cq + c0 + lane) : 0;
            for (int r0 = 0; r0 < nloc; r0 += RR_ROWS) {
                int id[RR_ROWS];
#pragma unroll
                for (int r = 0; r < RR_ROWS; ++r) id[r] = __shfl_sync(0xffffffffu, myid, min(r0 + r, nloc - 1));
                double s[RR_ROWS], xn[RR_ROWS];
                score_rows<VEC2, METRIC>(X, d, qs, id, lane, s, xn);
#pragma unroll
                for (int r = 0; r < RR_ROWS; ++r) {
                    if (r0 + r < nloc) {
                        double v = s[r];
                        if (METRIC == DPF_METRIC_ANGULAR) v = v / (qn * sqrt(xn[r]));
                        const double key = (METRIC == DPF_METRIC_L2) ? -v : v;
                        if (key == key) warp_insert(mykeys, myids, count, K, key, id[r], lane);   // NaN is never ranked
                    }
                }
            }
        }
        // merge the per-warp lists into this unit's sorted partial list (K rounds of "best head", warp 0)
        if (lane == 0) s_counts[warp] = count;
        __syncthreads();
        if (warp == 0) {
            int head = 0;
            const int mycount = lane < RR_WARPS ? s_counts[lane] : 0;
            for (int r = 0; r < K; ++r) {
                double bk = 0;
                int bi = 0x7fffffff, bl = -1;
                if (lane < RR_WARPS && head < mycount) { bk = lkeys[lane * K + head]; bi = lids[lane * K + head]; bl = lane; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ok_ = __shfl_xor_sync(0xffffffffu, bk, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
                    if (ol >= 0 && (bl < 0 || better(ok_, oi, bk, bi) || (ok_ == bk && oi == bi && ol < bl))) { bk = ok_; bi = oi; bl = ol; }
                }
                if (lane == 0) { part_key[u * K + r] = bk; part_id[u * K + r] = bl >= 0 ? bi : -1; }
                if (lane == bl) head++;
            }
        }
        __syncthreads();
    }
}

// final selection: thread g walks partial list g of its query (lists are sorted); K rounds of block-wide best head
__global__ void __launch_bounds__(RR_MAX_UNITS)
k_topk_select(const int64_t* __restrict__ unit_off, int64_t q0, int K, int metric, const double* __restrict__ part_key,
              const int32_t* __restrict__ part_id, int32_t* __restrict__ ids_out, double* __restrict__ score_out) {
    __shared__ double wk[RR_MAX_UNITS / 32];
    __shared__ int wi[RR_MAX_UNITS / 32], wl[RR_MAX_UNITS / 32];
    const int64_t ql = blockIdx.x, q = q0 + ql;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t u0 = unit_off[ql];
    const int G = (int)(unit_off[ql + 1] - u0);
    int head = 0;
    for (int r = 0; r < K; ++r) {
        double bk = 0;
        int bi = 0x7fffffff, bl = -1;
        if (tid < G && head < K) {
            const int id = part_id[(u0 + tid) * K + head];
            if (id >= 0) { bk = part_key[(u0 + tid) * K + head]; bi = id; bl = tid; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ok_ = __shfl_xor_sync(0xffffffffu, bk, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
            if (ol >= 0 && (bl < 0 || better(ok_, oi, bk, bi) || (ok_ == bk && oi == bi && ol < bl))) { bk = ok_; bi = oi; bl = ol; }
        }
        if (lane == 0) { wk[warp] = bk; wi[warp] = bi; wl[warp] = bl; }
        __syncthreads();
        bk = wk[0]; bi = wi[0]; bl = wl[0];
#pragma unroll
        for (int w = 1; w < RR_MAX_UNITS / 32; ++w)
            if (wl[w] >= 0 && (bl < 0 || better(wk[w], wi[w], bk, bi))) { bk = wk[w]; bi = wi[w]; bl = wl[w]; }
        if (tid == 0) {
            ids_out[q * K + r] = bl >= 0 ? bi : -1;
            score_out[q * K + r] = bl >= 0 ? (metric == DPF_METRIC_L2 ? -bk : bk) : __longlong_as_double(0x7ff8000000000000LL);
        }
        if (tid == bl) head++;
        __syncthreads();
    }
}

__global__ void k_counts_from_offsets(const int64_t* __restrict__ off, int64_t nq, int32_t* __restrict__ cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) cnt[i] = (int32_t)(off[i + 1] - off[i]);
}

template <bool VEC2, int METRIC>
static void launch_rerank_units(dpf_index* h, int grid, size_t smem, const double* Qd, int64_t q0, int64_t nqc, int64_t base,
                                const int64_t* off, const int32_t* cnt, const int32_t* cand, int seg, int topk,
                                int* next_unit) {
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        DPF_CUDA(cudaFuncSetAttribute(k_rerank_units<VEC2, METRIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    k_rerank_units<VEC2, METRIC><<<grid, RR_THREADS, smem, h->stream>>>(h->Xdev, h->cfg.d, Qd, q0, nqc, base, off, cnt, cand,
                                                                        h->unit_off.p, seg, topk, h->part_key.p,
                                                                        h->part_id.p, next_unit);
    DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
}

// max_cnt: upper bound of the candidate count of any query in [q0, q1) (sizes the units)
void rerank_topk(dpf_index* h, const double* Qd, int64_t q0, int64_t q1, int64_t base, const int64_t* off,
                 const int32_t* cnt, const int32_t* cand, int64_t max_cnt, int64_t total_ub, int topk, int metric,
                 int32_t* ids_out, double* score_out) {
    const int64_t nqc = q1 - q0;
    DPF_REQUIRE(topk >= 1 && topk <= RR_MAXK, DPF_ERR_INVALID, "topk must be in 1..256");
    DPF_REQUIRE(h->dense && h->Xdev, DPF_ERR_STATE, "re-rank needs a dense index");
    if (nqc <= 0) return;
    StageTimer tm(h, DPF_T_RERANK);
    const int d = h->cfg.d;
    cudaStream_t st = h->stream;
    const int seg = (int)std::max<int64_t>(RR_MIN_SEG, (max_cnt + RR_MAX_UNITS - 1) / RR_MAX_UNITS);
    const int64_t units_ub = total_ub / seg + nqc + 1;
    h->ucnt.reserve(nqc + 1);
    h->unit_off.reserve(nqc + 1);
    h->part_key.reserve((size_t)units_ub * topk);
    h->part_id.reserve((size_t)units_ub * topk);
    h->counters.reserve(64);
    int* next_unit = h->counters.p + 17;
    DPF_CUDA(cudaMemsetAsync(next_unit, 0, sizeof(int), st));
    k_unit_counts<<<(unsigned)((nqc + 255) / 256), 256, 0, st>>>(cnt, q0, nqc, seg, h->ucnt.p); DPF_LAUNCHED();
    exclusive_scan_i64(h, h->ucnt.p, h->unit_off.p, nqc);
    const size_t smem = (size_t)((d + 1) & ~1) * sizeof(double) + (size_t)RR_WARPS * topk * (sizeof(double) + sizeof(int));
    const bool vec2 = (d % 2 == 0) && ((reinterpret_cast<uintptr_t>(h->Xdev) & 15) == 0);
    const int grid = (int)std::min<int64_t>(units_ub, (int64_t)h->num_sms * 6);
#define DPF_RR(V, M) launch_rerank_units<V, M>(h, grid, smem, Qd, q0, nqc, base, off, cnt, cand, seg, topk, next_unit)
    if (vec2) {
        if (metric == DPF_METRIC_DOT) DPF_RR(true, DPF_METRIC_DOT);
        else if (metric == DPF_METRIC_ANGULAR) DPF_RR(true, DPF_METRIC_ANGULAR);
        else DPF_RR(true, DPF_METRIC_L2);
    } else {
        if (metric == DPF_METRIC_DOT) DPF_RR(false, DPF_METRIC_DOT);
        else if (metric == DPF_METRIC_ANGULAR) DPF_RR(false, DPF_METRIC_ANGULAR);
        else DPF_RR(false, DPF_METRIC_L2);
    }
#undef DPF_RR
    k_topk_select<<<(unsigned)nqc, RR_MAX_UNITS, 0, st>>>(h->unit_off.p, q0, topk, metric, h->part_key.p, h->part_id.p,
                                                           ids_out, score_out); DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
}

void counts_from_offsets(dpf_index* h, const int64_t* off, int64_t nq, int32_t* cnt) {
    k_counts_from_offsets<<<(unsigned)((nq + 255) / 256), 256, 0, h->stream>>>(off, nq, cnt); DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------
// multi-GPU: merge G per-GPU top-k lists per query; the same id can arrive from several GPUs (reached through
// tables whose sub-index lives on different GPUs) with a bit-identical score, and is kept once.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_merge_topk(const int32_t* __restrict__ gids, const double* __restrict__ gsc, int G, int64_t nq, int K, int metric,
             int32_t* __restrict__ ids_out, double* __restrict__ score_out) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    // lane g < G walks list g (each list is already in result order)
    int head = 0, last = -1;
    for (int r = 0; r < K; ++r) {
        double bk;
        int bi, bl;
        for (;;) {
            bk = 0; bi = 0x7fffffff; bl = -1;
            for (int g0 = 0; g0 < G; g0 += 32) {       // G <= 32 in practice; loop kept for generality
                const int g = g0 + lane;
                if (g < G && head < K) {
                    const int id = gids[((int64_t)g * nq + q) * K + head];
                    if (id >= 0) {
                        const double s = gsc[((int64_t)g * nq + q) * K + head];
                        bk = (metric == DPF_METRIC_L2) ? -s : s; bi = id; bl = lane;
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ok_ = __shfl_xor_sync(0xffffffffu, bk, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
                if (ol >= 0 && (bl < 0 || better(ok_, oi, bk, bi) || (ok_ == bk && oi == bi && ol < bl))) { bk = ok_; bi = oi; bl = ol; }
            }
            if (bl < 0) break;
            if (lane == bl) head++;
            if (bi != last) break;                     // duplicate of the id just emitted: skip it
        }
        if (lane == 0) {
            ids_out[q * K + r] = bl >= 0 ? bi : -1;
            score_out[q * K + r] = bl >= 0 ? (metric == DPF_METRIC_L2 ? -bk : bk) : __longlong_as_double(0x7ff8000000000000LL);
        }
        if (bl < 0) {
            for (int r2 = r + 1; r2 < K; ++r2)
                if (lane == 0) { ids_out[q * K + r2] = -1; score_out[q * K + r2] = __longlong_as_double(0x7ff8000000000000LL); }
            break;
        }
        last = bi;
    }
}

void merge_topk(dpf_index* h, const int32_t* gids, const double* gsc, int G, int64_t nq, int topk, int metric,
                int32_t* ids_out, double* score_out) {
    DPF_REQUIRE(G >= 1 && G <= 32, DPF_ERR_INVALID, "merge supports 1..32 lists");
    if (nq <= 0) return;
    k_merge_topk<<<(unsigned)((nq + 3) / 4), 128, 0, h->stream>>>(gids, gsc, G, nq, topk, metric, ids_out, score_out); DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
}


// ---------------------------------------------------------------------------------------------------------
// K5b: bucket-major re-rank on the FP64 tensor pipe.
//
// The row-major kernel above reads a candidate row once per (query, candidate): 8d bytes of HBM for 2d flops.
// In a batch, many queries probe the same leaf buckets (a bucket is probed by every query whose key falls in
// it or one bit away), so the same rows are fetched again and again.  Here the batch is regrouped by bucket:
// the (bucket, query) pairs found by the probe are sorted by bucket; a CTA takes QT consecutive pairs, and for
// each run of pairs that share a bucket it stages the bucket's rows in shared memory once and multiplies them
// against the run's queries with DMMA (rows = M, queries = N), writing one score per (pair, row).  A second
// kernel selects each query's top-k from its score segments, de-duplicating ids that were reached through
// several tables (their scores are bit-identical: same rows, same k order).  Per step this replaces
// nC_q * 8d bytes per query by ~(bucket rows * 8d) per ~QT queries plus 16 B per (query, candidate).
// Candidate *sets* are unchanged (same probe), so results equal the row-major path up to summation order.
// ---------------------------------------------------------------------------------------------------------
constexpr int BM_QT = 32;              // pairs (queries) per CTA group = MMA N extent
constexpr int BM_RT = 32;              // bucket rows per stage       = MMA M extent
constexpr int BM_KC = 128;             // largest supported d
constexpr int BM_PITCH = BM_KC + 4;    // (4g + t) mod 16 distinct => conflict-free LDS.64 fragment loads
constexpr int BM_THREADS = 256;        // 8 warps: 4 row blocks x 2 query halves

bool bucket_major_supported(const dpf_index* h, int metric, int topk) {
    const char* e = getenv("DPF_RERANK");
    if (e && e[0] == 'r') return false;                        // DPF_RERANK=rowmajor forces the row-major kernel
    return h->dense && h->Xdev && h->cfg.d <= BM_KC && (metric == DPF_METRIC_DOT || metric == DPF_METRIC_ANGULAR) &&
           topk <= RR_MAXK;
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

__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// stage `nrows` rows (ids[0..nrows)) of X into dst[r][0..dpad) (pitch BM_PITCH); columns d..dpad are pre-zeroed
__device__ __forceinline__ void stage_rows(double* dst, const double* __restrict__ X, const int32_t* __restrict__ ids,
                                           int nrows, int d, bool vec16, int tid) {
    if (vec16) {
        const int cpr = d >> 1;                                  // 16-byte chunks per row
        for (int i = tid; i < nrows * cpr; i += BM_THREADS) {
            const int r = i / cpr, c = i - r * cpr;
            cp_async_16(dst + r * BM_PITCH + 2 * c, X + (int64_t)__ldg(ids + r) * d + 2 * c);
        }
    } else {
        for (int i = tid; i < nrows * d; i += BM_THREADS) {
            const int r = i / d, c = i - r * d;
            cp_async_8(dst + r * BM_PITCH + c, X + (int64_t)__ldg(ids + r) * d + c);
        }
    }
}

template <bool ANGULAR>
__global__ void __launch_bounds__(BM_THREADS, 2)
k_score_groups(const double* __restrict__ X, int d, const double* __restrict__ Q,
               const unsigned long long* __restrict__ pair_key /* sorted by bucket */, int64_t npairs,
               const int32_t* __restrict__ pair_q, const uint32_t* __restrict__ pair_len,
               const uint32_t* __restrict__ pair_seg /* exclusive scan of pair_len in pair-index order */,
               const int32_t* __restrict__ ids_sorted, double* __restrict__ scores,
               unsigned long long* __restrict__ stat /* [0] runs, [1] rows staged */) {
    extern __shared__ double bsm[];
    double* Rs = bsm;                                  // [2][BM_RT][BM_PITCH]
    double* Qs = bsm + 2 * BM_RT * BM_PITCH;           // [BM_QT][BM_PITCH]
    __shared__ unsigned long long s_key[BM_QT];
    __shared__ uint32_t s_seg[BM_QT];
    __shared__ double s_qn[BM_QT], s_xn[2][BM_RT];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;           // rows 8*wm.., queries 16*wn..
    const int64_t p0 = (int64_t)blockIdx.x * BM_QT;
    const int np = (int)min((int64_t)BM_QT, npairs - p0);
    const int dpad = (d + 3) & ~3;
    const bool vec16 = (d & 1) == 0;
    if (tid < np) s_key[tid] = pair_key[p0 + tid];
    // zero the k padding (and unused rows/queries) once: cp.async only ever writes columns < d
    for (int i = tid; i < (2 * BM_RT + BM_QT) * BM_PITCH; i += BM_THREADS) bsm[i] = 0.0;
    __syncthreads();
    int j0 = 0;
    while (j0 < np) {
        // run of pairs that share the bucket of pair j0
        const unsigned long long k0 = s_key[j0];
        const uint32_t bstart = (uint32_t)(k0 >> 32);
        int j1 = j0 + 1;
        while (j1 < np && (uint32_t)(s_key[j1] >> 32) == bstart) j1++;
        const int m = j1 - j0;                          // queries of this run (MMA N extent, <= BM_QT)
        const uint32_t i0 = (uint32_t)k0;               // pair index of the run's first pair
        const int blen = (int)pair_len[i0];             // bucket length (identical for the whole run)
        const int32_t* bids = ids_sorted + bstart;
        if (tid == 0) { atomicAdd(&stat[0], 1ULL); atomicAdd(&stat[1], (unsigned long long)blen); }
        // stage the run's queries
        for (int i = tid; i < m * d; i += BM_THREADS) {
            const int r = i / d, cc = i - r * d;
            const uint32_t pi = (uint32_t)s_key[j0 + r];
            Qs[r * BM_PITCH + cc] = __ldg(Q + (int64_t)pair_q[pi] * d + cc);
        }
        if (tid < m) s_seg[tid] = pair_seg[(uint32_t)s_key[j0 + tid]];
        const int nblk = (blen + BM_RT - 1) / BM_RT;
        stage_rows(Rs, X, bids, min(BM_RT, blen), d, vec16, tid);
        cp_async_commit();
        __syncthreads();
        if (ANGULAR) {
            for (int r = warp; r < m; r += BM_THREADS / 32) {
                double s = 0;
                for (int cc = lane; cc < d; cc += 32) s = fma(Qs[r * BM_PITCH + cc], Qs[r * BM_PITCH + cc], s);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) s_qn[r] = sqrt(s);
            }
        }
        const int nb_n = (m + 7) >> 3;                  // 8-query blocks in use
        for (int b = 0; b < nblk; ++b) {
            const int buf = b & 1;
            const int rows = min(BM_RT, blen - b * BM_RT);
            if (b + 1 < nblk) {
                stage_rows(Rs + (buf ^ 1) * BM_RT * BM_PITCH, X, bids + (b + 1) * BM_RT, min(BM_RT, blen - (b + 1) * BM_RT), d,
                           vec16, tid);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            const double* rs = Rs + buf * BM_RT * BM_PITCH;
            if (ANGULAR) {
                for (int r = warp; r < rows; r += BM_THREADS / 32) {
                    double s = 0;
                    for (int cc = lane; cc < d; cc += 32) s = fma(rs[r * BM_PITCH + cc], rs[r * BM_PITCH + cc], s);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    if (lane == 0) s_xn[buf][r] = sqrt(s);
                }
                __syncthreads();
            }
            if (8 * wm < rows && 2 * wn < nb_n) {
                double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
                const double* ap = rs + (8 * wm + g) * BM_PITCH + t;
                const double* bp = Qs + (16 * wn + g) * BM_PITCH + t;
                const bool two = (2 * wn + 1) < nb_n;
#pragma unroll 8
                for (int k = 0; k < dpad; k += 4) {
                    const double a = ap[k];
                    const double b0 = bp[k];
                    dmma884(acc[0][0], acc[0][1], a, b0);
                    if (two) {
                        const double b1 = bp[8 * BM_PITCH + k];
                        dmma884(acc[1][0], acc[1][1], a, b1);
                    }
                }
                // thread holds (row 8wm+g, query 16wn + 8jb + 2t + e)
                const int row = 8 * wm + g;
                if (row < rows) {
#pragma unroll
                    for (int jb = 0; jb < 2; ++jb)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int qi = 16 * wn + 8 * jb + 2 * t + e;
                            if (qi < m) {
                                double v = acc[jb][e];
                                if (ANGULAR) v = v / (s_qn[qi] * s_xn[buf][row]);
                                scores[(int64_t)s_seg[qi] + b * BM_RT + row] = v;
                            }
                        }
                }
            }
            __syncthreads();   // the buffer computed on is refilled two iterations later; Qs/s_seg reused by the next run
        }
        j0 = j1;
    }
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
    for (uint32_t p = pbeg + warp; p < pend; p += RR_WARPS) {
        const uint32_t bstart = (uint32_t)(pair_key_unsorted[p] >> 32);
        const int len = (int)pair_len[p];
        const double* sc = scores + pair_seg[p];
        const int32_t* ids = ids_sorted + bstart;
        for (int j0 = 0; j0 < len; j0 += 32) {
            const int j = j0 + lane;
            double key = 0;
            int id = -1;
            bool cand = false;
            if (j < len) {
                key = sc[j];
                id = __ldg(ids + j);
                cand = (key == key) && !(excl && id == qid);
                if (cand && count == K) cand = better(key, id, mykeys[K - 1], myids[K - 1]);
            }
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

void topk_bucket_major(dpf_index* h, const double* Qd, const QueryKeys& qk, int steps, int probe_mode, int64_t q0, int64_t q1,
                       int64_t entries_ub, int topk, int metric, int32_t* ids_out, double* score_out) {
    const ProbeCtx c = make_ctx(h, steps, probe_mode);
    cudaStream_t st = h->stream;
    const int64_t nqc = q1 - q0;
    const int L = c.L, d = h->cfg.d;
    if (nqc <= 0) return;
    DPF_REQUIRE(h->h_table_base[L] < (1LL << 32), DPF_ERR_INVALID, "bucket-major re-rank: more than 2^32 forest entries");
    DPF_REQUIRE(entries_ub < (1LL << 32), DPF_ERR_INVALID, "bucket-major re-rank: chunk too large");
    // pair offsets of this chunk = exclusive scan of the per-(query, table) bucket counts from the probe pass
    const int64_t nslots = nqc * L + 1;
    h->pair_base.reserve(nslots);
    {
        StageTimer tm(h, DPF_T_EXPAND);
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
            k_topk_select<<<(unsigned)nqc, RR_MAX_UNITS, 0, st>>>(h->unit_off.p, q0, topk, metric, nullptr, nullptr, ids_out, score_out); DPF_LAUNCHED();
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
    }
    {
        StageTimer tm(h, DPF_T_RERANK);
        const int64_t npairs = h->bm_npairs;
        const size_t smem = (size_t)(2 * BM_RT + BM_QT) * BM_PITCH * sizeof(double);
        static bool attr = false;
        if (!attr) {
            DPF_CUDA(cudaFuncSetAttribute(k_score_groups<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            DPF_CUDA(cudaFuncSetAttribute(k_score_groups<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr = true;
        }
        const unsigned grid = (unsigned)((npairs + BM_QT - 1) / BM_QT);
        unsigned long long* bm_stat = reinterpret_cast<unsigned long long*>(h->counters.p + 26);   // cleared by probe_count_all
        h->stats[DPF_STAT_BM_PAIRS] += npairs;
        const char* ev = getenv("DPF_BM_KERNEL");
        const bool use_warps = (d % 2 == 0) && !(ev && ev[0] == 'c');   // DPF_BM_KERNEL=cta selects the staged kernel
        if (use_warps) {
            const char* nbv = getenv("DPF_BM_NB");
            const int nb = nbv ? atoi(nbv) : 2;
            const bool ang = metric == DPF_METRIC_ANGULAR;
            if (nb <= 1) { if (ang) launch_score_warps<true, 1>(h, Qd, npairs, metric, bm_stat); else launch_score_warps<false, 1>(h, Qd, npairs, metric, bm_stat); }
            else if (nb == 2) { if (ang) launch_score_warps<true, 2>(h, Qd, npairs, metric, bm_stat); else launch_score_warps<false, 2>(h, Qd, npairs, metric, bm_stat); }
            else { if (ang) launch_score_warps<true, 4>(h, Qd, npairs, metric, bm_stat); else launch_score_warps<false, 4>(h, Qd, npairs, metric, bm_stat); }
        } else if (metric == DPF_METRIC_ANGULAR)
            k_score_groups<true><<<grid, BM_THREADS, smem, st>>>(h->Xdev, d, Qd, h->bm_sorted, npairs, h->pair_q.p, h->pair_len.p,
                                                                 h->pair_seg.p, h->ids_sorted.p, h->scores.p, bm_stat);
        else
            k_score_groups<false><<<grid, BM_THREADS, smem, st>>>(h->Xdev, d, Qd, h->bm_sorted, npairs, h->pair_q.p, h->pair_len.p,
                                                                  h->pair_seg.p, h->ids_sorted.p, h->scores.p, bm_stat);
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
