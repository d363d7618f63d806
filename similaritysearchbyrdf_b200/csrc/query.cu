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

#include "rerank_units.cuh"

namespace dpf {

// pass A: upper bound of the candidate count per query (sum of distinct bucket sizes over tables)
__global__ void __launch_bounds__(256)
k_probe_count(ProbeCtx c, const int32_t* __restrict__ qkeys, const uint8_t* __restrict__ qpids, int64_t ld, int64_t nq,
              int32_t* __restrict__ q_ub, unsigned long long* __restrict__ stat /* [0] nlz>28 */) {
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
    int total = 0;
    const int np = 1 << c.tp.pb;
    for (int sub = 0; sub < np; ++sub) {       // findStepWiseSubIndexIDs (RandomDrawTreeMap.java:613-621)
        if (__popc(sub ^ pid) > c.steps) continue;
        if (!c.own.has(t, sub)) continue;   // this GPU's sub-forest only
        bool leader;
        int ptr, cnt;
        warp_lookup(c, t, sub, seg, h, nprobes, lane, leader, ptr, cnt);
        total += leader ? cnt : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if (lane == 0 && total > 0) atomicAdd(&q_ub[q], total);      // (the batch total is the last scanned offset: no global counter)
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
                if (!c.own.has(t, sub)) continue;
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
    h->counters.reserve(CTR_COUNT);
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

// pass A over the whole batch: h->q_cnt = per-query upper bound, h->q_off = its exclusive scan (device) and
// off_host = the same offsets on the host, from which the caller cuts the batch into memory-bounded chunks
void probe_count_all(dpf_index* h, const QueryKeys& qk, int steps, int probe_mode, std::vector<int64_t>& off_host) {
    const int64_t nq = qk.nq;
    const ProbeCtx c = make_ctx(h, steps, probe_mode);
    cudaStream_t st = h->stream;
    h->q_cnt.reserve(nq + 1);
    h->q_off.reserve(nq + 1);
    h->counters.reserve(CTR_COUNT);
    unsigned long long* stat = reinterpret_cast<unsigned long long*>(h->counters.p + CTR_STAT_NLZ);
    DPF_CUDA(cudaMemsetAsync(h->q_cnt.p, 0, (nq + 1) * sizeof(int32_t), st));
    {
        StageTimer tm(h, DPF_T_PROBE_COUNT);
        const int64_t warps = nq * c.L;
        k_probe_count<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(c, qk.keys, h->qpids.p, qk.ld, nq, h->q_cnt.p, stat); DPF_LAUNCHED();
        DPF_CUDA(cudaGetLastError());
        exclusive_scan_i64(h, h->q_cnt.p, h->q_off.p, nq);
    }
    off_host.resize((size_t)nq + 1);
    DPF_CUDA(cudaMemcpyAsync(off_host.data(), h->q_off.p, (nq + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    DPF_CUDA(cudaMemcpyAsync(h->counters.p + CTR_ENTRIES, h->q_off.p + nq, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    DPF_CUDA(cudaStreamSynchronize(st));
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
        if (smem_need > 48 * 1024)   // per device, hence per launch
            DPF_CUDA(cudaFuncSetAttribute(k_expand<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_need));
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

// ---------------------------------------------------------------------------------------------------------
// k_topk_direct — exhaustive top k of one query over the buckets it probes, straight from the probe result (leaf numbers
// per table): the answer for the queries the threshold filter cannot serve (fewer than k sampled rows, survivors that
// did not fit the pool, unit records that did not fit; listed in a DirtySet on the device), with no scratch that depends
// on the data.  Work item = (dirty query, table group): persistent CTAs pull items, a warp per bucket, FP64 rows scored
// like the row-major kernel, one sorted partial list per item; k_merge_topk then merges a query's DIRECT_PARTS lists
// (an id reached through several tables has the same score every time and is kept once).  An empty dirty list costs two
// launches that return at once; a single dirty query is spread over DIRECT_PARTS CTAs instead of occupying one for
// milliseconds.
// ---------------------------------------------------------------------------------------------------------
constexpr int DIRECT_PARTS = 32;

template <bool VEC2, int METRIC>
__global__ void __launch_bounds__(RR_THREADS)
k_topk_direct(const double* __restrict__ X, int d, ChunkView cv, int L, const uint32_t* __restrict__ leaf_pos,
              const int32_t* __restrict__ leaf_len, const int32_t* __restrict__ ids_sorted, DirtySet dirty, int first, int cap,
              int self_exclude, int K, double* __restrict__ part_key, int32_t* __restrict__ part_id) {
    extern __shared__ double rsm[];
    double* qs = rsm;                                    // d (padded to even)
    double* lkeys = rsm + ((d + 1) & ~1);                // RR_WARPS x K
    int* lids = reinterpret_cast<int*>(lkeys + RR_WARPS * K);
    __shared__ int s_counts[RR_WARPS];
    __shared__ double s_qn;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* mykeys = lkeys + warp * K;
    int* myids = lids + warp * K;
    const int ndirty = min(*dirty.count - first, cap);   // this round's slice of the list
    for (int item = blockIdx.x; item < ndirty * DIRECT_PARTS; item += gridDim.x) {
        const int di = item / DIRECT_PARTS, part = item % DIRECT_PARTS;
        const int64_t q = dirty.list[first + di];
        __syncthreads();                                 // shared lists and query of the previous item are free
        for (int j = tid; j < d; j += RR_THREADS) qs[j] = cv.Q[q * d + j];
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
        const double qn = (METRIC == DPF_METRIC_ANGULAR) ? s_qn : 1.0;
        const int qid = cv.qids ? cv.qids[q] : INT32_MIN;
        const bool excl = self_exclude && cv.qids && qid >= -128 && qid <= 127;
        int count = 0;
        for (int t = part; t < L; t += DIRECT_PARTS) {
            const uint32_t nb = cv.pair_cnt[q * L + t];
            for (uint32_t e = 0; e < nb; ++e) {          // every warp takes 32-row slices of every bucket: balanced whatever
                const uint32_t leaf = cv.cache[(q * L + t) * cv.cap + e];   // the bucket sizes
                const int32_t* bids = ids_sorted + leaf_pos[leaf];
                const int len = leaf_len[leaf];
                for (int c0 = 32 * warp; c0 < len; c0 += 32 * RR_WARPS) {
                    const int nloc = min(32, len - c0);
                    const int myid = (lane < nloc) ? __ldg(bids + c0 + lane) : 0;
                    for (int r0 = 0; r0 < nloc; r0 += RR_ROWS) {
                        int id[RR_ROWS];
#pragma unroll
                        for (int r = 0; r < RR_ROWS; ++r) id[r] = __shfl_sync(0xffffffffu, myid, min(r0 + r, nloc - 1));
                        double s[RR_ROWS], xn[RR_ROWS];
                        score_rows<VEC2, METRIC>(X, d, qs, id, lane, s, xn);
#pragma unroll
                        for (int r = 0; r < RR_ROWS; ++r) {
                            if (r0 + r >= nloc) continue;
                            double v = s[r];
                            if (METRIC == DPF_METRIC_ANGULAR) v = v / (qn * sqrt(xn[r]));
                            const double key = (METRIC == DPF_METRIC_L2) ? -v : v;
                            if (!(key == key) || (excl && id[r] == qid)) continue;            // NaN is never ranked
                            if (count == K && !better(key, id[r], mykeys[K - 1], myids[K - 1])) continue;
                            bool dup = false;
                            for (int b2 = 0; b2 < count; b2 += 32) dup |= __any_sync(0xffffffffu, b2 + lane < count && myids[b2 + lane] == id[r]);
                            if (!dup) warp_insert(mykeys, myids, count, K, key, id[r], lane);
                        }
                    }
                }
            }
        }
        if (lane == 0) s_counts[warp] = count;
        __syncthreads();
        if (warp == 0) {                                 // merge the warps' lists into the item's sorted partial list
            int head = 0, last = -1;
            const int mycount = lane < RR_WARPS ? s_counts[lane] : 0;
            double* ok_out = part_key + (size_t)item * K;
            int32_t* oi_out = part_id + (size_t)item * K;
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
                    oi_out[r] = bl >= 0 ? bi : -1;
                    ok_out[r] = bl >= 0 ? (METRIC == DPF_METRIC_L2 ? -bk : bk) : __longlong_as_double(0x7ff8000000000000LL);
                }
                if (bl >= 0) last = bi;
            }
        }
    }
}

// warp per dirty query of the round: merge its DIRECT_PARTS partial lists into the query's row of the output
__global__ void __launch_bounds__(128)
k_direct_merge(const int32_t* __restrict__ part_id, const double* __restrict__ part_key, DirtySet dirty, int first, int cap, int K,
               int metric, int32_t* __restrict__ ids_out, double* __restrict__ score_out, int* __restrict__ stat_direct) {
    const int lane = threadIdx.x & 31;
    const int ndirty = min(*dirty.count - first, cap);
    if (blockIdx.x == 0 && threadIdx.x == 0 && ndirty > 0) atomicAdd(stat_direct, ndirty);
    for (int di = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); di < ndirty; di += gridDim.x * (blockDim.x >> 5)) {
        const int64_t q = dirty.list[first + di];
        const int32_t* lids = part_id + ((size_t)di * DIRECT_PARTS + lane) * K;
        const double* lsc = part_key + ((size_t)di * DIRECT_PARTS + lane) * K;
        int head = 0;
        for (int r = 0; r < K; ++r) {
            double bk;
            int bi, bl;
            for (;;) {
                bk = 0; bi = 0x7fffffff; bl = -1;
                if (head < K) {
                    const int id = lids[head];
                    if (id >= 0) { const double s = lsc[head]; bk = (metric == DPF_METRIC_L2) ? -s : s; bi = id; bl = lane; }
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
                bool dup = false;                        // the same id through tables of two groups: same score, kept once
                for (int j = lane; j < r; j += 32) dup |= ids_out[q * K + j] == bi;
                if (!__any_sync(0xffffffffu, dup)) break;
            }
            if (lane == 0) {
                ids_out[q * K + r] = bl >= 0 ? bi : -1;
                score_out[q * K + r] = bl >= 0 ? (metric == DPF_METRIC_L2 ? -bk : bk) : __longlong_as_double(0x7ff8000000000000LL);
            }
            __syncwarp();
        }
    }
}

void topk_direct(dpf_index* h, const ChunkView& cv, const DirtySet& dirty, int topk, int metric, int32_t* ids_out, double* score_out) {
    static_assert(DIRECT_PARTS == 32, "one lane per partial list in k_direct_merge");
    const int d = h->cfg.d;
    const size_t smem = (size_t)((d + 1) & ~1) * sizeof(double) + (size_t)RR_WARPS * topk * (sizeof(double) + sizeof(int));
    const bool vec2 = (d % 2 == 0) && ((reinterpret_cast<uintptr_t>(h->Xdev) & 15) == 0);
    // partial lists for `cap` dirty queries per round (<= 128 MB); the rounds beyond the list's length return at once
    const int64_t cap = std::max<int64_t>(1, std::min<int64_t>(cv.nqc, (128LL << 20) / ((int64_t)DIRECT_PARTS * topk * 12)));
    h->direct_keys.reserve((size_t)cap * DIRECT_PARTS * topk);
    h->direct_ids.reserve((size_t)cap * DIRECT_PARTS * topk);
    const int grid = (int)std::min<int64_t>(cap * DIRECT_PARTS, (int64_t)h->num_sms * 4);
    for (int64_t first = 0; first < cv.nqc; first += cap) {
        auto go = [&](auto kern) {
            if (smem > 48 * 1024) DPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, RR_THREADS, smem, h->stream>>>(h->Xdev, d, cv, h->cfg.L, h->leaf_pos.p, h->leaf_len.p, h->ids_sorted.p, dirty,
                                                        (int)first, (int)cap, h->cfg.self_exclude_small_ids, topk, h->direct_keys.p,
                                                        h->direct_ids.p);
            DPF_LAUNCHED();
        };
        if (vec2) {
            if (metric == DPF_METRIC_DOT) go(k_topk_direct<true, DPF_METRIC_DOT>);
            else if (metric == DPF_METRIC_ANGULAR) go(k_topk_direct<true, DPF_METRIC_ANGULAR>);
            else go(k_topk_direct<true, DPF_METRIC_L2>);
        } else {
            if (metric == DPF_METRIC_DOT) go(k_topk_direct<false, DPF_METRIC_DOT>);
            else if (metric == DPF_METRIC_ANGULAR) go(k_topk_direct<false, DPF_METRIC_ANGULAR>);
            else go(k_topk_direct<false, DPF_METRIC_L2>);
        }
        k_direct_merge<<<(unsigned)std::min<int64_t>((cap + 3) / 4, 1024), 128, 0, h->stream>>>(
            h->direct_ids.p, h->direct_keys.p, dirty, (int)first, (int)cap, topk, metric, ids_out, score_out, h->counters.p + CTR_DIRECT);
        DPF_LAUNCHED();
        DPF_CUDA(cudaGetLastError());
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
    // (function attributes are per device: set per launch, not once per process)
    if (smem > 48 * 1024)
        DPF_CUDA(cudaFuncSetAttribute(k_rerank_units<VEC2, METRIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
    h->counters.reserve(CTR_COUNT);
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
k_merge_topk(const int32_t* __restrict__ gids, const double* __restrict__ gsc, int64_t stride_ids, int64_t stride_sc, int G,
             int64_t nq, int K, int metric, int32_t* __restrict__ ids_out, double* __restrict__ score_out) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    // lane g < G walks list g (each list is already in result order)
    int head = 0;
    for (int r = 0; r < K; ++r) {
        double bk;
        int bi, bl;
        for (;;) {
            bk = 0; bi = 0x7fffffff; bl = -1;
            for (int g0 = 0; g0 < G; g0 += 32) {       // G <= 32 in practice; loop kept for generality
                const int g = g0 + lane;
                if (g < G && head < K) {
                    const int id = gids[(int64_t)g * stride_ids + q * K + head];
                    if (id >= 0) {
                        const double s = gsc[(int64_t)g * stride_sc + q * K + head];
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
            // the same id from two GPUs: emitted once.  Its two scores are the same dot product, but not necessarily from
            // the same kernel (one GPU may have answered the query exhaustively, the other through the filter), so they can
            // differ in the last bits and need not be neighbours in the merged order: look the id up in what was emitted
            bool dup = false;
            for (int j = lane; j < r; j += 32) dup |= ids_out[q * K + j] == bi;
            if (!__any_sync(0xffffffffu, dup)) break;
        }
        if (lane == 0) {
            ids_out[q * K + r] = bl >= 0 ? bi : -1;
            score_out[q * K + r] = bl >= 0 ? (metric == DPF_METRIC_L2 ? -bk : bk) : __longlong_as_double(0x7ff8000000000000LL);
        }
        __syncwarp();
        if (bl < 0) {
            for (int r2 = r + 1; r2 < K; ++r2)
                if (lane == 0) { ids_out[q * K + r2] = -1; score_out[q * K + r2] = __longlong_as_double(0x7ff8000000000000LL); }
            break;
        }
    }
}

void merge_topk_strided(dpf_index* h, const int32_t* gids, const double* gsc, int64_t stride_ids, int64_t stride_sc, int G, int64_t nq,
                        int topk, int metric, int32_t* ids_out, double* score_out) {
    DPF_REQUIRE(G >= 1 && G <= 32, DPF_ERR_INVALID, "merge supports 1..32 lists");
    if (nq <= 0) return;
    k_merge_topk<<<(unsigned)((nq + 3) / 4), 128, 0, h->stream>>>(gids, gsc, stride_ids, stride_sc, G, nq, topk, metric, ids_out,
                                                                  score_out); DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
}

void merge_topk(dpf_index* h, const int32_t* gids, const double* gsc, int G, int64_t nq, int topk, int metric,
                int32_t* ids_out, double* score_out) {
    merge_topk_strided(h, gids, gsc, nq * topk, nq * topk, G, nq, topk, metric, ids_out, score_out);
}


}  // namespace dpf
