// rerank_wide.cu — bucket-major re-rank for vectors wider than BM_KC (= 128) columns: GIST-shaped data (d = 960).
//
// Replaces, for large batches, the row-major gather of `topKAndPrecisionScore`
// (src/main/scala/mclab/deploy/DensevectorRDFInit.scala:472-507): there every query reads each of its candidates' rows
// (nC_q * 8d bytes per query); here a bucket's rows are read once per unit of <= 16 queries that probe the bucket, so the
// HBM bytes of a batch shrink by the number of queries that share a bucket (6 at 10k queries / steps = 0 on a 1M index,
// 16 with multi-step search).  The pipeline around the two kernels below is the one of rerank_bm.cu (probe, counting sort
// of the pairs by leaf, unit records, thresholds, survivor pool, selection, exhaustive fallback): all of it is
// independent of d.
//
//   k_score_wide       warp per unit; the bucket's rows in slabs of 32 (four DMMA m-tiles), the unit's <= 16 queries as two
//                      n-blocks.  A slab of whole rows does not fit in shared memory next to sixteen 7.5 KB queries, so
//                      rows and queries stream through a per-warp ring 32 columns (256 contiguous bytes of every row and
//                      query) at a time, copied with cp.async half a warp per row, one to two stages ahead (rows: HBM;
//                      queries: L2 / L1 — a slab of 32 rows re-reads the unit's queries once, +50 % / +25 % of L2 traffic
//                      on top of the rows for two / one n-blocks).  64 DMMA.8x8x4 per stage.  Only scores that reach the
//                      query's threshold leave the kernel (SurvivorSink).
//   k_threshold_wide   the threshold samples of k_threshold (rerank_bm.cu) for any d: warp per (query, sampled table), the
//                      lanes stride over the 16-byte chunks of four rows at a time, plain FP64 FMAs, lower bound = score
//                      minus twice the rounding bound of a length-d dot product in any order.
//
// k permutation of the DMMA (as in k_score_stream): thread t of a row group takes the 16-byte chunks 4j + t of the row,
// k-step 2j multiplies the first double of those chunks, k-step 2j + 1 the second; rows and queries use the same
// permutation, so the sum over k is the dot product.
#include <type_traits>

#include "rerank_units.cuh"

namespace dpf {

constexpr int WD_WARPS = 6;                // warps per CTA, one CTA per SM (3 ring stages of 12 KB per warp)
constexpr int WD_MT = 4;                   // m-tiles (8 rows) per slab
constexpr int WD_SLAB = 8 * WD_MT;

__device__ __forceinline__ double2 ldg_d2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }
// rows stream through once per unit: do not let them wash the queries out of L1
__device__ __forceinline__ double2 ldg_d2_stream(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// Ring geometry.  A stage = 32 columns = 256 contiguous bytes (two 128-byte lines) of each of the slab's 32 rows and of the
// unit's 16 queries (12 KB).  Earlier forms of this kernel, all measured on 1M x 960, 10k queries: every thread fetching the
// 16-byte fragment pieces it multiplies, straight into registers one column group ahead (184 ms) or through a per-thread
// cp.async ring three groups ahead (176 ms) — ncu: every unit idle, the warps throttled on the memory-instruction queue: a
// warp instruction of that pattern asks for half of eight lines and the SM tracks a bounded number of pending lines; an L2
// bulk prefetch of 512-byte row blocks ahead of that ring (198 ms); a quarter warp copying one whole line, four lines per
// instruction, 16 columns per stage (91 ms; 12 warps instead of 8 changed nothing).  Here half a warp copies 256 contiguous
// bytes of a row, and the fragment owners read them back from shared memory.
constexpr int WD_ST = 3;                   // stages per warp: one or two in flight while one is multiplied
constexpr int WD_SCOLS = 32;               // columns per stage
constexpr int WD_RB = WD_SCOLS * 8;        // bytes of a row per stage
constexpr int WD_LINES = WD_SLAB + 16;     // rows + queries per stage
constexpr size_t WD_STAGE_BYTES = (size_t)WD_LINES * WD_RB;
constexpr size_t WD_SMEM = (size_t)WD_WARPS * WD_ST * WD_STAGE_BYTES;      // 216 KB: one CTA per SM

__device__ __forceinline__ void cp16_cg(unsigned dst, const void* src) {      // rows: L2 only
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp16_ca(unsigned dst, const void* src) {      // queries: re-read by every slab of the unit
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// Shared-memory layout of a stage: entry e (rows 0..31, then queries 0..15) at e * 256; its 16-byte chunk c (0..15) at
// position c ^ 4(e & 1).  Producer: half a warp writes the 16 chunks of one entry = 256 contiguous bytes.  Consumer: thread
// (g, t) reads chunk 4jj + t of entry 8mt + g; a quarter warp covers two neighbouring entries whose 64-byte pieces the
// swizzle puts on disjoint banks: LDS.128 without conflicts, no padding.
template <bool ANGULAR>
__global__ void __launch_bounds__(WD_WARPS * 32, 1)
k_score_wide(const double* __restrict__ X, int d, const double* __restrict__ Q, const UnitRec* __restrict__ units,
             const uint32_t* __restrict__ nunits_p, const int32_t* __restrict__ ids_sorted, Filter flt,
             unsigned long long* __restrict__ stat /* [0] units, [1] rows staged */) {
    extern __shared__ __align__(128) unsigned char wd_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int sub = lane >> 4, cc = lane & 15;           // producer role: entry 2i + sub of an instruction, chunk cc
    unsigned char* wring = wd_smem + (size_t)warp * (WD_ST * WD_STAGE_BYTES);
    const unsigned wring_s = (unsigned)__cvta_generic_to_shared(wring);
    // producer: byte offset of (entry 2i + sub, chunk cc) in a stage = i * 512 + p_off (2i is even: the entry's parity is sub)
    const unsigned p_off = (unsigned)sub * WD_RB + (unsigned)((cc ^ (4 * sub)) * 16);
    // consumer: byte offset of (entry 8mt + g, chunk 4jj + t) = mt * 2048 + c_off(jj)
    const unsigned c_base = (unsigned)g * WD_RB;
    const unsigned c_sw = 4u * (g & 1);
    const int64_t nunits = *nunits_p;
    const int64_t W = (int64_t)gridDim.x * WD_WARPS;
    const int nst = (d + WD_SCOLS - 1) / WD_SCOLS;       // stages per row
    SurvivorSink sink;
    unsigned long long rows_staged = 0, nmine = 0;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (int64_t u = (int64_t)blockIdx.x * WD_WARPS + warp; u < nunits; u += W) {
        const UnitRec* r = units + u;
        const uint32_t bstart = __ldg(&r->bstart);
        const int len = (int)__ldg(&r->len), m = (int)__ldg(&r->m);
        ++nmine;
        rows_staged += (unsigned)len;
        // producer: queries 2i + sub of the unit (slots >= m repeat the last query: copied, never kept)
        int qid[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) qid[i] = __ldg(&r->q[2 * i + sub]);
        // results of this thread: queries 8nb + 2t + e
        int c_q[2][2];
        double c_tau[2][2];
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = 8 * nb + 2 * t + e;
                c_q[nb][e] = __ldg(&r->q[j]);
                c_tau[nb][e] = j < m ? __ldg(flt.tau + c_q[nb][e]) : inf;          // +inf: masked
            }
        auto slabs = [&](auto two_blocks) {
            constexpr bool TWO = decltype(two_blocks)::value;
            for (int row0 = 0; row0 < len; row0 += WD_SLAB) {
                const int myid = __ldg(ids_sorted + bstart + min(row0 + lane, len - 1));   // rows >= len repeat the last one
                const int nrows = min(WD_SLAB, len - row0);
                int xid[16];                             // producer: rows 2i + sub of the slab
#pragma unroll
                for (int i = 0; i < 16; ++i) xid[i] = __shfl_sync(0xffffffffu, myid, 2 * i + sub);
                double acc[WD_MT][2][2];
                double xn[WD_MT], qq[2] = {0.0, 0.0};
#pragma unroll
                for (int mt = 0; mt < WD_MT; ++mt) {
                    xn[mt] = 0.0;
#pragma unroll
                    for (int nb = 0; nb < 2; ++nb) acc[mt][nb][0] = acc[mt][nb][1] = 0.0;
                }
                // stage k of the slab -> ring slot; joins the cp.async group committed next.  Chunks beyond column d (last
                // stage of a d that is not a multiple of 32) are written as zeros.
                auto request = [&](int k, int slot) {
                    if (k < nst) {
                        const unsigned dst = wring_s + (unsigned)slot * (unsigned)WD_STAGE_BYTES + p_off;
                        const int col = WD_SCOLS * k + 2 * cc;     // this lane's chunk: columns col, col + 1
                        if (col < d) {
#pragma unroll
                            for (int i = 0; i < 16; ++i)     // (entries of rows past the bucket's end keep what they held: masked)
                                if (2 * i + sub < nrows) cp16_cg(dst + i * 512u, X + (int64_t)xid[i] * d + col);
#pragma unroll
                            for (int i = 0; i < (TWO ? 8 : 4); ++i) cp16_ca(dst + (16 + i) * 512u, Q + (int64_t)qid[i] * d + col);
                        } else {
                            unsigned char* z = wring + (size_t)slot * WD_STAGE_BYTES + p_off;
#pragma unroll
                            for (int i = 0; i < 24; ++i) *reinterpret_cast<double2*>(z + i * 512) = make_double2(0.0, 0.0);
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                };
#pragma unroll
                for (int k = 0; k < WD_ST - 1; ++k) request(k, k);
                int cs = 0;                              // ring slot of stage k
                for (int k = 0; k < nst; ++k) {
                    asm volatile("cp.async.wait_group %0;" ::"n"(WD_ST - 2) : "memory");   // stage k has landed (this lane's part)
                    __syncwarp();                        // ... and every lane's; all lanes are done with stage k - 1
                    request(k + WD_ST - 1, cs == 0 ? WD_ST - 1 : cs - 1);
                    const unsigned char* src = wring + (size_t)cs * WD_STAGE_BYTES + c_base;
                    cs = cs + 1 == WD_ST ? 0 : cs + 1;
#pragma unroll
                    for (int jj = 0; jj < WD_SCOLS / 8; ++jj) {
                        const unsigned co = (((unsigned)(4 * jj + t)) ^ c_sw) * 16u;
                        double2 a[WD_MT], b[2];
#pragma unroll
                        for (int mt = 0; mt < WD_MT; ++mt) a[mt] = *reinterpret_cast<const double2*>(src + mt * (8 * WD_RB) + co);
                        b[0] = *reinterpret_cast<const double2*>(src + 4 * (8 * WD_RB) + co);
                        if (TWO) b[1] = *reinterpret_cast<const double2*>(src + 5 * (8 * WD_RB) + co);
#pragma unroll
                        for (int mt = 0; mt < WD_MT; ++mt) {
                            dmma884(acc[mt][0][0], acc[mt][0][1], a[mt].x, b[0].x);
                            if (TWO) dmma884(acc[mt][1][0], acc[mt][1][1], a[mt].x, b[1].x);
                        }
#pragma unroll
                        for (int mt = 0; mt < WD_MT; ++mt) {
                            dmma884(acc[mt][0][0], acc[mt][0][1], a[mt].y, b[0].y);
                            if (TWO) dmma884(acc[mt][1][0], acc[mt][1][1], a[mt].y, b[1].y);
                        }
                        if (ANGULAR) {
#pragma unroll
                            for (int mt = 0; mt < WD_MT; ++mt) xn[mt] = fma(a[mt].y, a[mt].y, fma(a[mt].x, a[mt].x, xn[mt]));
                            qq[0] = fma(b[0].y, b[0].y, fma(b[0].x, b[0].x, qq[0]));
                            if (TWO) qq[1] = fma(b[1].y, b[1].y, fma(b[1].x, b[1].x, qq[1]));
                        }
                    }
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");     // (only empty groups are left)
                __syncwarp();                            // the last stage is free before the next slab's first requests
                double xnr[WD_MT], c_qn[2][2] = {{1.0, 1.0}, {1.0, 1.0}};
                if (ANGULAR) {
#pragma unroll
                    for (int mt = 0; mt < WD_MT; ++mt) {
                        xn[mt] += __shfl_xor_sync(0xffffffffu, xn[mt], 1);
                        xn[mt] += __shfl_xor_sync(0xffffffffu, xn[mt], 2);
                        xnr[mt] = sqrt(xn[mt]);
                    }
#pragma unroll
                    for (int nb = 0; nb < 2; ++nb) {
                        double s = qq[nb];               // |query 8nb + g|^2: the 4 threads of the group
                        s += __shfl_xor_sync(0xffffffffu, s, 1);
                        s += __shfl_xor_sync(0xffffffffu, s, 2);
                        const double nrm = sqrt(s);
#pragma unroll
                        for (int e = 0; e < 2; ++e) c_qn[nb][e] = __shfl_sync(0xffffffffu, nrm, (2 * t + e) * 4);
                    }
                } else {
#pragma unroll
                    for (int mt = 0; mt < WD_MT; ++mt) xnr[mt] = 1.0;
                }
                // thread (g, t) holds (row row0 + 8mt + g, queries 8nb + 2t + e); one vote skips the slab's epilogue
                double v[WD_MT][2][2];
                bool keep[WD_MT][2][2];
                bool any = false;
#pragma unroll
                for (int mt = 0; mt < WD_MT; ++mt)
#pragma unroll
                    for (int nb = 0; nb < (TWO ? 2 : 1); ++nb)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            double s = acc[mt][nb][e];
                            if (ANGULAR) s = s / (c_qn[nb][e] * xnr[mt]);
                            v[mt][nb][e] = s;
                            keep[mt][nb][e] = row0 + 8 * mt + g < len && s >= c_tau[nb][e];
                            any |= keep[mt][nb][e];
                        }
                if (__any_sync(0xffffffffu, any)) {
#pragma unroll
                    for (int mt = 0; mt < WD_MT; ++mt)
#pragma unroll
                        for (int nb = 0; nb < (TWO ? 2 : 1); ++nb)
#pragma unroll
                            for (int e = 0; e < 2; ++e)
                                sink.push(flt, keep[mt][nb][e], c_q[nb][e], bstart + (uint32_t)(row0 + 8 * mt + g), v[mt][nb][e], lane);
                }
            }
        };
        if (m > 8) slabs(std::true_type{});
        else slabs(std::false_type{});
    }
    sink.flush(flt, lane);
    if (lane == 0 && nmine) { atomicAdd(&stat[0], nmine); atomicAdd(&stat[1], rows_staged); }
}

// ---------------------------------------------------------------------------------------------------------
// k_threshold_wide — see k_threshold (rerank_bm.cu): the same samples, lists and lower bounds for any (even) d
// ---------------------------------------------------------------------------------------------------------
constexpr int TW_ROWS = 4;                 // rows scored at a time by one warp
constexpr int TW_UNROLL = 4;               // 16-byte chunks per lane, row and round trip

template <bool ANGULAR, bool REG /* K <= 32: the sample list lives in registers */>
__global__ void __launch_bounds__(RR_THREADS)
k_threshold_wide(const double* __restrict__ X, int d, ChunkView cv, int L, int NT, const uint32_t* __restrict__ leaf_pos,
                 const int32_t* __restrict__ leaf_len, const int32_t* __restrict__ ids_sorted, int self_exclude, int K,
                 double* __restrict__ tl_keys, int* __restrict__ tl_ids, int* __restrict__ tl_cnt) {
    extern __shared__ double rsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* mykeys = rsm + (size_t)warp * K;
    int* myids = reinterpret_cast<int*>(rsm + (size_t)RR_WARPS * K) + (size_t)warp * K;
    const int64_t wid = (int64_t)blockIdx.x * RR_WARPS + warp;
    if (wid >= cv.nqc * NT) return;
    const int64_t ql = wid / NT;
    const int sample = (int)(wid % NT);
    const int qid = cv.qids ? cv.qids[ql] : INT32_MIN;
    const bool excl = self_exclude && cv.qids && qid >= -128 && qid <= 127;
    const double* __restrict__ qrow = cv.Q + ql * d;
    const int nch = d >> 1;                  // 16-byte chunks of a row (d even, rows 16-byte aligned)
    double qn = 1.0;
    if (ANGULAR) {
        double qq = 0.0;
        for (int c = lane; c < nch; c += 32) {
            const double2 q2 = ldg_d2(qrow + 2 * c);
            qq = fma(q2.y, q2.y, fma(q2.x, q2.x, qq));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
        qn = sqrt(qq);
    }
    const double slack = 4.0 * (double)(d + 8) * 1.1102230246251565e-16;     // >= twice the bound, with room for the norms
    // the sample-th table (among the first 32) in which the query probes something
    uint32_t nonempty = __ballot_sync(0xffffffffu, lane < L && cv.pair_cnt[ql * L + min(lane, L - 1)] > 0u);
    for (int i = 0; i < sample && nonempty; ++i) nonempty &= nonempty - 1;
    int count = 0;
    double kth = 0.0;
    RegList rl;
    const int tb = nonempty ? __ffs(nonempty) - 1 : 0;
    const uint32_t nb_ = nonempty ? cv.pair_cnt[ql * L + tb] : 0u;
    int seen = 0;
    // the first bucket the query probes in that table (in practice its own) and, should it hold fewer than K rows, the next
    // ones until K rows have been seen
    for (uint32_t e_ = 0; e_ < nb_ && (e_ == 0 || seen < K); ++e_) {
        const uint32_t leaf = cv.cache[(ql * L + tb) * cv.cap + e_];
        const int32_t* bids = ids_sorted + leaf_pos[leaf];
        const int len = leaf_len[leaf];
        seen += len;
        for (int row0 = 0; row0 < len; row0 += TW_ROWS) {
            int id[TW_ROWS];
            const double* xr[TW_ROWS];
#pragma unroll
            for (int r = 0; r < TW_ROWS; ++r) {
                id[r] = __ldg(bids + min(row0 + r, len - 1));
                xr[r] = X + (int64_t)id[r] * d;
            }
            double dot[TW_ROWS], ab[TW_ROWS], xx[TW_ROWS];
#pragma unroll
            for (int r = 0; r < TW_ROWS; ++r) dot[r] = ab[r] = xx[r] = 0.0;
            // TW_UNROLL chunks x TW_ROWS rows = 16 independent 16-byte loads per lane and round trip (8 KB per warp in flight;
            // two chunks at a time left the kernel waiting on memory latency at 2.4 TB/s); chunks beyond the row count as 0
            for (int c0 = lane; c0 < nch; c0 += 32 * TW_UNROLL) {
                double2 q2[TW_UNROLL], x2[TW_UNROLL][TW_ROWS];
#pragma unroll
                for (int u = 0; u < TW_UNROLL; ++u) {
                    const int c = c0 + 32 * u;
                    const bool ok = c < nch;
                    const int cc = ok ? c : c0;
                    q2[u] = ldg_d2(qrow + 2 * cc);
                    if (!ok) q2[u] = make_double2(0.0, 0.0);
#pragma unroll
                    for (int r = 0; r < TW_ROWS; ++r) x2[u][r] = ldg_d2_stream(xr[r] + 2 * cc);
                }
#pragma unroll
                for (int u = 0; u < TW_UNROLL; ++u)
#pragma unroll
                    for (int r = 0; r < TW_ROWS; ++r) {
                        dot[r] = fma(x2[u][r].y, q2[u].y, fma(x2[u][r].x, q2[u].x, dot[r]));
                        ab[r] = fma(fabs(x2[u][r].y), fabs(q2[u].y), fma(fabs(x2[u][r].x), fabs(q2[u].x), ab[r]));
                        if (ANGULAR && c0 + 32 * u < nch) xx[r] = fma(x2[u][r].y, x2[u][r].y, fma(x2[u][r].x, x2[u][r].x, xx[r]));
                    }
            }
#pragma unroll
            for (int r = 0; r < TW_ROWS; ++r)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    dot[r] += __shfl_xor_sync(0xffffffffu, dot[r], o);
                    ab[r] += __shfl_xor_sync(0xffffffffu, ab[r], o);
                    if (ANGULAR) xx[r] += __shfl_xor_sync(0xffffffffu, xx[r], o);
                }
#pragma unroll
            for (int r = 0; r < TW_ROWS; ++r) {          // every lane holds the sums: the insertions are warp-uniform
                if (row0 + r >= len) break;
                double lb;                               // lower bound of the score the scoring kernel will compute
                if (ANGULAR) {
                    const double den = qn * sqrt(xx[r]);
                    lb = dot[r] / den - slack * (ab[r] / den);
                } else {
                    lb = dot[r] - slack * ab[r];
                }
                if (!(lb == lb) || (excl && id[r] == qid)) continue;
                if (count == K && lb < kth) continue;
                if (REG) {                               // the rows of one table are distinct: no duplicate check
                    rl.insert(K, lb, id[r], lane);
                    count = rl.count;
                    if (count == K) kth = __shfl_sync(0xffffffffu, rl.key, K - 1);
                } else {
                    if (count == K && !better(lb, id[r], mykeys[K - 1], myids[K - 1])) continue;
                    warp_insert(mykeys, myids, count, K, lb, id[r], lane);
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

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
// FP64 rows of any even width, dot product or cosine (squared L2 on real-valued data cancels: row-major path)
bool wide_supported(const dpf_index* h, int metric) {
    if (h->dbg[DPF_DBG_WIDE] == 1) return false;               // test hook: keep d > 128 on the row-major kernel
    return h->cfg.d > BM_KC && (metric == DPF_METRIC_DOT || metric == DPF_METRIC_ANGULAR);
}

void launch_threshold_wide(dpf_index* h, cudaStream_t st, int metric, const ChunkView& cv, int NT, int topk, size_t list_smem) {
    const unsigned grid = (unsigned)((cv.nqc * NT + RR_WARPS - 1) / RR_WARPS);
    auto go = [&](auto kern) {
        kern<<<grid, RR_THREADS, list_smem, st>>>(h->Xdev, h->cfg.d, cv, h->cfg.L, NT, h->leaf_pos.p, h->leaf_len.p, h->ids_sorted.p,
                                                  h->cfg.self_exclude_small_ids, topk, h->bm_tl_keys.p, h->bm_tl_ids.p, h->bm_tl_cnt.p);
        DPF_LAUNCHED();
    };
    const bool ang = metric == DPF_METRIC_ANGULAR;
    if (topk <= 32) { if (ang) go(k_threshold_wide<true, true>); else go(k_threshold_wide<false, true>); }
    else { if (ang) go(k_threshold_wide<true, false>); else go(k_threshold_wide<false, false>); }
}

void launch_score_wide(dpf_index* h, const ChunkView& cv, const void* units, const uint32_t* nunits_p, int metric, const Filter& flt,
                       unsigned long long* bm_stat) {
    auto go = [&](auto kern) {
        DPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WD_SMEM));
        kern<<<h->num_sms, WD_WARPS * 32, WD_SMEM, h->stream>>>(h->Xdev, h->cfg.d, cv.Q, reinterpret_cast<const UnitRec*>(units), nunits_p,
                                                              h->ids_sorted.p, flt, bm_stat);
        DPF_LAUNCHED();
    };
    if (metric == DPF_METRIC_ANGULAR) go(k_score_wide<true>); else go(k_score_wide<false>);
}

}  // namespace dpf
