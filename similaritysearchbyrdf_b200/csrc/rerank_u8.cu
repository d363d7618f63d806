// rerank_u8.cu — K5c: bucket-major scoring from the uint8 compact store (store.cu).
//
// Same job as k_score_stream (rerank_bm.cu): for every unit (one leaf bucket x <= 16 queries that probe it) one
// score per (query, bucket row), replacing the gather + dgemv of topKAndPrecisionScore
// (src/main/scala/mclab/deploy/DensevectorRDFInit.scala:472-507) — but only scores that can still be among the query's
// best k leave the kernel (threshold filter, rerank_units.cuh).  When the store is bytes a row is <= 128 B, i.e. two
// 16-byte chunks per thread of a tensor-pipe row group, and the per-row issue cost of the TMA pipeline, which bounds
// k_score_stream once the bytes are cheap, can go.  Kernels:
//   k_score_u8s   every query of the batch is itself a vector of bytes (checked per batch, k_quantise_queries): all
//                 products and partial sums are integers below 2^31, so the integer tensor pipe (IMMA.16x8x32, u8 x u8 ->
//                 s32) computes the exact dot product and (double)dot is bit-identical to the FP64 sum in any order.
//                 This is the SIFT case (descriptors and queries are bytes).  Rows and unit records stream through
//                 shared memory with cp.async; dot, angular and (exactly, as integers) squared L2.
//   k_score_u8i   the same arithmetic with the rows loaded into registers, kept for comparison (DPF_U8I_KERNEL=lean).
//   k_score_u8d   queries are arbitrary doubles: bytes are widened to the identical doubles in registers (PRMT + DADD)
//                 and multiplied on the FP64 tensor pipe (DMMA.8x8x4) against the FP64 query fragments.
//   k_threshold_u8i   the threshold samples for byte queries, also on the integer tensor pipe.
// k permutation (all): thread t of a row group holds bytes [16t, 16t+16) and [64+16t, 64+16t+16) of its row; the
// query operand uses the same columns, so the sum over k is unchanged.
#include <cstdlib>
#include <type_traits>

#include "rerank_units.cuh"

namespace dpf {

constexpr int U8_WARPS = 8;
constexpr int U8_INT_CTAS = 3;          // CTAs per SM of the integer variant (<= 85 registers): latency is hidden by occupancy
constexpr int U8_QPITCH = 128;          // row pitch of the byte copy of the queries (zero padded)

// ---------------------------------------------------------------------------------------------------------
// queries -> bytes (when they are bytes)
// ---------------------------------------------------------------------------------------------------------
// warp per query: Q8[q][c] = (uint8)Q[q][c] (0 for the pad columns), qsq[q] = sum of squares (exact integer), qnorm[q] =
// its square root; *flag |= 1 if some value is not a byte
__global__ void __launch_bounds__(256)
k_quantise_queries(const double* __restrict__ Q, int64_t nq, int d, int pitch, unsigned char* __restrict__ Q8,
                   double* __restrict__ qnorm, int32_t* __restrict__ qsq, int* __restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const double* x = Q + q * d;
    unsigned long long sq = 0;
    int bad = 0;
    for (int c = lane; c < pitch; c += 32) {
        unsigned char b = 0;
        if (c < d) {
            const double v = x[c];
            b = (unsigned char)min(max(v, 0.0), 255.0);
            if (__double_as_longlong((double)b) != __double_as_longlong(v)) bad = 1;
            sq += (unsigned)b * (unsigned)b;
        }
        Q8[q * pitch + c] = b;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        qnorm[q] = sqrt((double)sq);
        qsq[q] = (int32_t)sq;                          // < 2^31 for d * 255^2 < 2^31 (checked by the caller)
        if (bad) atomicOr(flag, 1);
    }
}

__device__ __forceinline__ void imma_u8(int (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// byte b of w as a double: 2^52 + v has mantissa v; minus 2^52 is exact
template <int B>
__device__ __forceinline__ double byte_to_double(unsigned w) {
    return __hiloint2double(0x43300000, (int)__byte_perm(w, 0u, 0x4440u | (unsigned)B)) - 4503599627370496.0;
}

__device__ __forceinline__ uint4 ldg_u4(const unsigned char* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// ---------------------------------------------------------------------------------------------------------
// k_score_u8d — byte rows x FP64 queries on the FP64 tensor pipe
// ---------------------------------------------------------------------------------------------------------
template <bool ANGULAR>
__global__ void __launch_bounds__(U8_WARPS * 32, 1)
k_score_u8d(const unsigned char* __restrict__ X8, unsigned pitch /* row bytes: multiple of 16, <= 128 */, int d,
            const double* __restrict__ Q, const int* __restrict__ q8_bad, int gate, const UnitRec* __restrict__ units,
            const uint32_t* __restrict__ nunits_p, const int32_t* __restrict__ ids_sorted, Filter flt,
            unsigned long long* __restrict__ stat /* [0] units, [1] rows staged */) {
    if (gate && *q8_bad == 0) return;            // the batch is byte vectors: the integer kernel scores it
    constexpr int TR = 8;                        // rows per tile = DMMA M
    constexpr int TPW = 32 / TR;                 // tiles per id window
    constexpr int PF = 3;                        // tiles in flight per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const bool has0 = 16u * t < pitch, has1 = 64u + 16u * t < pitch;   // this thread's chunks exist
    const bool two_chunks = pitch > 64;                              // warp-uniform
    // every row / query load below is unconditional and its address clamped into valid memory: a select on a loaded
    // value makes the compiler wait for the load where it is issued (measured: 40 % of the kernel), whereas a garbage
    // operand is harmless as long as the other operand of the product is 0
    const unsigned off0 = has0 ? 16u * t : 0u, off1 = has1 ? 64u + 16u * t : 0u;
    const int64_t nunits = *nunits_p;
    const int64_t W = (int64_t)gridDim.x * U8_WARPS;
    const int64_t gw = (int64_t)blockIdx.x * U8_WARPS + warp;
    const int64_t nmine = nunits > gw ? (nunits - gw + W - 1) / W : 0;   // units gw + k * W
    if (nmine == 0) return;
    unsigned long long rows_staged = 0;
    SurvivorSink sink;

    // record of the next unit, loaded one unit ahead: lane j < 16 keeps query j, every lane one first-window id
    uint32_t n_bstart, n_len, n_m;
    int n_q, n_id;
    auto fetch_rec = [&](int64_t k) {
        const UnitRec* r = units + (gw + k * W);
        n_bstart = __ldg(&r->bstart); n_len = __ldg(&r->len); n_m = __ldg(&r->m);
        n_q = __ldg(&r->q[lane & (SS_UQ - 1)]);
        n_id = __ldg(&r->ids0[lane]);
    };
    fetch_rec(0);

    for (int64_t k = 0; k < nmine; ++k) {
        const uint32_t bstart = n_bstart;
        const int len = (int)n_len, m = (int)n_m;
        const int my_q = n_q, id_first = n_id;
        if (k + 1 < nmine) fetch_rec(k + 1);
        rows_staged += (unsigned)len;
        const int nb_used = (m + 7) >> 3;

        // ---- query operand: k-step c * 16 + e <-> column 64c + 16t + e -----------------------------------------
        double B[2][32];
        double c_qn[2][2] = {{1.0, 1.0}, {1.0, 1.0}}, c_tau[2][2];
        int c_q[2][2];
        bool c_ok[2][2];
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            const int qi = __shfl_sync(0xffffffffu, my_q, 8 * nb + g);      // slots >= m repeat the last query
            const double* qp = Q + (int64_t)qi * d;
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    const int col = 64 * c + 16 * t + e;          // even; d is even
                    const double2 v = __ldg(reinterpret_cast<const double2*>(qp + min(col, d - 2)));
                    B[nb][16 * c + e] = v.x;
                    B[nb][16 * c + e + 1] = v.y;
                }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = 8 * nb + 2 * t + e;
                c_q[nb][e] = __shfl_sync(0xffffffffu, my_q, j);
                c_ok[nb][e] = j < m;
                c_tau[nb][e] = __ldg(flt.tau + c_q[nb][e]);
            }
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if (64 * c + 16 * t + e >= d) B[nb][16 * c + e] = 0.0;      // columns >= d: the row operand is garbage there
            if (ANGULAR) {
                double sq = 0.0;             // ||query 8nb + g||^2: this thread's columns, then the 4 threads of the group
#pragma unroll
                for (int w = 0; w < 32; ++w) sq = fma(B[nb][w], B[nb][w], sq);
                sq += __shfl_xor_sync(0xffffffffu, sq, 1);
                sq += __shfl_xor_sync(0xffffffffu, sq, 2);
                const double nrm = sqrt(sq);
#pragma unroll
                for (int e = 0; e < 2; ++e) c_qn[nb][e] = __shfl_sync(0xffffffffu, nrm, (2 * t + e) * 4);
            }
        }

        // ---- rows: ids by 32-row windows (one window ahead), tiles of 8 rows, PF tiles in flight -------------------
        const int ntiles = (len + TR - 1) / TR;
        const int32_t* bids = ids_sorted + bstart;
        int win_lo = 0;                      // idA = ids of window win_lo, idB = window win_lo + 1
        int idA = id_first;
        int idB = 32 < len ? __ldg(bids + min(32 + lane, len - 1)) : 0;
        uint4 a[PF][2];
        auto load_tile = [&](uint4 (&dst)[2], int tile) {
            const int wi = tile / TPW;
            if (wi > win_lo) {               // warp-uniform; tiles are loaded in increasing order
                idA = idB;
                win_lo = wi;
                idB = 32 * (wi + 1) < len ? __ldg(bids + min(32 * (wi + 1) + lane, len - 1)) : 0;
            }
            const int row = min(tile * TR + g, len - 1);
            const int id = __shfl_sync(0xffffffffu, idA, row - 32 * wi);
            const unsigned char* xp = X8 + (size_t)id * pitch;
            dst[0] = ldg_u4(xp + off0);      // a chunk this thread does not have reads chunk 0 instead: its query
            dst[1] = ldg_u4(xp + off1);      // operand is 0 there
        };
        auto compute_tile = [&](const uint4 (&src)[2], int tile) {
            double acc[2][2][2];            // [n-block][even / odd k-step chain][column]
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                for (int c = 0; c < 2; ++c) acc[nb][c][0] = acc[nb][c][1] = 0.0;
            double xn = 0.0;
            auto steps = [&](auto chunk, auto two_blocks) {
                constexpr int C = decltype(chunk)::value;
                auto word = [&](auto wi, unsigned w) {
                    constexpr int WI = decltype(wi)::value;
                    const double a0 = byte_to_double<0>(w), a1 = byte_to_double<1>(w);
                    const double a2 = byte_to_double<2>(w), a3 = byte_to_double<3>(w);
                    constexpr int S = 16 * C + 4 * WI;
                    dmma884(acc[0][0][0], acc[0][0][1], a0, B[0][S]);
                    if (decltype(two_blocks)::value) dmma884(acc[1][0][0], acc[1][0][1], a0, B[1][S]);
                    dmma884(acc[0][1][0], acc[0][1][1], a1, B[0][S + 1]);
                    if (decltype(two_blocks)::value) dmma884(acc[1][1][0], acc[1][1][1], a1, B[1][S + 1]);
                    dmma884(acc[0][0][0], acc[0][0][1], a2, B[0][S + 2]);
                    if (decltype(two_blocks)::value) dmma884(acc[1][0][0], acc[1][0][1], a2, B[1][S + 2]);
                    dmma884(acc[0][1][0], acc[0][1][1], a3, B[0][S + 3]);
                    if (decltype(two_blocks)::value) dmma884(acc[1][1][0], acc[1][1][1], a3, B[1][S + 3]);
                    if (ANGULAR && (C == 0 ? has0 : has1)) { xn = fma(a0, a0, xn); xn = fma(a1, a1, xn); xn = fma(a2, a2, xn); xn = fma(a3, a3, xn); }
                };
                word(std::integral_constant<int, 0>{}, src[C].x);
                word(std::integral_constant<int, 1>{}, src[C].y);
                word(std::integral_constant<int, 2>{}, src[C].z);
                word(std::integral_constant<int, 3>{}, src[C].w);
            };
            if (nb_used == 2) {
                steps(std::integral_constant<int, 0>{}, std::true_type{});
                if (two_chunks) steps(std::integral_constant<int, 1>{}, std::true_type{});
            } else {
                steps(std::integral_constant<int, 0>{}, std::false_type{});
                if (two_chunks) steps(std::integral_constant<int, 1>{}, std::false_type{});
            }
            double xnr = 1.0;
            if (ANGULAR) {
                xn += __shfl_xor_sync(0xffffffffu, xn, 1);
                xn += __shfl_xor_sync(0xffffffffu, xn, 2);
                xnr = sqrt(xn);
            }
            // thread (g, t) holds (row tile * 8 + g, queries 8nb + 2t, 8nb + 2t + 1)
            const int row = tile * TR + g;
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double v = acc[nb][0][e] + acc[nb][1][e];
                    if (ANGULAR) v = v / (c_qn[nb][e] * xnr);
                    sink.push(flt, row < len && nb < nb_used && c_ok[nb][e] && v >= c_tau[nb][e], c_q[nb][e], bstart + (uint32_t)row, v, lane);
                }
        };

#pragma unroll
        for (int p = 0; p < PF; ++p)
            if (p < ntiles) load_tile(a[p], p);
        for (int base = 0; base < ntiles; base += PF) {
#pragma unroll
            for (int p = 0; p < PF; ++p) {
                const int tile = base + p;
                if (tile < ntiles) {             // warp-uniform
                    compute_tile(a[p], tile);
                    if (tile + PF < ntiles) load_tile(a[p], tile + PF);
                }
            }
        }
    }
    sink.flush(flt, lane);
    if (lane == 0) { atomicAdd(&stat[0], (unsigned long long)nmine); atomicAdd(&stat[1], rows_staged); }
}

// ---------------------------------------------------------------------------------------------------------
// k_score_u8i — byte rows x byte queries on the integer tensor pipe.  The arithmetic is trivial and the loop is a
// chain of dependent loads (record -> ids -> rows), so this kernel is kept lean (<= 85 registers, 24 warps per SM)
// and hides the latency with occupancy instead of a deep per-warp pipeline.  IMMA.16x8x32: 16 rows per tile, thread
// (g, t) loads its two chunks of rows g and 8 + g and holds (rows g, 8 + g) x (queries 2t, 2t + 1) per n-block.
// ---------------------------------------------------------------------------------------------------------
template <bool ANGULAR>
__global__ void __launch_bounds__(U8_WARPS * 32, ANGULAR ? U8_INT_CTAS - 1 : U8_INT_CTAS)
k_score_u8i(const unsigned char* __restrict__ X8, unsigned pitch, const unsigned char* __restrict__ Q8,
            const double* __restrict__ qnorm, const int* __restrict__ q8_bad, const UnitRec* __restrict__ units,
            const uint32_t* __restrict__ nunits_p, const int32_t* __restrict__ ids_sorted, Filter flt,
            unsigned long long* __restrict__ stat) {
    if (*q8_bad != 0) return;                    // some query is not a byte vector: k_score_u8d scores the batch
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const bool has0 = 16u * t < pitch, has1 = 64u + 16u * t < pitch;
    const bool two_chunks = pitch > 64;
    const unsigned off0 = has0 ? 16u * t : 0u, off1 = has1 ? 64u + 16u * t : 0u;   // clamped, see k_score_u8d
    const uint32_t nunits = *nunits_p;
    const uint32_t W = gridDim.x * U8_WARPS;
    unsigned rows_staged = 0, nmine = 0;
    SurvivorSink sink;
    for (uint32_t u = blockIdx.x * U8_WARPS + warp; u < nunits; u += W) {
        const UnitRec* r = units + u;
        const uint32_t bstart = __ldg(&r->bstart);
        const int len = (int)__ldg(&r->len), m = (int)__ldg(&r->m);
        const int my_q = __ldg(&r->q[lane & (SS_UQ - 1)]);
        int idA = __ldg(&r->ids0[lane]);
        rows_staged += (unsigned)len;
        nmine++;
        const bool two_blocks = m > 8;
        // queries g and 8 + g (slots >= m repeat the last query: harmless, their scores are masked)
        uint4 bq[2][2];
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            const unsigned char* qp = Q8 + (size_t)__shfl_sync(0xffffffffu, my_q, 8 * nb + g) * U8_QPITCH + 16 * t;   // zero beyond column d
            bq[nb][0] = ldg_u4(qp);
            bq[nb][1] = ldg_u4(qp + 64);
        }
        // thresholds of queries 8nb + 2t + e; dot: the score is an integer, compare integers
        int c_q[2][2];
        double c_tau[2][2];
        int c_taui[2][2];
        double c_qn[2][2];
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = 8 * nb + 2 * t + e;
                c_q[nb][e] = __shfl_sync(0xffffffffu, my_q, j);
                const double tv = j < m ? __ldg(flt.tau + c_q[nb][e]) : __longlong_as_double(0x7ff0000000000000LL);   // +inf: masked
                if (ANGULAR) { c_tau[nb][e] = tv; c_qn[nb][e] = __ldg(qnorm + c_q[nb][e]); }
                // a dot product of bytes is below 2^31 - 1: INT_MAX masks the slot
                else c_taui[nb][e] = tv >= 2147483647.0 ? 0x7fffffff : (tv <= -2147483648.0 ? (int)0x80000000 : (int)ceil(tv));
            }
        const int32_t* bids = ids_sorted + bstart;
        for (int row0 = 0; row0 < len; row0 += 32) {                 // one id window = 2 tiles of 16 rows
            if (row0) idA = __ldg(bids + min(row0 + lane, len - 1));
            uint4 a[2][2][2];                                        // [tile][row g / 8 + g][chunk]
#pragma unroll
            for (int tl = 0; tl < 2; ++tl)
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int row = min(row0 + 16 * tl + 8 * rr + g, len - 1);
                    const unsigned char* xp = X8 + (size_t)__shfl_sync(0xffffffffu, idA, row - row0) * pitch;
                    a[tl][rr][0] = ldg_u4(xp + off0);
                    a[tl][rr][1] = ldg_u4(xp + off1);
                }
#pragma unroll
            for (int tl = 0; tl < 2; ++tl) {
                if (row0 + 16 * tl >= len) break;                    // warp-uniform
                int acc[2][4];
#pragma unroll
                for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[nb][i] = 0;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c == 1 && !two_chunks) break;
                    imma_u8(acc[0], a[tl][0][c].x, a[tl][1][c].x, a[tl][0][c].y, a[tl][1][c].y, bq[0][c].x, bq[0][c].y);
                    imma_u8(acc[0], a[tl][0][c].z, a[tl][1][c].z, a[tl][0][c].w, a[tl][1][c].w, bq[0][c].z, bq[0][c].w);
                    if (two_blocks) {
                        imma_u8(acc[1], a[tl][0][c].x, a[tl][1][c].x, a[tl][0][c].y, a[tl][1][c].y, bq[1][c].x, bq[1][c].y);
                        imma_u8(acc[1], a[tl][0][c].z, a[tl][1][c].z, a[tl][0][c].w, a[tl][1][c].w, bq[1][c].z, bq[1][c].w);
                    }
                }
                double xnr[2] = {1.0, 1.0};
                if (ANGULAR) {
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        unsigned s = 0;
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            unsigned sc = 0;
                            sc = __dp4a(a[tl][rr][c].x, a[tl][rr][c].x, sc); sc = __dp4a(a[tl][rr][c].y, a[tl][rr][c].y, sc);
                            sc = __dp4a(a[tl][rr][c].z, a[tl][rr][c].z, sc); sc = __dp4a(a[tl][rr][c].w, a[tl][rr][c].w, sc);
                            if (c == 0 ? has0 : has1) s += sc;
                        }
                        s += __shfl_xor_sync(0xffffffffu, s, 1);
                        s += __shfl_xor_sync(0xffffffffu, s, 2);
                        xnr[rr] = sqrt((double)s);
                    }
                }
                // c0, c1 = (row g, queries 2t, 2t + 1), c2, c3 = (row 8 + g, same queries)
#pragma unroll
                for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int row = row0 + 16 * tl + 8 * rr + g;
                            const int dot = acc[nb][2 * rr + e];
                            bool keep;
                            double v = (double)dot;
                            if (ANGULAR) { v = v / (c_qn[nb][e] * xnr[rr]); keep = v >= c_tau[nb][e]; }
                            else keep = dot >= c_taui[nb][e];
                            sink.push(flt, keep && row < len && (nb == 0 || two_blocks), c_q[nb][e], bstart + (uint32_t)row, v, lane);
                        }
            }
        }
    }
    sink.flush(flt, lane);
    if (lane == 0) { atomicAdd(&stat[0], (unsigned long long)nmine); atomicAdd(&stat[1], (unsigned long long)rows_staged); }
}

// ---------------------------------------------------------------------------------------------------------
// k_score_u8s — the same computation as k_score_u8i with rows and unit records streamed through shared memory.
// k_score_u8i keeps one id window (32 rows) of loads in flight per warp and waits for it before it multiplies: with
// ~2 us from request to data it is latency-bound even at 24 warps per SM (ncu: 40 % of the samples on the first use of
// a loaded row).  Here every thread copies its own 16-byte chunks with cp.async into a per-warp ring of US_D tiles (no
// registers held while the data is in flight; a thread only reads back what it copied itself) and the ring keeps
// running across unit boundaries: the producer side of the warp is US_D - 1 tiles — often several units — ahead of its
// consumer side.  The unit records are copied the same way, US_D units ahead of the producer, into a ring of US_R
// records that both sides read; id windows beyond the first (which rides in the record) are requested three windows
// ahead, the next unit's query operand and thresholds one unit ahead.
// cp.async groups: iteration s commits one group = tile s (+ possibly one record); the wait before tile c is consumed
// leaves the US_D - 1 younger groups pending, so at the top of iteration s everything committed up to s - US_D is
// complete — a record requested US_D production steps before its unit is opened has arrived.
// ---------------------------------------------------------------------------------------------------------
constexpr int US_D = 5;                // tile ring depth, tiles of 16 rows (2 KB each)
constexpr int US_R = 2 * US_D;         // record ring depth: records of units [consumer's, producer's + US_D] are live
constexpr int US_REC_CHUNKS = sizeof(UnitRec) / 16;
constexpr int US_CTAS = 2;             // CTAs of 8 warps per SM (<= 128 registers)
constexpr size_t US_WARP_SMEM = (size_t)US_D * 4 * 32 * 16 + (size_t)US_R * sizeof(UnitRec);
constexpr size_t US_SMEM = U8_WARPS * US_WARP_SMEM;
static_assert(US_CTAS * US_SMEM <= 227 * 1024, "two CTAs per SM");
static_assert(US_REC_CHUNKS <= 32, "one lane per 16-byte chunk of a record");

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

template <int METRIC>
__global__ void __launch_bounds__(U8_WARPS * 32, US_CTAS)
k_score_u8s(const unsigned char* __restrict__ X8, unsigned pitch, const unsigned char* __restrict__ Q8,
            const double* __restrict__ qnorm, const int32_t* __restrict__ qsq, const int* __restrict__ q8_bad,
            const UnitRec* __restrict__ units, const uint32_t* __restrict__ nunits_p, const int32_t* __restrict__ ids_sorted,
            Filter flt, unsigned long long* __restrict__ stat) {
    if (*q8_bad != 0) return;                    // some query is not a byte vector: k_score_u8d scores the batch
    // DOT: key = dot.  ANGULAR: key = dot / (|q| |x|).  L2: key = -|q - x|^2 = 2 dot - |x|^2 - |q|^2, an exact integer.
    constexpr bool ANGULAR = METRIC == DPF_METRIC_ANGULAR, L2 = METRIC == DPF_METRIC_L2;
    extern __shared__ __align__(16) unsigned char us_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    unsigned char* wbase = us_smem + (size_t)warp * US_WARP_SMEM;
    uint4* ring = reinterpret_cast<uint4*>(wbase) + lane;                 // slot s, vector j: ring[(4 s + j) * 32]
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);      // the same as a shared-window address
    UnitRec* recs = reinterpret_cast<UnitRec*>(wbase + (size_t)US_D * 4 * 32 * 16);
    const bool has0 = 16u * t < pitch, has1 = 64u + 16u * t < pitch;
    const bool two_chunks = pitch > 64;
    const unsigned off0 = has0 ? 16u * t : 0u, off1 = has1 ? 64u + 16u * t : 0u;   // clamped, see k_score_u8d
    unsigned long long xb0 = (unsigned long long)(X8 + off0), xb1 = (unsigned long long)(X8 + off1);
    asm volatile("" : "+l"(xb0), "+l"(xb1));     // kept as two 64-bit registers (the compiler would re-add base + offset per address)
    // (Tried: a quarter warp copying one whole 128-byte row per instruction instead of this fragment pattern — the same
    // 1.17 ms: both halves of a row are requested back to back here and the kernel moves 6.4 TB/s from L2 to the SMs (ncu
    // l1tex__m_xbar2l1tex_read_bytes), the rate of the measured device copy; 17 % fewer instructions did not change it either.)
    // (unit counts are 32-bit: the record array of a chunk holds < 2^31 units; 32-bit counters keep the ring bookkeeping short)
    const int nunits = (int)*nunits_p;
    const int W = (int)gridDim.x * U8_WARPS;
    const int u0 = (int)blockIdx.x * U8_WARPS + warp;
    const int nmine = nunits > u0 ? (nunits - u0 + W - 1) / W : 0;         // this warp's units: u0 + k W, k < nmine
    if (nmine == 0) return;
    unsigned rows_staged = 0;
    SurvivorSink sink;

    // record k -> ring slot k % US_R (joins the cp.async group that is committed next)
    auto rec_request = [&](int k, int rslot) {
        if (k < nmine && lane < US_REC_CHUNKS)
            cp_async16(reinterpret_cast<unsigned char*>(&recs[rslot]) + 16 * lane,
                       reinterpret_cast<const unsigned char*>(units + (u0 + k * W)) + 16 * lane);
    };
    for (int k = 0; k <= US_D; ++k) rec_request(k, k);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    // ---- producer side: row copies, US_D - 1 tiles ahead -----------------------------------------------------------
    auto win_load = [&](const UnitRec* r, int w) {       // ids of rows [32w, 32w + 32) of the record's bucket, 0 past the end
        const int len = (int)r->len;
        return 32 * w < len ? __ldg(ids_sorted + r->bstart + min(32 * w + lane, len - 1)) : 0;
    };
    int pk = 0;                        // unit being produced
    int p_rslot = 0;                   // = pk % US_R; the record requested when pk is opened goes to (p_rslot + US_D) % US_R
    int p_tile = 0, p_len = (int)recs[0].len;
    int p_w0 = recs[0].ids0[lane], p_w1 = win_load(&recs[0], 1), p_w2 = win_load(&recs[0], 2), p_w3 = win_load(&recs[0], 3);
    int nx_w1 = 0, nx_w2 = 0, nx_w3 = 0;                  // windows 1..3 of unit pk + 1
    if (nmine > 1) { nx_w1 = win_load(&recs[1], 1); nx_w2 = win_load(&recs[1], 2); nx_w3 = win_load(&recs[1], 3); }
    auto issue = [&](int slot) {
        if (pk < nmine) {
            const unsigned dst = ring_s + (unsigned)slot * (4 * 32 * 16);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int row = min(16 * p_tile + 8 * rr + g, p_len - 1);
                // (the chunk offsets are folded into the two base pointers: one multiply-add per address)
                const unsigned long long ro = (unsigned long long)(unsigned)__shfl_sync(0xffffffffu, p_w0, row & 31) * pitch;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (2 * rr) * 512), "l"(xb0 + ro) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (2 * rr + 1) * 512), "l"(xb1 + ro) : "memory");
            }
            ++p_tile;
            if (16 * p_tile >= p_len) {                          // unit done: open the next one
                ++pk;
                p_tile = 0;
                p_rslot = p_rslot + 1 == US_R ? 0 : p_rslot + 1;
                rec_request(pk + US_D, p_rslot >= US_D ? p_rslot - US_D : p_rslot + US_D);   // arrives US_D steps before it is opened
                if (pk < nmine) {
                    __syncwarp();                                // records copied by other lanes (complete: see header)
                    const UnitRec* r = &recs[p_rslot];
                    p_len = (int)r->len;
                    p_w0 = r->ids0[lane]; p_w1 = nx_w1; p_w2 = nx_w2; p_w3 = nx_w3;
                    if (pk + 1 < nmine) {
                        const UnitRec* rn = &recs[p_rslot + 1 == US_R ? 0 : p_rslot + 1];
                        nx_w1 = win_load(rn, 1); nx_w2 = win_load(rn, 2); nx_w3 = win_load(rn, 3);
                    }
                }
            } else if (!(p_tile & 1)) {                          // next id window of this unit
                p_w0 = p_w1; p_w1 = p_w2; p_w2 = p_w3;
                p_w3 = win_load(&recs[p_rslot], (p_tile >> 1) + 3);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int s = 0; s < US_D - 1; ++s) issue(s);

    // ---- consumer side ----------------------------------------------------------------------------------------------
    // IMMA.16x8x32 with the QUERIES as the 16-row A operand (queries g and 8 + g of the unit, kept in registers) and 8
    // bucket rows as the B operand: thread (g, t) supplies row g of the group, and the two B registers of a k-step are
    // two consecutive words of one of its 16-byte chunks — an LDS.128 feeds two IMMAs with no register shuffling.
    // Results: c0, c1 = (query g, rows 2t, 2t + 1), c2, c3 = (query 8 + g, same rows).
    uint4 bq_nx[2][2];                 // [query g / 8 + g][chunk] of the unit after the current one
    double tau_nx[2], qn_nx[2];
    int qj_nx[2], qq_nx[2];
    auto stage2 = [&](int rslot) {
        const UnitRec* r = &recs[rslot];
        const int m = (int)r->m;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            const int j = 8 * nb + g;
            qj_nx[nb] = r->q[j];                                 // slots >= m repeat the last query
            const unsigned char* qp = Q8 + (size_t)qj_nx[nb] * U8_QPITCH + 16 * t;
            bq_nx[nb][0] = ldg_u4(qp);
            bq_nx[nb][1] = ldg_u4(qp + 64);
            tau_nx[nb] = j < m ? __ldg(flt.tau + qj_nx[nb]) : __longlong_as_double(0x7ff0000000000000LL);   // +inf: masked
            if (ANGULAR) qn_nx[nb] = __ldg(qnorm + qj_nx[nb]);
            if (L2) qq_nx[nb] = __ldg(qsq + qj_nx[nb]);
        }
    };
    stage2(0);
    int c_slot = 0, c_rslot = 0;       // consumer's tile slot and record slot; the producer writes tile slot c_slot - 1
    for (int k = 0; k < nmine; ++k) {
        __syncwarp();                                            // records are copied by other lanes
        const UnitRec* r = &recs[c_rslot];
        const uint32_t bstart = r->bstart;
        const int len = (int)r->len;
        const bool two_blocks = r->m > 8;
        // A fragments of the four k-steps (chunk c, half h): {Qg.w[2h], Q8g.w[2h], Qg.w[2h+1], Q8g.w[2h+1]}
        unsigned A[2][2][4];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            A[c][0][0] = bq_nx[0][c].x; A[c][0][1] = bq_nx[1][c].x; A[c][0][2] = bq_nx[0][c].y; A[c][0][3] = bq_nx[1][c].y;
            A[c][1][0] = bq_nx[0][c].z; A[c][1][1] = bq_nx[1][c].z; A[c][1][2] = bq_nx[0][c].w; A[c][1][3] = bq_nx[1][c].w;
        }
        int c_q[2], c_taui[2], c_qq[2];
        double c_tau[2], c_qn[2];
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            c_q[nb] = qj_nx[nb];
            const double tv = tau_nx[nb];
            c_tau[nb] = tv;
            c_qn[nb] = ANGULAR ? qn_nx[nb] : 1.0;
            c_qq[nb] = L2 ? qq_nx[nb] : 0;
            // a dot product of bytes is below 2^31 - 1: INT_MAX masks the slot
            c_taui[nb] = tv >= 2147483647.0 ? 0x7fffffff : (tv <= -2147483648.0 ? (int)0x80000000 : (int)ceil(tv));
        }
        rows_staged += (unsigned)len;
        c_rslot = c_rslot + 1 == US_R ? 0 : c_rslot + 1;
        if (k + 1 < nmine) stage2(c_rslot);                      // its record is in the ring since before unit k was produced
        for (int tile = 0; 16 * tile < len; ++tile) {
            issue(c_slot == 0 ? US_D - 1 : c_slot - 1);
            asm volatile("cp.async.wait_group %0;" ::"n"(US_D - 1) : "memory");
            const uint4* src = ring + c_slot * (4 * 32);
            c_slot = c_slot + 1 == US_D ? 0 : c_slot + 1;
            int acc[2][4];                                       // [row half][c0..c3]
            int xx[2][2];                                        // |row|^2 of rows 8rr + 2t + e
            double xnr[2][2];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const uint4 b0 = src[(2 * rr) * 32], b1 = src[(2 * rr + 1) * 32];     // row 8rr + g: chunks t and 4 + t
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[rr][i] = 0;
                imma_u8(acc[rr], A[0][0][0], A[0][0][1], A[0][0][2], A[0][0][3], b0.x, b0.y);
                imma_u8(acc[rr], A[0][1][0], A[0][1][1], A[0][1][2], A[0][1][3], b0.z, b0.w);
                if (two_chunks) {
                    imma_u8(acc[rr], A[1][0][0], A[1][0][1], A[1][0][2], A[1][0][3], b1.x, b1.y);
                    imma_u8(acc[rr], A[1][1][0], A[1][1][1], A[1][1][2], A[1][1][3], b1.z, b1.w);
                }
                xx[rr][0] = xx[rr][1] = 0;
                xnr[rr][0] = xnr[rr][1] = 1.0;
                if (ANGULAR || L2) {                             // this thread loaded row 8rr + g; its results are rows 2t, 2t + 1
                    unsigned sq = 0, sc = 0;
                    sc = __dp4a(b0.x, b0.x, sc); sc = __dp4a(b0.y, b0.y, sc); sc = __dp4a(b0.z, b0.z, sc); sc = __dp4a(b0.w, b0.w, sc);
                    if (has0) sq += sc;
                    sc = 0;
                    sc = __dp4a(b1.x, b1.x, sc); sc = __dp4a(b1.y, b1.y, sc); sc = __dp4a(b1.z, b1.z, sc); sc = __dp4a(b1.w, b1.w, sc);
                    if (has1) sq += sc;
                    sq += __shfl_xor_sync(0xffffffffu, sq, 1);
                    sq += __shfl_xor_sync(0xffffffffu, sq, 2);   // lanes with the same g hold |row 8rr + g|^2
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        xx[rr][e] = (int)__shfl_sync(0xffffffffu, sq, (2 * t + e) * 4);
                        if (ANGULAR) xnr[rr][e] = sqrt((double)xx[rr][e]);
                    }
                }
            }
            // Nearly every tile has no score above its thresholds: one vote over all eight scores of the thread decides
            // whether to look closer.  Dot product: the largest of a query's four scores against its bar is all the vote
            // needs (4 instructions per query instead of a compare + row / slot masks per score); rows past the bucket's
            // end repeat its last row and unused query slots carry INT_MAX as their bar, so the only false alarms are
            // tiles the exact test below would also open.
            if (!ANGULAR && !L2) {
                const int m0 = max(max(acc[0][0], acc[0][1]), max(acc[1][0], acc[1][1]));
                const int m1 = max(max(acc[0][2], acc[0][3]), max(acc[1][2], acc[1][3]));
                if (!__any_sync(0xffffffffu, m0 >= c_taui[0] || m1 >= c_taui[1])) continue;
            }
            int key[2][2][2];                                    // [query g / 8 + g][row half][row 2t + e]
            bool pass[2][2][2];
            bool any = false;
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int dot = acc[rr][2 * nb + e];
                        key[nb][rr][e] = L2 ? 2 * dot - xx[rr][e] - c_qq[nb] : dot;
                        bool p;
                        if (ANGULAR) {
                            // dot / den >= tau  <=  dot >= tau * den up to rounding: a slightly lower bar here, the exact
                            // quotient decides below
                            const double bar = c_tau[nb] * (c_qn[nb] * xnr[rr][e]);
                            p = (double)dot >= bar - 1e-12 * fabs(bar) || !(bar == bar);
                        } else {
                            p = key[nb][rr][e] >= c_taui[nb];
                        }
                        p = p && 16 * tile + 8 * rr + 2 * t + e < len && (nb == 0 || two_blocks);
                        pass[nb][rr][e] = p;
                        any |= p;
                    }
            if (__any_sync(0xffffffffu, any)) {
#pragma unroll
                for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            double v = (double)key[nb][rr][e];
                            bool keep = pass[nb][rr][e];
                            if (ANGULAR) { v = v / (c_qn[nb] * xnr[rr][e]); keep = keep && v >= c_tau[nb]; }
                            sink.push(flt, keep, c_q[nb], bstart + (uint32_t)(16 * tile + 8 * rr + 2 * t + e), v, lane);
                        }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    sink.flush(flt, lane);
    if (lane == 0) { atomicAdd(&stat[0], (unsigned long long)nmine); atomicAdd(&stat[1], (unsigned long long)rows_staged); }
}

// ---------------------------------------------------------------------------------------------------------
// k_threshold_u8i — the threshold samples (see k_threshold, rerank_bm.cu) for byte rows and byte queries on the integer
// tensor pipe: one warp per (query, sampled table), the query in column 0 of the B operand (the other 7 columns are 0),
// 16 rows per IMMA tile exactly as in k_score_u8i, so a sample row costs 1/20 of the instructions of the DP4A /
// shuffle-reduction form and the kernel is bound by the row gather alone.  The dot products are the exact integers the
// scoring kernel will produce; cosines are computed with the same operations.
// ---------------------------------------------------------------------------------------------------------
template <int METRIC, bool REG /* K <= 32: the sample list lives in registers */>
__global__ void __launch_bounds__(RR_THREADS, 4)   // 64 registers: measured 2.15 ms per configs[1] step against 2.22 (3 CTAs, 80 registers) and 2.24 (5 CTAs, spills)
k_threshold_u8i(const unsigned char* __restrict__ X8, unsigned pitch, ChunkView cv, int L, int NT,
                const uint32_t* __restrict__ leaf_pos, const int32_t* __restrict__ leaf_len,
                const int32_t* __restrict__ ids_sorted, int self_exclude, int K,
                double* __restrict__ tl_keys, int* __restrict__ tl_ids, int* __restrict__ tl_cnt) {
    constexpr bool ANGULAR = METRIC == DPF_METRIC_ANGULAR, L2 = METRIC == DPF_METRIC_L2;
    if (*cv.q8_bad != 0) return;                 // some query is not a byte vector: k_threshold<.., false> samples the batch
    const unsigned char* __restrict__ Q8 = cv.Q8;
    const double* __restrict__ qnorm = cv.qnorm8;
    const int32_t* __restrict__ qsq = cv.qsq8;
    const int32_t* __restrict__ qids = cv.qids;
    const int64_t nqc = cv.nqc;
    extern __shared__ double rsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    double* mykeys = rsm + (size_t)warp * K;
    int* myids = reinterpret_cast<int*>(rsm + (size_t)RR_WARPS * K) + (size_t)warp * K;
    const int64_t wid = (int64_t)blockIdx.x * RR_WARPS + warp;
    if (wid >= nqc * NT) return;
    const int64_t ql = wid / NT;
    const int sample = (int)(wid % NT);
    const int64_t q = ql;
    const int qid = qids ? qids[q] : INT32_MIN;
    const bool excl = self_exclude && qids && qid >= -128 && qid <= 127;
    const bool has0 = 16u * t < pitch, has1 = 64u + 16u * t < pitch;
    const bool two_chunks = pitch > 64;
    const unsigned off0 = has0 ? 16u * t : 0u, off1 = has1 ? 64u + 16u * t : 0u;   // clamped, see k_score_u8d
    // B operand: column g = 0 is the query, the rest 0
    uint4 bq0 = make_uint4(0, 0, 0, 0), bq1 = make_uint4(0, 0, 0, 0);
    if (g == 0) {
        const unsigned char* qp = Q8 + (size_t)q * U8_QPITCH + 16 * t;            // zero beyond column d
        bq0 = ldg_u4(qp);
        bq1 = ldg_u4(qp + 64);
    }
    const double qn = ANGULAR ? __ldg(qnorm + q) : 1.0;
    const int qq = L2 ? __ldg(qsq + q) : 0;
    // the sample-th table (among the first 32) in which the query probes something
    uint32_t nonempty = __ballot_sync(0xffffffffu, lane < L && cv.pair_cnt[ql * L + min(lane, L - 1)] > 0u);
    for (int i = 0; i < sample && nonempty; ++i) nonempty &= nonempty - 1;
    int count = 0;
    double kth = 0.0;
    RegList rl;
    // the first bucket the query probes in that table (in practice its own bucket); should it hold fewer than K rows, the
    // next ones too, until K rows have been seen (a sample that cannot reach K rows would leave the query without a threshold)
    const int st_ = nonempty ? __ffs(nonempty) - 1 : 0;
    const uint32_t nb_ = nonempty ? cv.pair_cnt[ql * L + st_] : 0u;
    int seen = 0;
    for (uint32_t e_ = 0; e_ < nb_ && (e_ == 0 || seen < K); ++e_) {
        const uint32_t leaf = cv.cache[(ql * L + st_) * cv.cap + e_];
        const uint32_t bstart = leaf_pos[leaf];
        const int len = leaf_len[leaf];
        seen += len;
        const int32_t* bids = ids_sorted + bstart;
        // (Tried: the rows of window w + 1 gathered while window w is multiplied — 148 registers, or 128 with spills: one or
        // two CTAs per SM instead of three, and the kernel went from 0.57 to 0.97 ms inside a configs[1] step.)
        int idA = __ldg(bids + min(lane, len - 1));
        for (int row0 = 0; row0 < len; row0 += 32) {                 // one id window = 2 tiles of 16 rows
            int id_next = 0;
            if (row0 + 32 < len) id_next = __ldg(bids + min(row0 + 32 + lane, len - 1));
            uint4 a[2][2][2];                                        // [tile][row g / 8 + g][chunk]
#pragma unroll
            for (int tl = 0; tl < 2; ++tl)
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int row = min(row0 + 16 * tl + 8 * rr + g, len - 1);
                    const unsigned char* xp = X8 + (size_t)__shfl_sync(0xffffffffu, idA, row - row0) * pitch;
                    a[tl][rr][0] = ldg_u4(xp + off0);
                    a[tl][rr][1] = ldg_u4(xp + off1);
                }
#pragma unroll
            for (int tl = 0; tl < 2; ++tl) {
                if (row0 + 16 * tl >= len) break;                    // warp-uniform
                int acc[4] = {0, 0, 0, 0};
                imma_u8(acc, a[tl][0][0].x, a[tl][1][0].x, a[tl][0][0].y, a[tl][1][0].y, bq0.x, bq0.y);
                imma_u8(acc, a[tl][0][0].z, a[tl][1][0].z, a[tl][0][0].w, a[tl][1][0].w, bq0.z, bq0.w);
                if (two_chunks) {
                    imma_u8(acc, a[tl][0][1].x, a[tl][1][1].x, a[tl][0][1].y, a[tl][1][1].y, bq1.x, bq1.y);
                    imma_u8(acc, a[tl][0][1].z, a[tl][1][1].z, a[tl][0][1].w, a[tl][1][1].w, bq1.z, bq1.w);
                }
                double xnr[2] = {1.0, 1.0};
                int xx[2] = {0, 0};
                if (ANGULAR || L2) {
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        unsigned sq = 0;
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            unsigned sc = 0;
                            sc = __dp4a(a[tl][rr][c].x, a[tl][rr][c].x, sc); sc = __dp4a(a[tl][rr][c].y, a[tl][rr][c].y, sc);
                            sc = __dp4a(a[tl][rr][c].z, a[tl][rr][c].z, sc); sc = __dp4a(a[tl][rr][c].w, a[tl][rr][c].w, sc);
                            if (c == 0 ? has0 : has1) sq += sc;
                        }
                        sq += __shfl_xor_sync(0xffffffffu, sq, 1);
                        sq += __shfl_xor_sync(0xffffffffu, sq, 2);
                        if (ANGULAR) xnr[rr] = sqrt((double)sq);
                        xx[rr] = (int)sq;
                    }
                }
                // lanes t = 0 hold column 0: acc[0] = row g, acc[2] = row 8 + g
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int row = row0 + 16 * tl + 8 * rr + g;
                    double lb = (double)(L2 ? 2 * acc[2 * rr] - xx[rr] - qq : acc[2 * rr]);
                    if (ANGULAR) { lb = lb / (qn * xnr[rr]); lb -= 8.0 * 1.1102230246251565e-16 * fabs(lb); }
                    const int id = __shfl_sync(0xffffffffu, idA, min(row, len - 1) - row0);
                    bool ok = t == 0 && row < len && lb == lb && !(excl && id == qid);
                    if (ok && count == K) ok = lb >= kth;
                    uint32_t todo = __ballot_sync(0xffffffffu, ok);
                    while (todo) {                                   // rare once the list is full
                        const int src = __ffs(todo) - 1;
                        todo &= todo - 1;
                        const double lb_s = __shfl_sync(0xffffffffu, lb, src);
                        const int id_s = __shfl_sync(0xffffffffu, id, src);
                        if (REG) {
                            rl.insert(K, lb_s, id_s, lane);          // rows of one table are distinct; a row that is not among
                            count = rl.count;                        // the K best leaves the list unchanged
                            if (count == K) kth = __shfl_sync(0xffffffffu, rl.key, K - 1);
                        } else {
                            if (count == K && !better(lb_s, id_s, mykeys[K - 1], myids[K - 1])) continue;
                            warp_insert(mykeys, myids, count, K, lb_s, id_s, lane);
                            if (count == K) kth = mykeys[K - 1];
                        }
                    }
                }
            }
            idA = id_next;
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

void launch_threshold_u8i(dpf_index* h, cudaStream_t st, int metric, const ChunkView& cv, int NT, int topk, size_t list_smem) {
    const unsigned grid = (unsigned)((cv.nqc * NT + RR_WARPS - 1) / RR_WARPS);
    auto go = [&](auto kern) {
        kern<<<grid, RR_THREADS, list_smem, st>>>(h->Xc.p, (unsigned)h->Xc_row_bytes, cv, h->cfg.L, NT, h->leaf_pos.p, h->leaf_len.p,
                                                  h->ids_sorted.p, h->cfg.self_exclude_small_ids, topk, h->bm_tl_keys.p,
                                                  h->bm_tl_ids.p, h->bm_tl_cnt.p);
        DPF_LAUNCHED();
    };
    if (topk <= 32) {
        if (metric == DPF_METRIC_ANGULAR) go(k_threshold_u8i<DPF_METRIC_ANGULAR, true>);
        else if (metric == DPF_METRIC_L2) go(k_threshold_u8i<DPF_METRIC_L2, true>);
        else go(k_threshold_u8i<DPF_METRIC_DOT, true>);
    } else {
        if (metric == DPF_METRIC_ANGULAR) go(k_threshold_u8i<DPF_METRIC_ANGULAR, false>);
        else if (metric == DPF_METRIC_L2) go(k_threshold_u8i<DPF_METRIC_L2, false>);
        else go(k_threshold_u8i<DPF_METRIC_DOT, false>);
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
int u8_query_pitch() { return U8_QPITCH; }

bool score_u8_usable(const dpf_index* h) {
    if (h->dbg[DPF_DBG_BM_KERNEL] == 1) return false;         // test hook: the TMA ring kernel on the byte rows
    return h->Xc_kind == DPF_STORE_KIND_U8 && h->Xc_row_bytes <= 128;
}

// byte copy of the batch's queries + the device flag "some value is not a byte" (CTR_Q8_BAD).  The flag stays on the
// device: the kernels of both forms are launched and read it.  Only squared L2 needs it on the host (`need_host_flag`),
// because there the alternative to the integer pipeline is a different path altogether (row-major).
void prepare_queries_u8(dpf_index* h, const double* Qd, int64_t nq, bool need_host_flag) {
    h->Q8_valid = false;
    if (h->dbg[DPF_DBG_U8_IMMA] == 0) return;                 // test hook: always multiply on the FP64 tensor pipe
    if ((int64_t)h->cfg.d * 255 * 255 >= (1LL << 31)) return;  // integer dot products must stay below 2^31
    cudaStream_t st = h->stream;
    const int pitch = U8_QPITCH;
    h->Q8.reserve((size_t)nq * pitch);
    h->qnorm8.reserve((size_t)nq);
    h->qsq8.reserve((size_t)nq);
    int* flag = h->counters.p + CTR_Q8_BAD;
    k_quantise_queries<<<(unsigned)((nq + 7) / 8), 256, 0, st>>>(Qd, nq, h->cfg.d, pitch, h->Q8.p, h->qnorm8.p, h->qsq8.p, flag); DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
    if (!need_host_flag) return;
    int f = 1;
    DPF_CUDA(cudaMemcpyAsync(&f, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    DPF_CUDA(cudaStreamSynchronize(st));
    h->Q8_valid = f == 0;
}

// int_kernel: also launch the mma.sync integer kernel (it runs when the batch is byte vectors; false when the tcgen05
// kernel has that case); the FP64-query kernel is launched for dot / angular and runs when the batch is not bytes
void launch_score_u8(dpf_index* h, const ChunkView& cv, const void* units_v, const uint32_t* nunits_p, int metric,
                     const Filter& flt, unsigned long long* bm_stat, bool int_kernel) {
    const bool angular = metric == DPF_METRIC_ANGULAR;
    const UnitRec* units = reinterpret_cast<const UnitRec*>(units_v);
    const unsigned pitch = (unsigned)h->Xc_row_bytes;
    cudaStream_t st = h->stream;
    const bool try_int = h->dbg[DPF_DBG_U8_IMMA] != 0 && (int64_t)h->cfg.d * 255 * 255 < (1LL << 31);
    if (int_kernel) {                                             // runs when the batch is byte vectors (device flag)
        if (h->dbg[DPF_DBG_U8I_KERNEL] == 1 && metric != DPF_METRIC_L2) {       // test hook: the occupancy-based variant
            auto launch = [&](auto kern) {
                kern<<<h->num_sms * (angular ? U8_INT_CTAS - 1 : U8_INT_CTAS), U8_WARPS * 32, 0, st>>>(
                    h->Xc.p, pitch, cv.Q8, cv.qnorm8, cv.q8_bad, units, nunits_p, h->ids_sorted.p, flt, bm_stat);
                DPF_LAUNCHED();
            };
            if (angular) launch(k_score_u8i<true>); else launch(k_score_u8i<false>);
        } else {
            auto launch_s = [&](auto kern) {
                DPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)US_SMEM));
                kern<<<h->num_sms * US_CTAS, U8_WARPS * 32, US_SMEM, st>>>(h->Xc.p, pitch, cv.Q8, cv.qnorm8, cv.qsq8, cv.q8_bad, units,
                                                                         nunits_p, h->ids_sorted.p, flt, bm_stat);
                DPF_LAUNCHED();
            };
            if (angular) launch_s(k_score_u8s<DPF_METRIC_ANGULAR>);
            else if (metric == DPF_METRIC_L2) launch_s(k_score_u8s<DPF_METRIC_L2>);
            else launch_s(k_score_u8s<DPF_METRIC_DOT>);
        }
    }
    if (metric != DPF_METRIC_L2) {                                // runs when some query is not a byte vector
        auto launch = [&](auto kern) {
            kern<<<h->num_sms, U8_WARPS * 32, 0, st>>>(h->Xc.p, pitch, h->cfg.d, cv.Q, cv.q8_bad, try_int ? 1 : 0, units, nunits_p,
                                                       h->ids_sorted.p, flt, bm_stat);
            DPF_LAUNCHED();
        };
        if (angular) launch(k_score_u8d<true>); else launch(k_score_u8d<false>);
    }
}

}  // namespace dpf
