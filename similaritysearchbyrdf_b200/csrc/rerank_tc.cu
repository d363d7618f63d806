// rerank_tc.cu — K5t: bucket-major scoring of byte rows x byte queries on the 5th-generation tensor cores
// (tcgen05.mma.kind::i8, accumulators in tensor memory).
//
// Same job as k_score_u8s (rerank_u8.cu) — for every unit (one leaf bucket x the queries that probe it) the exact
// integer dot product of every (query, bucket row), replacing the gather + dgemv of topKAndPrecisionScore
// (src/main/scala/mclab/deploy/DensevectorRDFInit.scala:472-507); only scores that reach the query's threshold leave
// the kernel (rerank_units.cuh) — but the multiplication is one instruction per 128 rows x <= 64 queries x 32 bytes,
// issued by one thread, so a unit is four times wider than the mma.sync kernel's (a bucket is staged once per 64
// queries instead of once per 16) and the per-score instruction count drops from ~23 to ~2.
//
// One persistent CTA per SM, three roles:
//   producers (4 warps)   gather the bucket's rows by id from the byte store into shared memory in the UMMA K-major
//                         128-byte-swizzle layout: a row is 128 bytes = one swizzle row, thread (row group, chunk c)
//                         copies the 16-byte chunks c of 8 rows with cp.async to chunk position c ^ (row & 7); a ring of
//                         TC_S tiles of 128 rows, TC_D - 1 of them in flight per thread; the unit's queries (rows of the
//                         byte copy of the batch) go to one of TC_NB operand buffers the same way.  A thread waits for
//                         its own copies of a tile (cp.async.wait_group), fences them towards the async proxy and
//                         arrives on the tile's mbarrier.
//   MMA issuer (1 thread) waits for a tile, issues <= 4 tcgen05.mma (K = 32 bytes each) into one of TC_NACC accumulator
//                         stages of tensor memory (128 lanes = rows, <= 64 columns = queries), and commits the tile's
//                         shared-memory slot back to the producers and the accumulator to the epilogue (tcgen05.commit
//                         -> mbarrier).
//   epilogue (8 warps)    tcgen05.ld the accumulator (warp w reads lanes 32 (w % 4) .., 16 columns at a time; the two
//                         warps of a lane quarter split the column groups), compare with the integer thresholds of the
//                         unit's queries (a private copy per warp in shared memory), and push the few survivors.
// No demand load sits between a wait and the work it guards: unit descriptors, query lists, thresholds and row ids are
// all requested one unit / one tile ahead.
#include "rerank_units.cuh"

namespace dpf {

constexpr int TC_ROWS = 128;                      // rows per tile = UMMA M
constexpr int TC_S = 8;                           // row-tile ring
constexpr int TC_D = 7;                           // cp.async groups a producer thread keeps in flight (< TC_S)
constexpr int TC_NB = 8;                          // query-operand buffers: >= TC_D, because a thread's arrival for a tile lags its
                                                  // copies by TC_D - 1 tiles and buffer k % TC_NB is only free once the tiles of
                                                  // unit k - TC_NB have been multiplied (units can be one tile long)
constexpr int TC_NACC = 4;                        // accumulator stages
constexpr int TC_TMEM_COLS = TC_NACC * TC_TQ;     // 256 of the 512 columns
constexpr int TC_EPI_WARPS = 8, TC_PROD_WARPS = 4;
constexpr int TC_THREADS = (TC_EPI_WARPS + 1 + TC_PROD_WARPS) * 32;
constexpr int TC_A_BYTES = TC_ROWS * 128, TC_B_BYTES = TC_TQ * 128;
constexpr size_t TC_SMEM = 1024 /* alignment slack */ + (size_t)TC_S * TC_A_BYTES + (size_t)TC_NB * TC_B_BYTES +
                           (size_t)TC_EPI_WARPS * 2 * TC_TQ * 8 + 256;
static_assert(TC_NB >= TC_D && TC_D < TC_S, "see TC_NB");
static_assert(TC_TQ % 16 == 0 && TC_TQ <= 256 && TC_TMEM_COLS <= 512, "UMMA N / tensor-memory budget");
static_assert((TC_TMEM_COLS & (TC_TMEM_COLS - 1)) == 0 && TC_TMEM_COLS >= 32, "tensor memory is allocated in powers of two");

namespace tc {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Every wait in this kernel is bounded: a role that waits longer than ~1 s for a barrier (a protocol error — it never
// happens in a correct run) records which barrier it was in g_tc_diag, raises the CTA's abort flag and every role of
// the CTA leaves its loops, so a bug shows up as a wrong result and a non-zero dpf_debug_tc_diag, not as a hung GPU.
__device__ unsigned long long g_tc_diag[8];
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, unsigned parity, int tag, volatile int* abort_flag) {
    const long long t0 = clock64();
    for (;;) {
        unsigned ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return true;
        if (*abort_flag) return false;
        if (clock64() - t0 > 2000000000LL) {
            *abort_flag = 1;
            atomicAdd(&g_tc_diag[tag], 1ULL);
            atomicCAS(&g_tc_diag[0], 0ULL, ((unsigned long long)tag << 32) | (unsigned long long)blockIdx.x);
            return false;
        }
    }
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// shared-memory matrix descriptor, K-major, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ULL << 16) /* LBO (unused for swizzled K-major) */ |
           ((uint64_t)(1024 >> 4) << 32) /* SBO */ | (1ULL << 46) /* descriptor version */ | (2ULL << 61) /* SWIZZLE_128B */;
}
// instruction descriptor, kind::i8: D = s32, A = B = unsigned 8 bit, both K-major, M = 128
__device__ __forceinline__ uint32_t idesc_u8(int n) {
    return (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
}  // namespace tc

struct TcUnit { uint32_t bstart, len, pair0, m; };

__global__ void __launch_bounds__(TC_THREADS, 1)
k_score_u8t(const unsigned char* __restrict__ X8, unsigned pitch /* row bytes: multiple of 16, <= 128 */,
            const unsigned char* __restrict__ Q8 /* pitch 128, zero padded */, const int* __restrict__ q8_bad,
            const UnitDesc* __restrict__ units, const uint32_t* __restrict__ nunits_p, const int32_t* __restrict__ pair_q,
            const int32_t* __restrict__ ids_sorted, Filter flt, unsigned long long* __restrict__ stat) {
    using namespace tc;
    if (*q8_bad != 0) return;                    // some query is not a byte vector: k_score_u8d scores the batch
    extern __shared__ unsigned char tc_smem_raw[];
    __shared__ uint64_t a_full[TC_S], a_empty[TC_S], b_empty[TC_NB], acc_full[TC_NACC], acc_empty[TC_NACC];
    __shared__ uint32_t s_tmem;
    __shared__ int s_abort;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* a_tiles = base;                                         // TC_S x 16 KB, 1024-byte aligned
    unsigned char* b_tiles = base + (size_t)TC_S * TC_A_BYTES;            // TC_NB x 8 KB
    int* epi_meta = reinterpret_cast<int*>(b_tiles + (size_t)TC_NB * TC_B_BYTES);   // per epilogue warp: 2 x (taui[TQ], q[TQ])

    const int64_t nunits = *nunits_p;
    const int64_t G = gridDim.x;
    const int64_t nmine = nunits > blockIdx.x ? (nunits - blockIdx.x + G - 1) / G : 0;   // units blockIdx.x + k G
    // rows narrower than 128 bytes: the chunks beyond the row are never copied and must read as zero
    for (int i = tid; i < (TC_S * TC_A_BYTES + TC_NB * TC_B_BYTES) / 16; i += TC_THREADS)
        reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        s_abort = 0;
        for (int i = 0; i < TC_S; ++i) { mbar_init(&a_full[i], TC_PROD_WARPS * 32); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < TC_NB; ++i) mbar_init(&b_empty[i], 1);
        for (int i = 0; i < TC_NACC; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {                             // tensor memory: allocated and released by warp 0
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // the zero fill, towards the tensor core's reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;
    const int nk = (int)((pitch + 31) / 32);                               // K steps of 32 bytes that hold data

    auto unit_at = [&](int64_t k) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(units + (blockIdx.x + k * G)));
        return TcUnit{v.x, v.y, v.z, v.w};
    };

    if (warp < TC_EPI_WARPS) {
        // ================================ epilogue ========================================================================
        const int quarter = warp & 3, half = warp >> 2;
        int* meta = epi_meta + (size_t)warp * 2 * (2 * TC_TQ);
        SurvivorSink sink;
        unsigned long long rows_scored = 0;
        // unit k + 1 is requested while unit k is processed: descriptor, this lane's two queries, their thresholds
        TcUnit nx = {0, 0, 0, 0};
        int nx_q[2] = {0, 0};
        double nx_tau[2] = {0.0, 0.0};
        auto request = [&](int64_t k) {
            if (k >= nmine) return;
            nx = unit_at(k);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                nx_q[e] = __ldg(pair_q + nx.pair0 + min((uint32_t)(lane + 32 * e), nx.m - 1u));
                nx_tau[e] = __ldg(flt.tau + nx_q[e]);
            }
        };
        request(0);
        int64_t g = 0;
        for (int64_t k = 0; k < nmine && !s_abort; ++k) {
            const TcUnit u = nx;
            int* taui = meta + (k & 1) * (2 * TC_TQ);
            int* qv = taui + TC_TQ;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = lane + 32 * e;
                const double tv = (uint32_t)j < u.m ? nx_tau[e] : __longlong_as_double(0x7ff0000000000000LL);   // +inf: masked
                // a dot product of bytes is below 2^31 - 1: INT_MAX masks the slot
                taui[j] = tv >= 2147483647.0 ? 0x7fffffff : (tv <= -2147483648.0 ? (int)0x80000000 : (int)ceil(tv));
                qv[j] = nx_q[e];
            }
            __syncwarp();
            request(k + 1);
            const int ntiles = (int)((u.len + TC_ROWS - 1) / TC_ROWS);
            const int ngroups = (int)((u.m + 15) >> 4);
            if (half == 0) rows_scored += u.len;
            for (int t = 0; t < ntiles; ++t, ++g) {
                const int a = (int)(g % TC_NACC);
                if (!mbar_wait(&acc_full[a], (unsigned)((g / TC_NACC) & 1), 3, &s_abort)) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int row = t * TC_ROWS + 32 * quarter + lane;
                const bool valid = row < (int)u.len;
                const uint32_t pos = u.bstart + (uint32_t)row;
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(a * TC_TQ);
                int v[2][16];                                            // <= 2 column groups per warp (TC_TQ = 64, two halves)
                static_assert(TC_TQ / 16 <= 4, "column groups per warp pair");
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int cg = 2 * i + half;
                    if (cg < ngroups) tmem_ld16(taddr + 16 * cg, v[i]);   // warp-uniform
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[a]);               // the accumulator stage is free again
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int cg = 2 * i + half;
                    if (cg >= ngroups) continue;
                    bool any = false;
#pragma unroll
                    for (int jj = 0; jj < 16; jj += 4) {
                        const int4 tq = *reinterpret_cast<const int4*>(taui + 16 * cg + jj);
                        any |= v[i][jj] >= tq.x || v[i][jj + 1] >= tq.y || v[i][jj + 2] >= tq.z || v[i][jj + 3] >= tq.w;
                    }
                    // survivors cluster in the few buckets near a query: most groups have none
                    if (__any_sync(0xffffffffu, any && valid)) {
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj)
                            sink.push(flt, valid && v[i][jj] >= taui[16 * cg + jj], qv[16 * cg + jj], pos, (double)v[i][jj], lane);
                    }
                }
            }
        }
        sink.flush(flt, lane);
        if (lane == 0 && quarter == 0 && half == 0) { atomicAdd(&stat[0], (unsigned long long)nmine); atomicAdd(&stat[1], rows_scored); }
    } else if (warp == TC_EPI_WARPS) {
        // ================================ MMA issuer ======================================================================
        if (lane == 0 && nmine > 0) {
            TcUnit nx = unit_at(0);
            int64_t g = 0;
            for (int64_t k = 0; k < nmine && !s_abort; ++k) {
                const TcUnit u = nx;
                if (k + 1 < nmine) nx = unit_at(k + 1);
                const int ntiles = (int)((u.len + TC_ROWS - 1) / TC_ROWS);
                const uint32_t idesc = idesc_u8((int)((u.m + 15) & ~15u));
                const int bslot = (int)(k % TC_NB);
                const uint64_t bdesc = smem_desc_sw128(smem_u32(b_tiles + (size_t)bslot * TC_B_BYTES));
                for (int t = 0; t < ntiles; ++t, ++g) {
                    const int s = (int)(g % TC_S), a = (int)(g % TC_NACC);
                    if (!mbar_wait(&a_full[s], (unsigned)((g / TC_S) & 1), 1, &s_abort)) break;
                    if (g >= TC_NACC && !mbar_wait(&acc_empty[a], (unsigned)(((g / TC_NACC) - 1) & 1), 2, &s_abort)) break;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t adesc = smem_desc_sw128(smem_u32(a_tiles + (size_t)s * TC_A_BYTES));
                    for (int kk = 0; kk < nk; ++kk)                       // + 32 bytes along K = + 2 in the 16-byte address field
                        umma_i8(tmem_base + (uint32_t)(a * TC_TQ), adesc + 2 * kk, bdesc + 2 * kk, idesc, kk > 0 ? 1u : 0u);
                    umma_commit(&a_empty[s]);                             // the tile's shared memory, back to the producers
                    umma_commit(&acc_full[a]);                            // the accumulator, to the epilogue
                }
                umma_commit(&b_empty[bslot]);                             // the unit's query operand
            }
        }
    } else {
        // ================================ producers =======================================================================
        const int p = tid - (TC_EPI_WARPS + 1) * 32;                      // 0 .. 127
        const int c = p & 7, rg = p >> 3;                                 // 16-byte chunk, row group: rows rg + 16 i
        const bool has_chunk = 16u * c < pitch;
        const uint32_t swz = (uint32_t)((c ^ (rg & 7)) << 4);             // (rg + 16 i) & 7 == rg & 7
        const uint32_t a0 = smem_u32(a_tiles) + (uint32_t)rg * 128u + swz;
        const uint32_t b0 = smem_u32(b_tiles) + (uint32_t)rg * 128u + swz;
        const unsigned char* xsrc = X8 + 16 * c;
        const unsigned char* qsrc = Q8 + 16 * c;
        if (nmine > 0) {
            TcUnit u0 = unit_at(0), u1 = nmine > 1 ? unit_at(1) : TcUnit{0, 1, 0, 1}, u2 = TcUnit{0, 1, 0, 1};
            int qi[4], qn[4], id[8], idn[8];
            auto load_q = [&](const TcUnit& u, int (&dst)[4]) {
#pragma unroll
                for (int i = 0; i < 4; ++i) dst[i] = __ldg(pair_q + u.pair0 + min((uint32_t)(rg + 16 * i), u.m - 1u));
            };
            auto load_ids = [&](const TcUnit& u, int t, int (&dst)[8]) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    dst[i] = __ldg(ids_sorted + u.bstart + min((uint32_t)(t * TC_ROWS + rg + 16 * i), u.len - 1u));
            };
            load_q(u0, qi);
            load_ids(u0, 0, id);
            int64_t g = 0;
            for (int64_t k = 0; k < nmine && !s_abort; ++k) {
                if (k + 2 < nmine) u2 = unit_at(k + 2);
                if (k + 1 < nmine) load_q(u1, qn);
                const int ntiles = (int)((u0.len + TC_ROWS - 1) / TC_ROWS);
                // the unit's queries -> operand buffer k % TC_NB (free once the MMAs of unit k - TC_NB are done)
                const int bslot = (int)(k % TC_NB);
                if (k >= TC_NB && !mbar_wait(&b_empty[bslot], (unsigned)(((k / TC_NB) - 1) & 1), 5, &s_abort)) break;
                const int npad = (int)((u0.m + 15) & ~15u);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (rg + 16 * i < npad) cp16(b0 + (uint32_t)bslot * TC_B_BYTES + (uint32_t)i * 2048u, qsrc + (size_t)qi[i] * 128);
                for (int t = 0; t < ntiles; ++t, ++g) {
                    const int s = (int)(g % TC_S);
                    if (g >= TC_S && !mbar_wait(&a_empty[s], (unsigned)(((g / TC_S) - 1) & 1), 4, &s_abort)) break;
                    if (has_chunk) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            cp16(a0 + (uint32_t)s * TC_A_BYTES + (uint32_t)i * 2048u, xsrc + (size_t)(unsigned)id[i] * pitch);
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    // ids of the next tile (of this unit, or the first of the next one)
                    if (t + 1 < ntiles) load_ids(u0, t + 1, idn);
                    else if (k + 1 < nmine) load_ids(u1, 0, idn);
                    if (g >= TC_D - 1) {                                  // tile g - (TC_D - 1): this thread's copies have landed
                        asm volatile("cp.async.wait_group %0;" ::"n"(TC_D - 1) : "memory");
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        mbar_arrive(&a_full[(g - (TC_D - 1)) % TC_S]);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) id[i] = idn[i];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) qi[i] = qn[i];
                u0 = u1;
                u1 = u2;
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (int64_t gg = g > TC_D - 1 ? g - (TC_D - 1) : 0; gg < g; ++gg) mbar_arrive(&a_full[gg % TC_S]);
        }
    }
    // ---- teardown: everything issued has been consumed (the epilogue waited for every accumulator) ------------------------
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS) : "memory");
    }
}

void tc_diag_read(unsigned long long* out8) {
    DPF_CUDA(cudaMemcpyFromSymbol(out8, tc::g_tc_diag, 8 * sizeof(unsigned long long)));
}

bool score_u8t_usable(const dpf_index* h, int metric) {
    // dot product only (the reference's re-rank metric); angular and squared L2 need the row norms next to the
    // accumulator and stay on the mma.sync kernel
    if (metric != DPF_METRIC_DOT) return false;
    const int64_t sel = h->dbg[DPF_DBG_U8I_KERNEL];
    return sel == 0 || sel == 3;
}

void launch_score_u8t(dpf_index* h, const ChunkView& cv, const UnitDesc* descs, const uint32_t* nunits_p, const Filter& flt,
                      unsigned long long* bm_stat) {
    DPF_CUDA(cudaFuncSetAttribute(k_score_u8t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    k_score_u8t<<<h->num_sms, TC_THREADS, TC_SMEM, h->stream>>>(h->Xc.p, (unsigned)h->Xc_row_bytes, cv.Q8, cv.q8_bad, descs, nunits_p,
                                                               h->pair_q.p, h->ids_sorted.p, flt, bm_stat);
    DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
}

}  // namespace dpf
