// rerank_tc.cu — K5t: bucket-major scoring of byte rows x byte queries on the 5th-generation tensor cores
// (tcgen05.mma.kind::i8, accumulators in tensor memory).
//
// Same job as k_score_u8s (rerank_u8.cu) — for every unit (one leaf bucket x the queries that probe it) the exact
// integer dot product of every (query, bucket row), replacing the gather + dgemv of topKAndPrecisionScore
// (src/main/scala/mclab/deploy/DensevectorRDFInit.scala:472-507); only scores that reach the query's threshold leave
// the kernel (rerank_units.cuh) — but the multiplication is one instruction per 128 rows x <= 32 queries x 32 bytes,
// issued by one thread: a unit is twice as wide as the mma.sync kernel's (a bucket is staged once per 32 queries instead
// of once per 16: 33.7M instead of 46.9M rows on configs[1]) and the per-score instruction count drops from ~23 to ~2.
//
// A unit here = <= 128 rows of one bucket x <= 32 of the queries that probe it, described by one self-contained record
// (k_emit_tc_recs: header, query list, row ids).  One persistent CTA per SM, four roles connected by mbarrier rings:
//   record loader (1 warp) streams the CTA's unit records into a ring of TC_R slots with cp.async, as far ahead as the
//                         ring allows — every index the other roles need is in shared memory
//                         long before they need it (the first version fetched descriptors, query lists and row ids with
//                         ordinary loads one unit ahead and spent 1.5 us per unit waiting for them)
//   row producers (4 warps) gather the unit's rows by id from the byte store into shared memory in the UMMA K-major
//                         128-byte-swizzle layout: a row is 128 bytes = one swizzle row, thread (row group, chunk c)
//                         copies the 16-byte chunks c of 8 rows with cp.async to chunk position c ^ (row & 7); the unit's
//                         queries (rows of the byte copy of the batch) go behind them the same way; TC_S stages, all of
//                         them in flight; the stage's mbarrier gets a thread's arrival when its copies have landed
//                         (cp.async.mbarrier.arrive), so nobody waits for data it does not need yet.
//   MMA issuer (1 thread) waits for a stage, issues <= 4 tcgen05.mma (K = 32 bytes each) into one of TC_NACC accumulator
//                         stages of tensor memory (128 lanes = rows, <= 32 columns = queries), and commits the stage's
//                         shared memory back to the producers and the accumulator to the epilogue (tcgen05.commit).
//   epilogue (16 warps)   four sets of four warps take the units round-robin: tcgen05.ld the accumulator (warp w reads
//                         lanes 32 (w % 4) .., all 32 columns), compare with the integer thresholds of the unit's queries
//                         (gathered next to the record by the loader), push the few survivors, free the unit's record.
#include "rerank_units.cuh"

namespace dpf {

constexpr int TC_ROWS = 128;                      // rows per unit = UMMA M
constexpr int TC_S = 8;                           // stage ring: a unit's rows (16 KB) + its queries (4 KB), all of them in flight
constexpr int TC_R = 32;                          // unit-record ring (loader -> producers -> epilogue)
constexpr int TC_TL = 6;                          // units the threshold gather trails the record copy by
constexpr int TC_NACC = 8;                        // accumulator stages
constexpr int TC_TMEM_COLS = TC_NACC * TC_TQ;     // 256 of the 512 columns
constexpr int TC_EPI_SETS = 4;                    // epilogue warp sets; set j takes the units k = j (mod 4) of the CTA
constexpr int TC_EPI_WARPS = 4 * TC_EPI_SETS, TC_PROD_WARPS = 4;
constexpr int TC_THREADS = (TC_EPI_WARPS + 2 + TC_PROD_WARPS) * 32;
constexpr int TC_A_BYTES = TC_ROWS * 128, TC_B_BYTES = TC_TQ * 128, TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
struct __align__(16) TcSlot {                     // a unit's record in shared memory + the thresholds of its queries
    TcRec rec;
    int32_t tau[TC_TQ];
};
constexpr size_t TC_SMEM = 1024 /* alignment slack */ + (size_t)TC_S * TC_STAGE_BYTES + (size_t)TC_R * sizeof(TcSlot) + 256;
static_assert(TC_S + TC_NACC + TC_TL + TC_EPI_SETS + 2 <= TC_R, "ring depths");
static_assert(TC_TQ == 32, "one threshold per loader lane, one tcgen05.ld.x32 per epilogue warp, two query rows per producer thread");
static_assert(TC_STAGE_BYTES % 1024 == 0 && sizeof(TcRec) % 16 == 0, "swizzle atoms / 16-byte record chunks");
static_assert((TC_TMEM_COLS & (TC_TMEM_COLS - 1)) == 0 && TC_TMEM_COLS >= 32 && TC_TMEM_COLS <= 512, "tensor memory is allocated in powers of two");

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
// cycles spent waiting, per barrier tag (lane 0 of every warp), and [15] = cycles of the kernel summed over CTAs:
// which role is the bottleneck is the one that does not wait (dpf_debug_tc_diag returns these after the watchdog words)
__device__ unsigned long long g_tc_prof[16];
struct WaitClock {                       // per-thread wait cycles by barrier tag; lane 0 of a warp reports them once, at the end
    long long c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    __device__ void report() {
        if ((threadIdx.x & 31) == 0)
            for (int i = 0; i < 8; ++i)
                if (c[i]) atomicAdd(&g_tc_prof[i], (unsigned long long)c[i]);
    }
};
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, unsigned parity, int tag, volatile int* abort_flag, WaitClock& wc) {
    const long long t0 = clock64();
    struct Acc {
        long long t0; long long& dst;
        __device__ ~Acc() { dst += clock64() - t0; }
    } acc{t0, wc.c[tag]};
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
// the mbarrier gets one arrival from this thread when all its cp.async issued so far have landed (.noinc: the arrival
// is one of the barrier's expected count, not an extra one)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void cp4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
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

__global__ void __launch_bounds__(TC_THREADS, 1)
k_score_u8t(const unsigned char* __restrict__ X8, unsigned pitch /* row bytes: multiple of 16, <= 128 */,
            const unsigned char* __restrict__ Q8 /* pitch 128, zero padded */, const int* __restrict__ q8_bad,
            const TcRec* __restrict__ recs, const uint32_t* __restrict__ nunits_p, uint32_t cap,
            const int32_t* __restrict__ taui /* per query, [tau_sentinel] = INT_MAX */, int tau_sentinel, Filter flt,
            unsigned long long* __restrict__ stat) {
    using namespace tc;
    if (*q8_bad != 0) return;                    // some query is not a byte vector: k_score_u8d scores the batch
    extern __shared__ unsigned char tc_smem_raw[];
    __shared__ uint64_t rec_full[TC_R], tau_full[TC_R], rec_empty[TC_R], a_full[TC_S], a_empty[TC_S], acc_full[TC_NACC], acc_empty[TC_NACC];
    __shared__ uint32_t s_tmem;
    __shared__ int s_abort;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* stages = base;                                          // TC_S x (16 KB rows + 4 KB queries), 1024-byte aligned
    TcSlot* ring = reinterpret_cast<TcSlot*>(base + (size_t)TC_S * TC_STAGE_BYTES);  // TC_R unit records + thresholds

    const int64_t nunits = min(*nunits_p, cap);
    const int64_t G = gridDim.x;
    const int64_t nmine = nunits > blockIdx.x ? (nunits - blockIdx.x + G - 1) / G : 0;   // units blockIdx.x + k G
    // rows narrower than 128 bytes: the chunks beyond the row are never copied and must read as zero
    for (int i = tid; i < TC_S * TC_STAGE_BYTES / 16; i += TC_THREADS) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        s_abort = 0;
        for (int i = 0; i < TC_R; ++i) { mbar_init(&rec_full[i], 32); mbar_init(&tau_full[i], 32); mbar_init(&rec_empty[i], 4); }
        for (int i = 0; i < TC_S; ++i) { mbar_init(&a_full[i], TC_PROD_WARPS * 32); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < TC_NACC; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
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
    const long long t_start = clock64();
    WaitClock wc;
    const int nk = (int)((pitch + 31) / 32);                               // K steps of 32 bytes that hold data

    if (warp < TC_EPI_WARPS) {
        // ================================ epilogue ========================================================================
        // A unit's epilogue is a chain of dependent long-latency steps (barrier check, TMEM load, threshold compare) that one
        // warp cannot overlap with itself — measured ~2400 cycles per unit — so TC_EPI_SETS sets of 4 warps take the CTA's
        // units round-robin; warp (set, quarter) reads TMEM lanes 32 quarter .. of the units k = set (mod TC_EPI_SETS).
        const int quarter = warp & 3, set = warp >> 2;
        SurvivorSink sink;
        unsigned long long rows_scored = 0;
        for (int64_t k = set; k < nmine && !s_abort; k += TC_EPI_SETS) {
            const int slot = (int)(k % TC_R), a = (int)(k % TC_NACC);
            if (!mbar_wait(&tau_full[slot], (unsigned)((k / TC_R) & 1), 6, &s_abort, wc)) break;   // record and thresholds are in
            const TcSlot* r = &ring[slot];
            const uint32_t bstart = r->rec.bstart, nrows = r->rec.nrows, row0 = r->rec.row0;
            if (!mbar_wait(&acc_full[a], (unsigned)((k / TC_NACC) & 1), 3, &s_abort, wc)) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = 32 * quarter + lane;
            const bool valid = (uint32_t)row < nrows;
            const uint32_t pos = bstart + row0 + (uint32_t)row;
            int v[32];
            tmem_ld32(tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(a * TC_TQ), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);                   // the accumulator stage is free again
            bool any = false;                                            // columns past the unit's queries carry INT_MAX
#pragma unroll
            for (int jj = 0; jj < 32; jj += 4) {
                const int4 t4 = *reinterpret_cast<const int4*>(r->tau + jj);
                any |= v[jj] >= t4.x || v[jj + 1] >= t4.y || v[jj + 2] >= t4.z || v[jj + 3] >= t4.w;
            }
            // survivors cluster in the few buckets near a query: most units have none
            if (__any_sync(0xffffffffu, any && valid)) {
#pragma unroll
                for (int jj = 0; jj < 32; ++jj)
                    sink.push(flt, valid && v[jj] >= r->tau[jj], r->rec.q[jj], pos, (double)v[jj], lane);
            }
            if (quarter == 0) rows_scored += nrows;
            __syncwarp();
            if (lane == 0) mbar_arrive(&rec_empty[slot]);                // the last readers of the unit's record
        }
        sink.flush(flt, lane);
        if (lane == 0 && quarter == 0) { atomicAdd(&stat[0], (unsigned long long)((nmine + TC_EPI_SETS - 1 - set) / TC_EPI_SETS)); atomicAdd(&stat[1], rows_scored); }
    } else if (warp == TC_EPI_WARPS) {
        // ================================ MMA issuer ======================================================================
        if (lane == 0) {
            for (int64_t k = 0; k < nmine && !s_abort; ++k) {
                const int s = (int)(k % TC_S), a = (int)(k % TC_NACC);
                if (!mbar_wait(&a_full[s], (unsigned)((k / TC_S) & 1), 1, &s_abort, wc)) break;
                if (k >= TC_NACC && !mbar_wait(&acc_empty[a], (unsigned)(((k / TC_NACC) - 1) & 1), 2, &s_abort, wc)) break;
                // (no proxy fence between the producers' cp.async writes and the tensor core's reads: the stage's mbarrier
                // completes when the copies have landed — the same protocol as CUTLASS's sm100 cp.async mainloop; the
                // full-batch bit-exactness tests against the mma.sync kernels guard it)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sbase = smem_u32(stages + (size_t)s * TC_STAGE_BYTES);
                const uint64_t adesc = smem_desc_sw128(sbase), bdesc = smem_desc_sw128(sbase + TC_A_BYTES);
                // N = TC_TQ always: the columns past the unit's queries multiply whatever the stage held before and are
                // never read (the issuer then needs nothing from the unit's record)
                const uint32_t idesc = idesc_u8(TC_TQ);
                for (int kk = 0; kk < nk; ++kk)                           // + 32 bytes along K = + 2 in the 16-byte address field
                    umma_i8(tmem_base + (uint32_t)(a * TC_TQ), adesc + 2 * kk, bdesc + 2 * kk, idesc, kk > 0 ? 1u : 0u);
                umma_commit(&a_empty[s]);                                 // the stage's shared memory, back to the producers
                umma_commit(&acc_full[a]);                                // the accumulator, to the epilogue
            }
        }
    } else if (warp == TC_EPI_WARPS + 1) {
        // ================================ record loader ===================================================================
        // unit records stream into the ring with cp.async as far ahead as the ring allows (its slots are freed by the
        // epilogue); TC_TL units behind, lane j gathers the integer threshold of the unit's query j next to the record
        // (4-byte cp.async; slots past the unit's queries take the sentinel INT_MAX at taui[nq]).  Nothing downstream ever
        // waits for a global load of an index or a threshold.
        constexpr int CH = (int)(sizeof(TcRec) / 16);
        static_assert(CH > 32 && CH <= 64, "two chunks per lane at most");
        auto gather_tau = [&](int64_t k) {
            const int slot = (int)(k % TC_R);
            if (!mbar_wait(&rec_full[slot], (unsigned)((k / TC_R) & 1), 6, &s_abort, wc)) return false;
            const TcSlot* r = &ring[slot];
            const int qj = (uint32_t)lane < r->rec.m ? r->rec.q[lane] : tau_sentinel;
            cp4(smem_u32(&ring[slot].tau[lane]), taui + qj);
            cp_async_arrive(&tau_full[slot]);
            return true;
        };
        int64_t k = 0;
        for (; k < nmine && !s_abort; ++k) {
            const int slot = (int)(k % TC_R);
            if (k >= TC_R && !mbar_wait(&rec_empty[slot], (unsigned)(((k / TC_R) - 1) & 1), 7, &s_abort, wc)) break;
            const unsigned char* src = reinterpret_cast<const unsigned char*>(recs + (blockIdx.x + k * G));
            const uint32_t dst = smem_u32(&ring[slot].rec);
            cp16(dst + 16 * lane, src + 16 * lane);
            if (lane + 32 < CH) cp16(dst + 16 * (lane + 32), src + 16 * (lane + 32));
            cp_async_arrive(&rec_full[slot]);                             // published the moment the copies land
            if (k >= TC_TL && !gather_tau(k - TC_TL)) break;
        }
        for (int64_t kk = k > TC_TL ? k - TC_TL : 0; kk < k && !s_abort; ++kk)
            if (!gather_tau(kk)) break;
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else {
        // ================================ row producers ===================================================================
        const int p = tid - (TC_EPI_WARPS + 2) * 32;                      // 0 .. 127
        const int c = p & 7, rg = p >> 3;                                 // 16-byte chunk, row group: rows rg + 16 i
        const bool has_chunk = 16u * c < pitch;
        const uint32_t swz = (uint32_t)((c ^ (rg & 7)) << 4);             // (rg + 16 i) & 7 == rg & 7
        const uint32_t a0 = smem_u32(stages) + (uint32_t)rg * 128u + swz;
        const unsigned char* xsrc = X8 + 16 * c;
        const unsigned char* qsrc = Q8 + 16 * c;
        int64_t k = 0;
        for (; k < nmine && !s_abort; ++k) {
            if (!mbar_wait(&rec_full[k % TC_R], (unsigned)((k / TC_R) & 1), 6, &s_abort, wc)) break;
            const TcRec* r = &ring[k % TC_R].rec;
            int id[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) id[i] = r->ids[rg + 16 * i];
            const uint32_t m = r->m;
            const int q0 = r->q[rg], q1 = r->q[rg + 16];
            const int s = (int)(k % TC_S);
            if (k >= TC_S && !mbar_wait(&a_empty[s], (unsigned)(((k / TC_S) - 1) & 1), 4, &s_abort, wc)) break;
            const uint32_t sa = a0 + (uint32_t)s * TC_STAGE_BYTES;
            if (has_chunk) {
#pragma unroll
                for (int i = 0; i < 8; ++i) cp16(sa + (uint32_t)i * 2048u, xsrc + (size_t)(unsigned)id[i] * pitch);
            }
            // the unit's queries: rows rg and rg + 16 of the B operand (N = m rounded up to 16)
            if ((uint32_t)rg < ((m + 15) & ~15u)) cp16(sa + TC_A_BYTES, qsrc + (size_t)q0 * 128);
            if ((uint32_t)(rg + 16) < ((m + 15) & ~15u)) cp16(sa + TC_A_BYTES + 2048u, qsrc + (size_t)q1 * 128);
            // the stage is published when this thread's copies have landed; all TC_S stages can be in flight, and the
            // producer never waits for its own data (an arrival that lags the copies would leave the MMA -> a_empty loop
            // only TC_S - lag stages of slack: the first version ran at one memory latency per unit because of that)
            cp_async_arrive(&a_full[s]);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    wc.report();
    if (tid == 0) atomicAdd(&g_tc_prof[15], (unsigned long long)(clock64() - t_start));
    // ---- teardown: everything issued has been consumed (the epilogue waited for every accumulator) ------------------------
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS) : "memory");
    }
}

void tc_diag_read(unsigned long long* out24) {
    DPF_CUDA(cudaMemcpyFromSymbol(out24, tc::g_tc_diag, 8 * sizeof(unsigned long long)));
    DPF_CUDA(cudaMemcpyFromSymbol(out24 + 8, tc::g_tc_prof, 16 * sizeof(unsigned long long)));
}

bool score_u8t_usable(const dpf_index* h, int metric) {
    // dot product only (the reference's re-rank metric); angular and squared L2 need the row norms next to the
    // accumulator and stay on the mma.sync kernel.  Opt-in (DPF_DBG_U8I_KERNEL = 3): on configs[1] this kernel is
    // correct but slower than k_score_u8s — the units of this workload are small (58 rows x 6 queries on average), and
    // every role of the asynchronous pipeline pays ~1000 cycles of barrier hand-offs per unit (profiles/r2_tcgen05_*).
    return metric == DPF_METRIC_DOT && h->dbg[DPF_DBG_U8I_KERNEL] == 3;
}

void launch_score_u8t(dpf_index* h, const ChunkView& cv, const TcRec* recs, const uint32_t* nunits_p, int64_t cap, const int32_t* taui,
                      const Filter& flt, unsigned long long* bm_stat) {
    DPF_CUDA(cudaFuncSetAttribute(k_score_u8t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    k_score_u8t<<<h->num_sms, TC_THREADS, TC_SMEM, h->stream>>>(h->Xc.p, (unsigned)h->Xc_row_bytes, cv.Q8, cv.q8_bad, recs, nunits_p, (uint32_t)cap,
                                                               taui, (int)cv.nqc, flt, bm_stat);
    DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
}

}  // namespace dpf
