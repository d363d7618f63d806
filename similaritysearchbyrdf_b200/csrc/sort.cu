// sort.cu — stable LSD radix sort and prefix scan used by the forest build (K3) and by the sorted-candidate
// output.  Hand-written for this path: keys are short bit-fields (table | root | slot), so only the bits that
// matter are sorted, with the digit width chosen per call.
//
// One pass = k_radix_hist (per-tile digit histogram) -> exclusive scan over (digit, tile) -> k_radix_scatter
// (re-read the tile, stable rank, scatter).  Stability is what lets the build keep ids in ascending order inside
// every tree node, which the reference's insertion-order-dependent split rule needs.
#include "common.cuh"

namespace dpf {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_STEPS = 16;                       // 32-item steps per warp
constexpr int RS_TILE = RS_THREADS * RS_STEPS;     // 4096 items per CTA
constexpr int RS_MAXBITS = 9;                      // digit width: 17 key bits of the build sort = 2 passes, 25 of the pair sort = 3
constexpr int RS_MAXBINS = 1 << RS_MAXBITS;

// ---- exclusive scan (uint32, in place) ---------------------------------------------------------------------
constexpr int SC_THREADS = 256, SC_ITEMS = 8, SC_TILE = SC_THREADS * SC_ITEMS;

__global__ void __launch_bounds__(SC_THREADS) k_scan_tile(uint32_t* __restrict__ data, int64_t n,
                                                          uint32_t* __restrict__ partials) {
    __shared__ uint32_t wsum[SC_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    uint32_t v[SC_ITEMS], s = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        v[i] = (base + i < n) ? data[base + i] : 0u;
        s += v[i];
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t ws = lane < SC_THREADS / 32 ? wsum[lane] : 0u, wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += y;
        }
        if (lane < SC_THREADS / 32) wsum[lane] = wi - ws;
        if (lane == SC_THREADS / 32 - 1 && partials) partials[blockIdx.x] = wi;
    }
    __syncthreads();
    uint32_t run = wsum[w] + inc - s;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        if (base + i < n) data[base + i] = run;
        run += v[i];
    }
}

// Single-pass scan for arrays of more than one tile: tiles take their index from a counter (so a tile only ever waits
// for tiles that are already running), publish their aggregate, and warp 0 looks back over the predecessors' status
// words 32 at a time — (flag << 32 | value), flag 1 = aggregate of that tile, 2 = inclusive prefix up to and including
// it — until it meets a prefix.  One read and one write of the data instead of the three kernels of the tile / add form.
constexpr int SL_ITEMS = 16, SL_TILE = SC_THREADS * SL_ITEMS;

__global__ void __launch_bounds__(SC_THREADS) k_scan_lookback(uint32_t* __restrict__ data, int64_t n, unsigned int* __restrict__ counter,
                                                              volatile unsigned long long* __restrict__ status) {
    __shared__ uint32_t wsum[SC_THREADS / 32];
    __shared__ uint32_t s_tile, s_total, s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t base = (int64_t)tile * SL_TILE + (int64_t)threadIdx.x * SL_ITEMS;
    uint32_t v[SL_ITEMS], s = 0;
#pragma unroll
    for (int i = 0; i < SL_ITEMS; ++i) {
        v[i] = (base + i < n) ? data[base + i] : 0u;
        s += v[i];
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t ws = lane < SC_THREADS / 32 ? wsum[lane] : 0u, wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += y;
        }
        if (lane < SC_THREADS / 32) wsum[lane] = wi - ws;
        const uint32_t total = __shfl_sync(0xffffffffu, wi, SC_THREADS / 32 - 1);
        uint32_t prefix = 0;
        if (tile == 0) {
            if (lane == 0) status[0] = (2ULL << 32) | total;
        } else {
            if (lane == 0) status[tile] = (1ULL << 32) | total;
            long long p = (long long)tile - 1;                        // look back: lane l reads tile p - l
            for (;;) {
                const long long mine = p - lane;
                unsigned long long st = mine >= 0 ? 0ULL : (2ULL << 32);   // before tile 0: an empty prefix
                if (mine >= 0)
                    do { st = status[mine]; } while ((st >> 32) == 0ULL);
                const uint32_t has_prefix = __ballot_sync(0xffffffffu, (st >> 32) == 2ULL);
                const int first = has_prefix ? __ffs(has_prefix) - 1 : 32;   // nearest predecessor with a full prefix
                uint32_t part = lane <= first ? (uint32_t)st : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                prefix += part;
                if (has_prefix) break;
                p -= 32;
            }
            if (lane == 0) status[tile] = (2ULL << 32) | (unsigned long long)(prefix + total);
        }
        if (lane == 0) { s_prefix = prefix; s_total = total; }
    }
    __syncthreads();
    uint32_t run = s_prefix + wsum[w] + inc - s;
#pragma unroll
    for (int i = 0; i < SL_ITEMS; ++i) {
        if (base + i < n) data[base + i] = run;
        run += v[i];
    }
}

// scratch: see scan_scratch_elems
static void scan_u32_inplace(uint32_t* data, int64_t n, uint32_t* scratch, cudaStream_t st) {
    if (n <= SC_TILE) {
        k_scan_tile<<<1, SC_THREADS, 0, st>>>(data, n, nullptr); DPF_LAUNCHED();
        return;
    }
    const int64_t nb = (n + SL_TILE - 1) / SL_TILE;
    // [0..1] tile counter (+ pad to 8 bytes), then one 64-bit status word per tile
    cudaMemsetAsync(scratch, 0, (size_t)(2 + 2 * nb) * sizeof(uint32_t), st);
    k_scan_lookback<<<(unsigned)nb, SC_THREADS, 0, st>>>(data, n, scratch, reinterpret_cast<unsigned long long*>(scratch + 2)); DPF_LAUNCHED();
}

static size_t scan_scratch_elems(int64_t n) {
    return (size_t)(2 + 2 * ((n + SL_TILE - 1) / SL_TILE)) + 8;
}

// int32 counts -> int64 exclusive offsets (n+1 entries), single CTA chained over tiles of 8192 (8 consecutive entries per
// thread: 15 rounds of barriers for the 122 880 depth-1 codes of a build instead of 120; per-query offsets where n is small)
constexpr int SC64_ITEMS = 8;
__global__ void __launch_bounds__(1024) k_scan_i32_to_i64(const int32_t* __restrict__ in, int64_t* __restrict__ out,
                                                          int64_t n) {
    __shared__ long long wsum[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += 1024 * SC64_ITEMS) {
        const int64_t i0 = base + (int64_t)threadIdx.x * SC64_ITEMS;
        long long v[SC64_ITEMS];
        long long tot = 0;
#pragma unroll
        for (int e = 0; e < SC64_ITEMS; ++e) {
            v[e] = i0 + e < n ? (long long)in[i0 + e] : 0;
            tot += v[e];
        }
        long long inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        if (w == 0) {
            long long ws = wsum[lane], wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long y = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += y;
            }
            wsum[lane] = wi - ws;
        }
        __syncthreads();
        const long long c = carry;
        long long run = c + wsum[w] + inc - tot;         // exclusive prefix of this thread's first entry
#pragma unroll
        for (int e = 0; e < SC64_ITEMS; ++e) {
            if (i0 + e < n) out[i0 + e] = run;
            run += v[e];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = run;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

// in-place exclusive scan of n uint32 values (sum must fit 32 bits); multi-CTA, for arrays of millions of entries
void exclusive_scan_u32(dpf_index* h, uint32_t* data, int64_t n) {
    if (n <= 0) return;
    h->scan_scratch.reserve(scan_scratch_elems(n));
    scan_u32_inplace(data, n, h->scan_scratch.p, h->stream);
    DPF_CUDA(cudaGetLastError());
}

void exclusive_scan_i64(dpf_index* h, const int32_t* in, int64_t* out, int64_t n) {
    k_scan_i32_to_i64<<<1, 1024, 0, h->stream>>>(in, out, n); DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
}

// ---- radix sort ---------------------------------------------------------------------------------------------
template <class K>
__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const K* __restrict__ keys, int64_t n, int shift,
                                                           int bits, uint32_t* __restrict__ hist, int64_t ntiles) {
    __shared__ uint32_t cnt[RS_MAXBINS];
    const int nbins = 1 << bits;
    for (int i = threadIdx.x; i < nbins; i += RS_THREADS) cnt[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
    const uint32_t mask = (uint32_t)nbins - 1u;
#pragma unroll 4
    for (int s = 0; s < RS_STEPS; ++s) {
        const int64_t i = base + (int64_t)s * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&cnt[(uint32_t)(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += RS_THREADS) hist[(int64_t)i * ntiles + blockIdx.x] = cnt[i];
}

template <class K, bool HAS_VAL>
__global__ void __launch_bounds__(RS_THREADS)
k_radix_scatter(const K* __restrict__ keys, const uint32_t* __restrict__ vals, K* __restrict__ keys_out,
                uint32_t* __restrict__ vals_out, int64_t n, int shift, int bits, const uint32_t* __restrict__ hist,
                int64_t ntiles) {
    __shared__ uint32_t wcnt[RS_WARPS][RS_MAXBINS];
    const int nbins = 1 << bits;
    const uint32_t mask = (uint32_t)nbins - 1u;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * RS_MAXBINS; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    // warp w owns items [base + w*512, +512), visited in ascending index order => stable
    const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)w * (32 * RS_STEPS);
    K kreg[RS_STEPS];
    uint16_t loc[RS_STEPS];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int s = 0; s < RS_STEPS; ++s) {
        const int64_t i = wbase + s * 32 + lane;
        const bool ok = i < n;
        kreg[s] = ok ? keys[i] : (K)0;
        const uint32_t dgt = ok ? ((uint32_t)(kreg[s] >> shift) & mask) : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(0xffffffffu, dgt);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (ok && lane == leader) {
            old = wcnt[w][dgt];
            wcnt[w][dgt] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        loc[s] = (uint16_t)(old + __popc(peers & lt));
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over warps on top of the tile's global base
    for (int dgt = threadIdx.x; dgt < nbins; dgt += RS_THREADS) {
        uint32_t run = hist[(int64_t)dgt * ntiles + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) {
            const uint32_t c = wcnt[ww][dgt];
            wcnt[ww][dgt] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < RS_STEPS; ++s) {
        const int64_t i = wbase + s * 32 + lane;
        if (i < n) {
            const uint32_t dgt = (uint32_t)(kreg[s] >> shift) & mask;
            const uint32_t pos = wcnt[w][dgt] + loc[s];
            keys_out[pos] = kreg[s];
            if (HAS_VAL) vals_out[pos] = vals[i];
        }
    }
}

template <class K, bool HAS_VAL>
static void radix_sort_impl(dpf_index* h, K** keys, K** keys_alt, uint32_t** vals, uint32_t** vals_alt, int64_t n,
                            int lo_bit, int hi_bit) {
    if (n <= 1 || hi_bit <= lo_bit) return;
    DPF_REQUIRE(n < (1LL << 32), DPF_ERR_INVALID, "radix sort: more than 2^32 items in one call");
    const int total = hi_bit - lo_bit;
    const int passes = (total + RS_MAXBITS - 1) / RS_MAXBITS;
    const int64_t ntiles = (n + RS_TILE - 1) / RS_TILE;
    const size_t hist_elems = (size_t)RS_MAXBINS * ntiles;
    h->hist.reserve(hist_elems + scan_scratch_elems((int64_t)hist_elems));
    int bit = lo_bit;
    for (int p = 0; p < passes; ++p) {
        const int bits = (total - (bit - lo_bit) + (passes - p) - 1) / (passes - p);   // balanced digit widths
        const int64_t used = (int64_t)(1 << bits) * ntiles;
        k_radix_hist<K><<<(unsigned)ntiles, RS_THREADS, 0, h->stream>>>(*keys, n, bit, bits, h->hist.p, ntiles); DPF_LAUNCHED();
        scan_u32_inplace(h->hist.p, used, h->hist.p + hist_elems, h->stream);
        k_radix_scatter<K, HAS_VAL><<<(unsigned)ntiles, RS_THREADS, 0, h->stream>>>(
            *keys, HAS_VAL ? *vals : nullptr, *keys_alt, HAS_VAL ? *vals_alt : nullptr, n, bit, bits, h->hist.p, ntiles); DPF_LAUNCHED();
        DPF_CUDA(cudaGetLastError());
        std::swap(*keys, *keys_alt);
        if (HAS_VAL) std::swap(*vals, *vals_alt);
        bit += bits;
    }
}

void radix_sort_pairs_u32(dpf_index* h, uint32_t** keys, uint32_t** keys_alt, uint32_t** vals, uint32_t** vals_alt,
                          int64_t n, int lo_bit, int hi_bit) {
    radix_sort_impl<uint32_t, true>(h, keys, keys_alt, vals, vals_alt, n, lo_bit, hi_bit);
}

void radix_sort_keys_u64(dpf_index* h, unsigned long long** keys, unsigned long long** keys_alt, int64_t n, int lo_bit,
                         int hi_bit) {
    radix_sort_impl<unsigned long long, false>(h, keys, keys_alt, nullptr, nullptr, n, lo_bit, hi_bit);
}

}  // namespace dpf
