// query_common.cuh — device helpers shared by query.cu (K4 probe / expand, K5 row-major re-rank) and rerank_bm.cu
// (K5b bucket-major re-rank): the flat-forest descent with the reference's probe list, and the warp-cooperative
// top-k list.
#pragma once
#include "common.cuh"

namespace dpf {

// ---------------------------------------------------------------------------------------------------------
// descent
// ---------------------------------------------------------------------------------------------------------
struct ProbeCtx {
    ForestView f;
    TreeParams tp;
    int L, steps, probe_mode, world, self_exclude;
    OwnMask own;       // this GPU's sub-forest
};

// bucket lookup for one probe key (RandomDrawTreeMap.java:940-994): empty slot -> nothing; leaf -> (ptr,cnt);
// directory -> descend; falling off level 0 -> nothing
__device__ __forceinline__ bool descend(const ForestView& f, const TreeParams& tp, int root_node, uint32_t probe,
                                        int& ptr, int& cnt, int64_t* slot_index = nullptr) {
    int node = root_node;
    for (int level = tp.MAXL; level >= 0; --level) {
        const int slot = (int)((probe >> (tp.nb * level)) & (uint32_t)(tp.W - 1));
        const int64_t idx = (int64_t)node * tp.W + slot;
        const int c = __ldg(f.child_cnt + idx);
        const int p = __ldg(f.child_ptr + idx);
        if (c == 0) return false;
        if (c > 0) { ptr = p; cnt = c; if (slot_index) *slot_index = idx; return true; }
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

// The same, returning the dense leaf number (ForestView::child_leaf) of every distinct bucket instead of its range.
__device__ __forceinline__ void warp_lookup_leaf(const ProbeCtx& c, int t, int sub, int seg, uint32_t h, int nprobes,
                                                 int lane, bool& leader, uint32_t& leaf, int& cnt) {
    int ptr = 0;
    cnt = 0;
    leaf = 0;
    bool ok = false;
    int64_t idx = 0;
    if (lane < nprobes) {
        const uint32_t probe = (c.probe_mode == DPF_PROBE_NONE) ? h : (h ^ (1u << lane));
        ok = descend(c.f, c.tp, t * c.tp.R + sub * c.tp.SEG + seg, probe, ptr, cnt, &idx);
    }
    const int key = ok ? ptr : (-1 - lane);
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    leader = ok && ((__ffs(peers) - 1) == lane);
    if (leader) leaf = __ldg(c.f.child_leaf + idx);
}

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

// The same list for K <= 32 kept in registers, entry i in lane i (valid for i < count, best first): an insertion is a
// ballot, one shuffle per field and two selects instead of the shared-memory shift loop (~12 instructions against ~80).
// The selection kernels are bound by exactly these insertions (ncu: k_threshold_u8i issues 25 instructions per sampled
// row, 15 of them here).
struct RegList {
    double key = 0.0;
    int id = 0x7fffffff;
    int count = 0;
    __device__ __forceinline__ bool full(int K) const { return count == K; }
    __device__ __forceinline__ void kth(int K, double& k, int& i) const {          // entry K - 1 (when full)
        k = __shfl_sync(0xffffffffu, key, K - 1);
        i = __shfl_sync(0xffffffffu, id, K - 1);
    }
    __device__ __forceinline__ bool contains(int x, int lane) const { return __any_sync(0xffffffffu, lane < count && id == x); }
    __device__ __forceinline__ void insert(int K, double nk, int ni, int lane) {
        const int pos = __popc(__ballot_sync(0xffffffffu, lane < count && better(key, id, nk, ni)));
        if (pos >= K) return;
        const double uk = __shfl_up_sync(0xffffffffu, key, 1);
        const int ui = __shfl_up_sync(0xffffffffu, id, 1);
        if (lane > pos && lane <= count && lane < K) { key = uk; id = ui; }
        if (lane == pos) { key = nk; id = ni; }
        count = min(count + 1, K);
    }
};

inline ProbeCtx make_ctx(dpf_index* h, int steps, int probe_mode) {
    ProbeCtx c;
    c.f = forest_view(h);
    c.tp = h->tp;
    c.L = h->cfg.L;
    c.steps = steps;
    c.probe_mode = probe_mode;
    c.world = h->cfg.world > 1 ? h->cfg.world : 1;
    c.own = h->own;
    c.self_exclude = h->cfg.self_exclude_small_ids;
    return c;
}

}  // namespace dpf
