// incremental.cu — put and remove on the flat forest without rebuilding it.
//
// Replaces, for a batch of ids, RandomDrawTreeMap.put / putInner (src/main/java/mclab/mapdb/RandomDrawTreeMap.java:
// 1558-1584, 1662-1790) and remove / removeInternal / recursiveDirDelete (:1817-1932) — the latter with the tree geometry
// of the configuration instead of the reference's hard-coded 4 levels x 7 bits (quirk Q12).  Used by the facade overload
// that first inserts its query vectors (src/main/scala/mclab/deploy/DensevectorRDFInit.scala:215-251, 372-399): a
// two-vector insert costs O(batch), not a rebuild of 30M entries.
//
// Operations on different leaf slots commute (an insert or a remove only touches the slot it lands in, the subtree it may
// create below it, and — a remove that empties a bucket — the chain of directory nodes above it), so a batch is grouped by
// the leaf slot each (table, id) reaches in the tree as it is, and one warp applies a slot's operations one after the
// other in ascending id order: exactly the sequential semantics, in parallel over the slots.
//   k_locate_slots     thread per (table, id): descend with the id's stored key -> (slot index << 32 | id)
//   radix sort         by (slot, id)
//   k_insert_runs /    warp per run of equal slot.  A bucket that grows moves to the end of ids_sorted (the arena: buckets
//   k_remove_runs      stay contiguous, id-ascending ranges, so every query kernel reads them unchanged); a bucket at
//                      BUCKET_OVERFLOW splits into a new directory node one level down with a stable partition of its ids;
//                      a removed id is squeezed out of its bucket in place
//   k_collapse_dirs    thread per root: directory nodes left without children are deleted from their parents, bottom up
// The arena and the node array have head-room; when it runs out the caller falls back to the full rebuild.
#include "common.cuh"

namespace dpf {

__device__ __forceinline__ int inc_slot_at(int32_t h, int level, int nb, int mask) {
    return (int)(((uint32_t)h >> (nb * level)) & (uint32_t)mask);
}

constexpr unsigned long long kSkipKey = ~0ULL;

// leaf slot (or empty slot) the id's key leads to in the current tree
__global__ void __launch_bounds__(256)
k_locate_slots(const int32_t* __restrict__ ids, int64_t m, int L, const int32_t* __restrict__ keys, const uint8_t* __restrict__ pids,
               int64_t ld, int64_t n, TreeParams tp, OwnMask own, const int32_t* __restrict__ child_ptr,
               const int32_t* __restrict__ child_cnt, unsigned long long* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m * L) return;
    const int t = (int)(i / m);
    const int32_t id = ids[i % m];
    unsigned long long key = kSkipKey;
    if (id >= 0 && id < n) {
        const int pid = pids[(int64_t)t * ld + id];
        if (own.has(t, pid)) {
            const int32_t h = keys[(int64_t)t * ld + id];
            const int seg = tp.seg_bits ? (int)((uint32_t)h >> tp.bucket_bits) : 0;
            int node = t * tp.R + pid * tp.SEG + seg;
            for (int level = tp.MAXL; level >= 0; --level) {
                const int64_t idx = (int64_t)node * tp.W + inc_slot_at(h, level, tp.nb, tp.W - 1);
                const int c = child_cnt[idx];
                if (c < 0) { node = child_ptr[idx]; continue; }
                key = ((unsigned long long)idx << 32) | (uint32_t)id;
                break;
            }
        }
    }
    out[i] = key;
}

// positions where a run of equal slots starts -> list (order irrelevant)
__global__ void __launch_bounds__(256)
k_run_heads(const unsigned long long* __restrict__ sorted, int64_t total, uint32_t* __restrict__ heads, int* __restrict__ nheads) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const unsigned long long k = sorted[i];
    if (k == kSkipKey) return;
    if (i == 0 || (sorted[i - 1] >> 32) != (k >> 32)) heads[atomicAdd(nheads, 1)] = (uint32_t)i;
}

struct IncCtx {
    TreeParams tp;
    int L;
    const int32_t* keys;
    const uint8_t* pids;
    int64_t ld;
    const int64_t* table_base;
    int32_t* child_ptr;
    int32_t* child_cnt;
    int32_t* node_table;
    int32_t* ids_sorted;
    int32_t* counters;       // [0] node counter, [1] overflow flag, [4..5] arena cursor (u64), [8..9] singleton splits (u64), [10] splits
    int32_t node_cap;
    long long arena_cap;     // entries of ids_sorted
    int64_t roots;
};

constexpr int INC_WARPS = 4;
constexpr int INC_MAXW = 256;

// one warp per run: the ids of a run in ascending order, each inserted like putInner
__global__ void __launch_bounds__(INC_WARPS * 32)
k_insert_runs(IncCtx c, const unsigned long long* __restrict__ sorted, int64_t total, const uint32_t* __restrict__ heads,
              const int* __restrict__ nheads) {
    __shared__ int s_cnt[INC_WARPS][INC_MAXW], s_off[INC_WARPS][INC_MAXW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int run = blockIdx.x * INC_WARPS + warp;
    if (run >= *nheads) return;
    const TreeParams tp = c.tp;
    const int W = tp.W, mask = W - 1;
    int* cntW = s_cnt[warp];
    int* offW = s_off[warp];
    unsigned long long* arena = reinterpret_cast<unsigned long long*>(c.counters + 4);
    const int64_t head = heads[run];
    const unsigned long long slot_key = sorted[head] >> 32;
    for (int64_t e = head; e < total && (sorted[e] >> 32) == slot_key; ++e) {
        if (c.counters[1]) return;                                   // out of room somewhere: the caller rebuilds
        const int32_t id = (int32_t)(uint32_t)sorted[e];
        // table of the slot: roots know it by position, the other nodes carry it
        const int64_t node0 = (int64_t)slot_key / W;
        const int t = node0 < c.roots ? (int)(node0 / tp.R) : c.node_table[node0];
        const int32_t* kt = c.keys + (int64_t)t * c.ld;
        const int32_t h = kt[id];
        const int pid = c.pids[(int64_t)t * c.ld + id];
        const long long tb = c.table_base[t];
        const int seg = tp.seg_bits ? (int)((uint32_t)h >> tp.bucket_bits) : 0;
        int node = t * tp.R + pid * tp.SEG + seg;
        for (int level = tp.MAXL; level >= 0; --level) {
            const int64_t idx = (int64_t)node * W + inc_slot_at(h, level, tp.nb, mask);
            const int cn = c.child_cnt[idx];
            const int p = c.child_ptr[idx];
            if (cn < 0) { node = p; continue; }
            const int32_t* old = c.ids_sorted + tb + p;
            if (cn >= tp.T && level >= 1) {
                // ---- split: new directory one level down; the bucket's ids and the new id by their next slot ----------
                int nd = 0;
                unsigned long long ab = 0;
                if (lane == 0) { nd = atomicAdd(&c.counters[0], 1); ab = atomicAdd(arena, (unsigned long long)(cn + 1)); }
                nd = __shfl_sync(0xffffffffu, nd, 0);
                ab = __shfl_sync(0xffffffffu, ab, 0);
                if (nd >= c.node_cap || (long long)(ab + cn + 1) > c.arena_cap || (long long)(ab + cn + 1) - tb >= (1LL << 31)) {
                    if (lane == 0) c.counters[1] = 1;
                    return;
                }
                for (int i = lane; i < W; i += 32) cntW[i] = 0;
                __syncwarp();
                const int newpos = inc_slot_at(h, level - 1, tp.nb, mask);
                for (int i = lane; i <= cn; i += 32) {
                    const int sl = i < cn ? inc_slot_at(kt[old[i]], level - 1, tp.nb, mask) : newpos;
                    atomicAdd(&cntW[sl], 1);
                }
                __syncwarp();
                if (lane == 0) {
                    int run_ = 0;
                    for (int i = 0; i < W; ++i) { offW[i] = run_; run_ += cntW[i]; }
                    // quirk Q1 (RandomDrawTreeMap.java:1733-1734): counted, the intended bucket flag is implemented
                    if (cntW[newpos] == 1) atomicAdd(reinterpret_cast<unsigned long long*>(c.counters + 8), 1ULL);
                    atomicAdd(&c.counters[10], 1);
                    c.node_table[nd] = t;
                }
                __syncwarp();
                int32_t* dst = c.ids_sorted + ab;
                for (int sl = lane; sl < W; sl += 32) {                  // children of the new node
                    const int64_t ci = (int64_t)nd * W + sl;
                    c.child_cnt[ci] = cntW[sl];
                    c.child_ptr[ci] = cntW[sl] ? (int32_t)((long long)ab + offW[sl] - tb) : 0;
                }
                __syncwarp();
                // stable scatter in ascending position (= ascending id; the new id is the largest)
                for (int base = 0; base <= cn; base += 32) {
                    const int i = base + lane;
                    const bool ok = i <= cn;
                    int32_t y = 0;
                    uint32_t sl = 0xffffffffu;
                    if (ok) { y = i < cn ? old[i] : id; sl = (uint32_t)(i < cn ? inc_slot_at(kt[y], level - 1, tp.nb, mask) : newpos); }
                    const uint32_t peers = __match_any_sync(0xffffffffu, sl);
                    const int rank = __popc(peers & ((1u << lane) - 1u));
                    if (ok) dst[offW[sl] + rank] = y;
                    __syncwarp();
                    if (ok && rank == 0) offW[sl] += __popc(peers);
                    __syncwarp();
                }
                if (lane == 0) { c.child_ptr[idx] = nd; c.child_cnt[idx] = -1; }
            } else {
                // ---- the bucket grows (or is created) by one id: a fresh range at the end of the arena ----------------
                unsigned long long ab = 0;
                if (lane == 0) ab = atomicAdd(arena, (unsigned long long)(cn + 1));
                ab = __shfl_sync(0xffffffffu, ab, 0);
                if ((long long)(ab + cn + 1) > c.arena_cap || (long long)(ab + cn + 1) - tb >= (1LL << 31)) {
                    if (lane == 0) c.counters[1] = 1;
                    return;
                }
                int32_t* dst = c.ids_sorted + ab;
                for (int i = lane; i < cn; i += 32) dst[i] = old[i];
                if (lane == 0) dst[cn] = id;
                __syncwarp();
                if (lane == 0) { c.child_ptr[idx] = (int32_t)((long long)ab - tb); c.child_cnt[idx] = cn + 1; }
            }
            break;
        }
        __threadfence();
        __syncwarp();
    }
}

// one warp per run: the ids of a run leave their bucket (searchable again by the next id of the run)
__global__ void __launch_bounds__(INC_WARPS * 32)
k_remove_runs(IncCtx c, const unsigned long long* __restrict__ sorted, int64_t total, const uint32_t* __restrict__ heads,
              const int* __restrict__ nheads, int* __restrict__ removed_entries) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int run = blockIdx.x * INC_WARPS + warp;
    if (run >= *nheads) return;
    const TreeParams tp = c.tp;
    const int W = tp.W;
    const int64_t head = heads[run];
    const unsigned long long slot_key = sorted[head] >> 32;
    const int64_t idx = (int64_t)slot_key;
    const int64_t node0 = idx / W;
    const int t = node0 < c.roots ? (int)(node0 / tp.R) : c.node_table[node0];
    const long long tb = c.table_base[t];
    for (int64_t e = head; e < total && (sorted[e] >> 32) == slot_key; ++e) {
        const int32_t id = (int32_t)(uint32_t)sorted[e];
        const int cn = c.child_cnt[idx];
        if (cn <= 0) break;
        int32_t* b = c.ids_sorted + tb + c.child_ptr[idx];
        int pos = -1;
        for (int i0 = 0; i0 < cn && pos < 0; i0 += 32) {
            const int i = i0 + lane;
            const uint32_t hit = __ballot_sync(0xffffffffu, i < cn && b[i] == id);
            if (hit) pos = i0 + __ffs(hit) - 1;
        }
        if (pos < 0) continue;                                       // not in the index (never inserted, or removed before)
        for (int i0 = pos; i0 < cn - 1; i0 += 32) {                  // squeeze it out, keeping the order
            const int i = i0 + lane;
            int32_t v = 0;
            if (i < cn - 1) v = b[i + 1];
            __syncwarp();
            if (i < cn - 1) b[i] = v;
            __syncwarp();
        }
        if (lane == 0) {
            c.child_cnt[idx] = cn - 1;
            if (cn - 1 == 0) c.child_ptr[idx] = 0;
            atomicAdd(removed_entries, 1);
        }
        __threadfence();
        __syncwarp();
    }
}

// recursiveDirDelete for the whole forest: thread per root, post-order walk; a directory node without children is
// unlinked from its parent (roots stay).  Depth <= MAXL + 1.
__global__ void __launch_bounds__(128)
k_collapse_dirs(TreeParams tp, int64_t roots, int32_t* __restrict__ child_ptr, int32_t* __restrict__ child_cnt) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= roots) return;
    constexpr int MAXD = 34;
    int32_t st_node[MAXD];
    int16_t st_slot[MAXD];
    int8_t st_any[MAXD];
    int depth = 0;
    st_node[0] = (int32_t)r; st_slot[0] = 0; st_any[0] = 0;
    while (depth >= 0) {
        const int node = st_node[depth];
        if (st_slot[depth] >= tp.W) {                                // node finished: report to the parent
            const bool empty = !st_any[depth];
            depth--;
            if (depth >= 0) {
                const int64_t pi = (int64_t)st_node[depth] * tp.W + (st_slot[depth] - 1);
                if (empty) { child_cnt[pi] = 0; child_ptr[pi] = 0; }
                else st_any[depth] = 1;
            }
            continue;
        }
        const int slot = st_slot[depth]++;
        const int64_t idx = (int64_t)node * tp.W + slot;
        const int cn = child_cnt[idx];
        if (cn > 0) st_any[depth] = 1;
        else if (cn < 0 && depth + 1 < MAXD) {
            depth++;
            st_node[depth] = child_ptr[idx]; st_slot[depth] = 0; st_any[depth] = 0;
        }
    }
}

// ids per sub-index (numberOfObjectsInEachPartition, RandomDrawTreeMap.java:1572-1573), owned and not removed, summed
// over the tables: the build derives it from its histogram, a put / remove recounts the sub-index ids (L bytes per id)
__global__ void __launch_bounds__(256)
k_occupancy(const uint8_t* __restrict__ pids, const uint8_t* __restrict__ removed, int64_t n, int64_t ld, OwnMask own,
            unsigned long long* __restrict__ hist /* 256 */) {
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int t = blockIdx.y;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int pid = pids[(int64_t)t * ld + i];
        if (own.has(t, pid) && !(removed && removed[i])) atomicAdd(&sh[pid], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

void update_occupancy(dpf_index* h) {
    DevBuf<unsigned long long> hist;
    hist.reserve(256);
    DPF_CUDA(cudaMemsetAsync(hist.p, 0, 256 * sizeof(unsigned long long), h->stream));
    const dim3 grid((unsigned)std::min<int64_t>((h->n + 255) / 256, 1024), h->cfg.L);
    k_occupancy<<<grid, 256, 0, h->stream>>>(h->pids.p, h->removed.p, h->n, h->key_ld, h->own, hist.p); DPF_LAUNCHED();
    unsigned long long occ[256];
    DPF_CUDA(cudaMemcpyAsync(occ, hist.p, sizeof(occ), cudaMemcpyDeviceToHost, h->stream));
    DPF_CUDA(cudaStreamSynchronize(h->stream));
    for (size_t p = 0; p < h->occupancy.size(); ++p) h->occupancy[p] = (double)occ[p] / h->cfg.L;
}

// ---- host side ----------------------------------------------------------------------------------------------------------
// sorted (slot, id) keys of the batch in h->sk64a / sk64b, run heads in h->work0; returns the pointer to the sorted keys
static unsigned long long* locate_and_group(dpf_index* h, const int32_t* ids_dev, int64_t m) {
    const TreeParams tp = h->tp;
    const int L = h->cfg.L;
    cudaStream_t st = h->stream;
    const int64_t total = m * L;
    h->sk64a.reserve((size_t)total);
    h->sk64b.reserve((size_t)total);
    h->work0.reserve((size_t)total + 1);
    k_locate_slots<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ids_dev, m, L, h->keys.p, h->pids.p, h->key_ld, h->n, tp, h->own,
                                                                   h->child_ptr.p, h->child_cnt.p, h->sk64a.p); DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
    unsigned long long *a = h->sk64a.p, *b = h->sk64b.p;
    int idbits = 1, sbits = 1;
    while ((1LL << idbits) < h->n) idbits++;
    while ((1LL << sbits) < (int64_t)h->num_nodes * tp.W) sbits++;
    radix_sort_keys_u64(h, &a, &b, total, 0, idbits);
    radix_sort_keys_u64(h, &a, &b, total, 32, std::min(64, 32 + sbits + 1));   // + 1: the skip key's bits sort last
    DPF_CUDA(cudaMemsetAsync(h->counters.p + 2, 0, sizeof(int32_t), st));
    k_run_heads<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, total, reinterpret_cast<uint32_t*>(h->work0.p), h->counters.p + 2); DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
    return a;
}

static IncCtx make_inc_ctx(dpf_index* h) {
    IncCtx c;
    c.tp = h->tp;
    c.L = h->cfg.L;
    c.keys = h->keys.p;
    c.pids = h->pids.p;
    c.ld = h->key_ld;
    c.table_base = h->table_base.p;
    c.child_ptr = h->child_ptr.p;
    c.child_cnt = h->child_cnt.p;
    c.node_table = h->node_table.p;
    c.ids_sorted = h->ids_sorted.p;
    c.counters = h->counters.p;
    c.node_cap = h->node_cap;
    c.arena_cap = (long long)h->ids_sorted.cap - 64;
    c.roots = (int64_t)h->cfg.L * h->tp.R;
    return c;
}

// inserts the ids [n_old, n_old + m) (keys already stored) one after the other, in ascending order; false = no room
// left in the arena / node array (nothing the caller cannot repair with a rebuild)
bool forest_insert_incremental(dpf_index* h, int64_t n_old, int64_t m) {
    if (!h->leaf_table || h->tp.W > INC_MAXW) return false;
    cudaStream_t st = h->stream;
    const int L = h->cfg.L;
    DevBuf<int32_t> ids;
    ids.reserve((size_t)m);
    std::vector<int32_t> hid((size_t)m);
    for (int64_t i = 0; i < m; ++i) hid[(size_t)i] = (int32_t)(n_old + i);
    DPF_CUDA(cudaMemcpyAsync(ids.p, hid.data(), (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    unsigned long long* sorted = locate_and_group(h, ids.p, m);
    // counters: [0] nodes, [1] overflow, [2] run count, [4..5] arena cursor, [8..9] singleton splits, [10] splits
    int32_t init[12] = {h->num_nodes, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const unsigned long long arena0 = (unsigned long long)h->arena_used;
    memcpy(&init[4], &arena0, sizeof(arena0));
    int32_t keep2 = 0;
    DPF_CUDA(cudaMemcpyAsync(&keep2, h->counters.p + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    DPF_CUDA(cudaStreamSynchronize(st));                                   // `hid` and the run count
    init[2] = keep2;
    DPF_CUDA(cudaMemcpyAsync(h->counters.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    if (keep2 > 0) {
        k_insert_runs<<<(unsigned)((keep2 + INC_WARPS - 1) / INC_WARPS), INC_WARPS * 32, 0, st>>>(
            make_inc_ctx(h), sorted, m * L, reinterpret_cast<uint32_t*>(h->work0.p), h->counters.p + 2); DPF_LAUNCHED();
        DPF_CUDA(cudaGetLastError());
    }
    int32_t out[12];
    DPF_CUDA(cudaMemcpyAsync(out, h->counters.p, sizeof(out), cudaMemcpyDeviceToHost, st));
    DPF_CUDA(cudaStreamSynchronize(st));
    if (out[1]) return false;
    unsigned long long arena1, single;
    memcpy(&arena1, &out[4], sizeof(arena1));
    memcpy(&single, &out[8], sizeof(single));
    h->arena_used = (int64_t)arena1;
    h->num_nodes = out[0];
    h->stats[DPF_STAT_SINGLETON_SPLITS] += (int64_t)single;
    h->stats[DPF_STAT_SPLITS] += out[10];
    h->stats[DPF_STAT_DIR_NODES] = h->num_nodes;
    return true;
}

int64_t forest_remove(dpf_index* h, const int32_t* ids_host, int64_t m) {
    DPF_REQUIRE(h->leaf_table, DPF_ERR_STATE, "remove needs a forest with fewer than 2^32 entries");
    cudaStream_t st = h->stream;
    const int L = h->cfg.L;
    DevBuf<int32_t> ids;
    ids.reserve((size_t)m);
    DPF_CUDA(cudaMemcpyAsync(ids.p, ids_host, (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    unsigned long long* sorted = locate_and_group(h, ids.p, m);
    int32_t nruns = 0;
    DPF_CUDA(cudaMemcpyAsync(&nruns, h->counters.p + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    DPF_CUDA(cudaMemsetAsync(h->counters.p + 3, 0, sizeof(int32_t), st));
    DPF_CUDA(cudaStreamSynchronize(st));
    if (nruns > 0) {
        k_remove_runs<<<(unsigned)((nruns + INC_WARPS - 1) / INC_WARPS), INC_WARPS * 32, 0, st>>>(
            make_inc_ctx(h), sorted, m * L, reinterpret_cast<uint32_t*>(h->work0.p), h->counters.p + 2, h->counters.p + 3); DPF_LAUNCHED();
        const int64_t roots = (int64_t)L * h->tp.R;
        k_collapse_dirs<<<(unsigned)((roots + 127) / 128), 128, 0, st>>>(h->tp, roots, h->child_ptr.p, h->child_cnt.p); DPF_LAUNCHED();
        DPF_CUDA(cudaGetLastError());
    }
    int32_t gone = 0;
    DPF_CUDA(cudaMemcpyAsync(&gone, h->counters.p + 3, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    DPF_CUDA(cudaStreamSynchronize(st));
    return gone;
}

}  // namespace dpf
