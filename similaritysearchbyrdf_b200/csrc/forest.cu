// forest.cu — K3: forest construction as sort + segmented histogram + level-wise split.
//
// Replaces RandomDrawTreeMap.put / putInner (src/main/java/mclab/mapdb/RandomDrawTreeMap.java:1558-1584,
// 1662-1790) for a whole batch: instead of N*L sequential inserts into bitmap-compressed directory records and
// linked-list buckets behind the MapDB engine, every (table, id) entry is sorted by its tree path and the
// directory nodes are produced level by level.  Result = flat arrays in HBM:
//     child_ptr/child_cnt [num_nodes x W]   (dense W-way nodes; roots first: node = t*R + pid*SEG + seg)
//     ids_sorted [sum_t count_t]            (every leaf bucket is one contiguous, id-ascending range)
//
// Bit-exactness with sequential ascending-id insertion: a bucket at level l >= 1
// that was created holding c0 ids splits when its population reaches s = max(c0, T) + 1; the ids present at
// that moment are the s smallest ids of the bucket.  Entries are kept id-ascending inside every node (stable
// sort / stable partition), so "present at the split" = "position < s" and each child's c0 is a histogram of the
// first s entries.  Children never split recursively at the moment of the split (RandomDrawTreeMap.java:1740-1758);
// they split later under the same rule, which is what the next level of the loop evaluates.
#include "common.cuh"

namespace dpf {

struct WorkItem {
    long long start;   // global offset of the segment in ids_sorted
    int32_t cnt;       // entries in the segment
    int32_t c0;        // entries the bucket held when it was created (0 for root-level buckets)
    int32_t node;      // directory node index allocated for this segment
    int32_t table;
};

__device__ __forceinline__ int slot_at(int32_t h, int level, int nb, int mask) {
    return (int)(((uint32_t)h >> (nb * level)) & (uint32_t)mask);
}

// depth-1 code of an entry: ((t*R + pid*SEG + seg) * W + slot(MAXL)).  One CTA counts a chunk of one table's
// entries: in a shared-memory histogram over the table's R*W codes when that fits (the default geometry has 4096),
// flushed with one global atomic per non-empty bin, else with global atomics per entry.
constexpr int CD_THREADS = 256;
constexpr int CD_ITEMS = 64;                       // entries per thread
constexpr int CD_CHUNK = CD_THREADS * CD_ITEMS;    // entries per CTA
constexpr int CD_MAX_SMEM_BINS = 8192;

template <bool SMEM>
__global__ void __launch_bounds__(CD_THREADS)
k_count_depth1(const int32_t* __restrict__ keys, const uint8_t* __restrict__ pids, int64_t n, int64_t ld, TreeParams tp,
               OwnMask own, const uint8_t* __restrict__ removed, int32_t* __restrict__ cnt1) {
    extern __shared__ int32_t hist[];
    const int t = blockIdx.y;
    const int bins = tp.R * tp.W;
    if (SMEM) {
        for (int b = threadIdx.x; b < bins; b += CD_THREADS) hist[b] = 0;
        __syncthreads();
    }
    const int64_t base = (int64_t)blockIdx.x * CD_CHUNK;
    const int64_t end = min(n, base + CD_CHUNK);
    for (int64_t i = base + threadIdx.x; i < end; i += CD_THREADS) {
        const int pid = pids[(int64_t)t * ld + i];
        if (!own.has(t, pid) || (removed && removed[i])) continue;
        const int32_t h = keys[(int64_t)t * ld + i];
        const int seg = tp.seg_bits ? (int)((uint32_t)h >> tp.bucket_bits) : 0;
        const int local = ((pid * tp.SEG + seg) << tp.nb) | slot_at(h, tp.MAXL, tp.nb, tp.W - 1);
        if (SMEM) atomicAdd(&hist[local], 1);
        else atomicAdd(&cnt1[(int64_t)t * bins + local], 1);
    }
    if (SMEM) {
        __syncthreads();
        for (int b = threadIdx.x; b < bins; b += CD_THREADS) {
            const int c = hist[b];
            if (c) atomicAdd(&cnt1[(int64_t)t * bins + b], c);
        }
    }
}

// sort keys of one table group [t0, t0+gt): field = depth-1 code relative to t0; not-owned entries get the
// bit above the field set so that a 1-bit stable pass moves them behind the owned ones
__global__ void __launch_bounds__(256)
k_make_sort_keys(const int32_t* __restrict__ keys, const uint8_t* __restrict__ pids, int64_t n, int64_t ld,
                 TreeParams tp, int t0, OwnMask own, const uint8_t* __restrict__ removed, int field_bits,
                 uint32_t* __restrict__ sk, uint32_t* __restrict__ sv) {
    const int tl = blockIdx.y;
    const int t = t0 + tl;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pid = pids[(int64_t)t * ld + i];
    const int32_t h = keys[(int64_t)t * ld + i];
    const int seg = tp.seg_bits ? (int)((uint32_t)h >> tp.bucket_bits) : 0;
    uint32_t code = (uint32_t)(((tl * tp.R + pid * tp.SEG + seg) << tp.nb) | slot_at(h, tp.MAXL, tp.nb, tp.W - 1));
    if (!own.has(t, pid) || (removed && removed[i])) code |= 1u << field_bits;
    sk[(int64_t)tl * n + i] = code;
    sv[(int64_t)tl * n + i] = (uint32_t)i;
}

// one thread per depth-1 code: empty / leaf bucket / directory (-> work item for the split pass)
__global__ void __launch_bounds__(256)
k_init_depth1(const int32_t* __restrict__ cnt1, const int64_t* __restrict__ start1, const int64_t* __restrict__ table_base,
              int64_t ncodes, TreeParams tp, int32_t* __restrict__ child_ptr, int32_t* __restrict__ child_cnt,
              int32_t* __restrict__ counters /* [0]=node counter, [1]=work count */, WorkItem* __restrict__ work,
              int node_cap, int32_t* __restrict__ node_table) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncodes) return;
    const int cnt = cnt1[c];
    const int t = (int)((c >> tp.nb) / tp.R);
    // depth-1 code == index of the slot in the root node array (root node = c >> nb, slot = c & (W-1))
    if (cnt == 0) { child_ptr[c] = 0; child_cnt[c] = 0; return; }
    if (tp.MAXL >= 1 && cnt >= tp.T + 1) {
        const int node = atomicAdd(&counters[0], 1);
        if (node < node_cap) {
            child_ptr[c] = node;
            child_cnt[c] = -1;
            node_table[node] = t;
            const int w = atomicAdd(&counters[1], 1);
            work[w] = WorkItem{start1[c], cnt, 0, node, t};
            return;
        }
    }
    child_ptr[c] = (int32_t)(start1[c] - table_base[t]);
    child_cnt[c] = cnt;
}

// One CTA per overflowing bucket: stable W-way partition of its ids by the next level's slot, creation of its
// directory node, and the split decision for every child.  `level` is the level of the children.
// Two sizes of CTA: a bucket of many thousand ids (clustered data put 100k+ ids under one depth-1 code) is a long chain of
// rounds for ONE CTA, and the level's time is that chain; buckets above SP_BIG ids go to CTAs of 1024 threads (both kernels
// are launched over all work items of a level and each CTA takes only the items of its class).
constexpr int SP_MAXW = 256;
constexpr int SP_BIG = 8192;    // (2048: 1.13 instead of 1.15 ms)
template <int SP_THREADS>
__global__ void __launch_bounds__(SP_THREADS)
k_split_level(const WorkItem* __restrict__ work, const int32_t* __restrict__ keys, int64_t ld, TreeParams tp, int level,
              const int64_t* __restrict__ table_base, int32_t* __restrict__ ids_sorted, int32_t* __restrict__ tmp,
              int32_t* __restrict__ child_ptr, int32_t* __restrict__ child_cnt, int32_t* __restrict__ counters,
              WorkItem* __restrict__ work_next, int node_cap, unsigned long long* __restrict__ stat_singleton,
              int32_t* __restrict__ node_table, uint8_t* __restrict__ slots /* per entry: its slot at this level */) {
    __shared__ int32_t cntW[SP_MAXW], c0W[SP_MAXW], offW[SP_MAXW], runW[SP_MAXW];
    __shared__ int32_t wcnt[SP_THREADS / 32][SP_MAXW];
    __shared__ int trigger_slot;
    const WorkItem it = work[blockIdx.x];
    if ((it.cnt > SP_BIG) != (SP_THREADS > 256)) return;
    const int W = tp.W, nb = tp.nb, mask = W - 1;
    const int s = max(it.c0, tp.T) + 1;              // population at the moment of the split
    const int32_t* kt = keys + (int64_t)it.table * ld;
    int32_t* seg = ids_sorted + it.start;
    int32_t* out = tmp + it.start;
    uint8_t* sl8 = slots + it.start;                 // the counting pass leaves every entry's slot here: the scatter pass reads
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;   // it back in order instead of gathering the key a second time

    for (int i = tid; i < W; i += SP_THREADS) { cntW[i] = 0; c0W[i] = 0; }
    __syncthreads();
    for (int i = tid; i < it.cnt; i += SP_THREADS) {
        const int sl = slot_at(kt[seg[i]], level, nb, mask);
        sl8[i] = (uint8_t)sl;
        atomicAdd(&cntW[sl], 1);
        if (i < s) atomicAdd(&c0W[sl], 1);
        if (i == s - 1) trigger_slot = sl;           // the id whose insertion overflowed the bucket
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int i = 0; i < W; ++i) { offW[i] = run; runW[i] = run; run += cntW[i]; }
        // quirk Q1 (RandomDrawTreeMap.java:1733-1734): the new id's slot is written with the directory flag; it is
        // only repaired if a redistributed id shares the slot.  We implement the intended flag and count the event.
        if (c0W[trigger_slot] == 1) atomicAdd(stat_singleton, 1ULL);
    }
    __syncthreads();
    // stable scatter, 256 entries per round in ascending position
    const uint32_t lt = (1u << lane) - 1u;
    for (int base = 0; base < it.cnt; base += SP_THREADS) {
        for (int i = tid; i < (SP_THREADS / 32) * W; i += SP_THREADS) (&wcnt[0][0])[(i / W) * SP_MAXW + (i % W)] = 0;
        __syncthreads();
        const int i = base + tid;
        const bool ok = i < it.cnt;
        int32_t id = 0;
        uint32_t sl = 0xFFFFFFFFu;
        if (ok) { id = seg[i]; sl = sl8[i]; }
        const uint32_t peers = __match_any_sync(0xffffffffu, sl);
        const int rank_in_warp = __popc(peers & lt);
        if (ok && rank_in_warp == 0) wcnt[w][sl] = __popc(peers);
        __syncthreads();
        int pos = 0;
        if (ok) {
            pos = runW[sl] + rank_in_warp;
            for (int ww = 0; ww < w; ++ww) pos += wcnt[ww][sl];
        }
        __syncthreads();
        if (ok) out[pos] = id;
        for (int sidx = tid; sidx < W; sidx += SP_THREADS) {
            int add = 0;
            for (int ww = 0; ww < SP_THREADS / 32; ++ww) add += wcnt[ww][sidx];
            runW[sidx] += add;
        }
        __syncthreads();
    }
    for (int i = tid; i < it.cnt; i += SP_THREADS) seg[i] = out[i];
    // children
    for (int sl = tid; sl < W; sl += SP_THREADS) {
        const int64_t ci = (int64_t)it.node * W + sl;
        const int cnt = cntW[sl];
        if (cnt == 0) { child_ptr[ci] = 0; child_cnt[ci] = 0; continue; }
        const int c0 = c0W[sl];
        bool made_dir = false;
        if (level >= 1 && cnt >= max(c0, tp.T) + 1) {
            const int node = atomicAdd(&counters[0], 1);
            if (node < node_cap) {
                child_ptr[ci] = node;
                child_cnt[ci] = -1;
                node_table[node] = it.table;
                const int wi = atomicAdd(&counters[2], 1);
                work_next[wi] = WorkItem{it.start + offW[sl], cnt, c0, node, it.table};
                made_dir = true;
            }
        }
        if (!made_dir) {
            child_ptr[ci] = (int32_t)(it.start + offW[sl] - table_base[it.table]);
            child_cnt[ci] = cnt;
        }
    }
}

// ---- leaf table: a dense number for every leaf bucket, its range in ids_sorted ---------------------------------
// (the bucket-major re-rank groups the (bucket, query) pairs of a batch with a counting sort over these numbers)
__global__ void __launch_bounds__(256)
k_leaf_flags(const int32_t* __restrict__ child_cnt, int64_t nslots, uint32_t* __restrict__ child_leaf) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= nslots) child_leaf[i] = (i < nslots && child_cnt[i] > 0) ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
k_leaf_table(const int32_t* __restrict__ child_ptr, const int32_t* __restrict__ child_cnt, const uint32_t* __restrict__ child_leaf,
             const int32_t* __restrict__ node_table, const int64_t* __restrict__ table_base, int64_t nslots, int W, int64_t roots,
             int R, uint32_t* __restrict__ leaf_pos, int32_t* __restrict__ leaf_len) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nslots) return;
    const int cnt = child_cnt[i];
    if (cnt <= 0) return;
    const int64_t node = i / W;
    const int t = node < roots ? (int)(node / R) : node_table[node];
    const uint32_t leaf = child_leaf[i];
    leaf_pos[leaf] = (uint32_t)(table_base[t] + child_ptr[i]);
    leaf_len[leaf] = cnt;
}

__global__ void __launch_bounds__(256)
k_leaf_shape(const int32_t* __restrict__ leaf_len, int64_t nleaves, unsigned long long* __restrict__ out /* [0] max len, [1] tiles */) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long len = i < nleaves ? (unsigned long long)leaf_len[i] : 0ULL, tiles = (len + 127) / 128;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
        tiles += __shfl_xor_sync(0xffffffffu, tiles, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicMax(&out[0], len); atomicAdd(&out[1], tiles); }
}

void rebuild_leaf_table(dpf_index* h);
static void build_leaf_table(dpf_index* h) { rebuild_leaf_table(h); }
void rebuild_leaf_table(dpf_index* h) {
    const TreeParams tp = h->tp;
    cudaStream_t st = h->stream;
    h->leaf_table = false;
    h->num_leaves = 0;
    if ((int64_t)h->ids_sorted.cap >= (1LL << 32)) return;   // positions are 32-bit: such an index re-ranks row-major
    const int64_t nslots = (int64_t)h->num_nodes * tp.W;
    h->child_leaf.reserve((size_t)nslots + 1);
    k_leaf_flags<<<(unsigned)((nslots + 256) / 256), 256, 0, st>>>(h->child_cnt.p, nslots, h->child_leaf.p); DPF_LAUNCHED();
    exclusive_scan_u32(h, h->child_leaf.p, nslots + 1);
    uint32_t nl = 0;
    DPF_CUDA(cudaMemcpyAsync(&nl, h->child_leaf.p + nslots, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    DPF_CUDA(cudaStreamSynchronize(st));
    h->leaf_pos.reserve((size_t)nl + 1);
    h->leaf_len.reserve((size_t)nl + 1);
    h->leaf_cnt.reserve((size_t)nl + 1);
    h->leaf_off.reserve((size_t)nl + 2);
    h->leaf_unit_off.reserve((size_t)nl + 2);
    h->leaf_unit_off_tc.reserve((size_t)nl + 2);
    DPF_CUDA(cudaMemsetAsync(h->leaf_cnt.p, 0, ((size_t)nl + 1) * sizeof(uint32_t), st));
    if (nslots > 0) {
        k_leaf_table<<<(unsigned)((nslots + 255) / 256), 256, 0, st>>>(h->child_ptr.p, h->child_cnt.p, h->child_leaf.p, h->node_table.p,
                                                                       h->table_base.p, nslots, tp.W, (int64_t)h->cfg.L * tp.R, tp.R,
                                                                       h->leaf_pos.p, h->leaf_len.p); DPF_LAUNCHED();
    }
    DPF_CUDA(cudaGetLastError());
    {   // shape of the leaves: sizes the tcgen05 kernel's unit records without reading anything back at query time
        unsigned long long* shape = reinterpret_cast<unsigned long long*>(h->counters.p + 8);
        DPF_CUDA(cudaMemsetAsync(shape, 0, 2 * sizeof(unsigned long long), st));
        if (nl > 0) { k_leaf_shape<<<(unsigned)((nl + 255) / 256), 256, 0, st>>>(h->leaf_len.p, nl, shape); DPF_LAUNCHED(); }
        unsigned long long hs[2] = {0, 0};
        DPF_CUDA(cudaMemcpyAsync(hs, shape, sizeof(hs), cudaMemcpyDeviceToHost, st));
        DPF_CUDA(cudaStreamSynchronize(st));
        h->max_leaf_len = (int64_t)hs[0];
        h->total_leaf_tiles = (int64_t)hs[1];
    }
    h->num_leaves = (int32_t)nl;
    h->leaf_table = true;
}

ForestView forest_view(const dpf_index* h) {
    return ForestView{h->child_ptr.p, h->child_cnt.p, h->ids_sorted.p, h->table_base.p, h->num_nodes,
                      h->child_leaf.p, h->leaf_pos.p, h->leaf_len.p, h->num_leaves};
}

void build_forest(dpf_index* h) {
    const TreeParams tp = h->tp;
    const int L = h->cfg.L;
    const int64_t n = h->n, ld = h->key_ld;
    const int world = h->cfg.world > 1 ? h->cfg.world : 1;
    const OwnMask own = h->own;
    DPF_REQUIRE(tp.W <= SP_MAXW, DPF_ERR_INVALID, "dirNodeSize above 256 is not supported");
    const int64_t ncodes = (int64_t)L * tp.R * tp.W;
    DPF_REQUIRE(ncodes < (1LL << 30), DPF_ERR_INVALID, "L * 2^pb * SEG * dirNodeSize too large");
    cudaStream_t st = h->stream;

    // ---- depth-1 histogram, per-table counts ---------------------------------------------------------------
    DevBuf<int32_t> cnt1;
    DevBuf<int64_t> start1;
    cnt1.reserve(ncodes);
    start1.reserve(ncodes + 1);
    DPF_CUDA(cudaMemsetAsync(cnt1.p, 0, ncodes * sizeof(int32_t), st));
    {
        StageTimer tm(h, DPF_T_SORT);
        if (n > 0) {
            const dim3 grid((unsigned)((n + CD_CHUNK - 1) / CD_CHUNK), L);
            const int bins = tp.R * tp.W;
            if (bins <= CD_MAX_SMEM_BINS)
                k_count_depth1<true><<<grid, CD_THREADS, bins * sizeof(int32_t), st>>>(h->keys.p, h->pids.p, n, ld, tp, own,
                                                                                      h->removed.p, cnt1.p);
            else
                k_count_depth1<false><<<grid, CD_THREADS, 0, st>>>(h->keys.p, h->pids.p, n, ld, tp, own, h->removed.p, cnt1.p);
            DPF_LAUNCHED();
            DPF_CUDA(cudaGetLastError());
        }
        exclusive_scan_i64(h, cnt1.p, start1.p, ncodes);
    }
    std::vector<int32_t> hcnt(ncodes);
    DPF_CUDA(cudaMemcpyAsync(hcnt.data(), cnt1.p, ncodes * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    DPF_CUDA(cudaStreamSynchronize(st));
    h->h_table_base.assign(L + 1, 0);
    const int64_t per_table = (int64_t)tp.R * tp.W;
    for (int t = 0; t < L; ++t) {
        int64_t c = 0;
        for (int64_t j = 0; j < per_table; ++j) c += hcnt[t * per_table + j];
        h->h_table_base[t + 1] = h->h_table_base[t] + c;
    }
    const int64_t E = h->h_table_base[L];
    h->table_base.reserve(L + 1);
    DPF_CUDA(cudaMemcpyAsync(h->table_base.p, h->h_table_base.data(), (L + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    // sub-index occupancy (numberOfObjectsInEachPartition, RandomDrawTreeMap.java:1572-1573; averaged over tables as
    // getDtAndHtNumDistribution does, DensevectorRDFInit.scala:515-530)
    h->occupancy.assign(1 << tp.pb, 0.0);
    for (int t = 0; t < L; ++t)
        for (int p = 0; p < (1 << tp.pb); ++p) {
            int64_t c = 0;
            for (int64_t j = 0; j < (int64_t)tp.SEG * tp.W; ++j) c += hcnt[t * per_table + (int64_t)p * tp.SEG * tp.W + j];
            h->occupancy[p] += (double)c / L;
        }

    // head-room behind the tables: buckets that grow by an incremental put move there (incremental.cu); + slack: the
    // re-rank copies 16-byte-aligned id windows
    h->ids_sorted.reserve((size_t)std::max<int64_t>(E, 1) + (size_t)(E / 4) + (1 << 20) + 64);
    h->arena_used = E;
    DevBuf<int32_t> tmp;
    tmp.reserve((size_t)std::max<int64_t>(E, 1));
    DevBuf<uint8_t> slot_buf;
    slot_buf.reserve((size_t)std::max<int64_t>(E, 1));

    // ---- stable sort by (table, root, slot(MAXL)), one table group at a time -------------------------------
    {
        StageTimer tm(h, DPF_T_SORT);
        int root_bits = 0;
        while ((1 << root_bits) < tp.R) root_bits++;
        const int64_t max_items = 1LL << 30;
        int gmax = (int)std::max<int64_t>(1, std::min<int64_t>(L, max_items / std::max<int64_t>(n, 1)));
        for (int t0 = 0; t0 < L && n > 0; t0 += gmax) {
            const int gt = std::min(gmax, L - t0);
            int tbits = 0;
            while ((1 << tbits) < gt) tbits++;
            const int field_bits = tbits + root_bits + tp.nb;
            DPF_REQUIRE(field_bits < 32, DPF_ERR_INVALID, "tree path field exceeds 31 bits");
            const int64_t items = (int64_t)gt * n;
            h->sk0.reserve(items); h->sk1.reserve(items); h->sv0.reserve(items); h->sv1.reserve(items);
            const dim3 grid((unsigned)((n + 255) / 256), gt);
            k_make_sort_keys<<<grid, 256, 0, st>>>(h->keys.p, h->pids.p, n, ld, tp, t0, own, h->removed.p, field_bits, h->sk0.p,
                                                   h->sv0.p); DPF_LAUNCHED();
            DPF_CUDA(cudaGetLastError());
            uint32_t *k0 = h->sk0.p, *k1 = h->sk1.p, *v0 = h->sv0.p, *v1 = h->sv1.p;
            const int64_t owned = h->h_table_base[t0 + gt] - h->h_table_base[t0];
            if (world > 1 || h->n_removed > 0) radix_sort_pairs_u32(h, &k0, &k1, &v0, &v1, items, field_bits, field_bits + 1);
            radix_sort_pairs_u32(h, &k0, &k1, &v0, &v1, owned, 0, field_bits);
            if (owned > 0)
                DPF_CUDA(cudaMemcpyAsync(h->ids_sorted.p + h->h_table_base[t0], v0, owned * sizeof(int32_t),
                                         cudaMemcpyDeviceToDevice, st));
        }
    }

    // ---- nodes ----------------------------------------------------------------------------------------------
    const int64_t roots = (int64_t)L * tp.R;
    int64_t split_bound = tp.MAXL >= 1 ? (int64_t)tp.MAXL * (E / (tp.T + 1)) : 0;
    const int64_t node_cap64 = roots + split_bound + split_bound / 4 + 4096;   // head-room for incremental puts
    DPF_REQUIRE(node_cap64 * tp.W < (1LL << 31), DPF_ERR_NOMEM, "forest would need more than 2^31 child slots");
    h->node_cap = (int32_t)node_cap64;
    h->child_ptr.reserve((size_t)node_cap64 * tp.W);
    h->child_cnt.reserve((size_t)node_cap64 * tp.W);
    h->node_table.reserve((size_t)node_cap64);
    h->counters.reserve(CTR_COUNT);
    int32_t init_counters[4] = {(int32_t)roots, 0, 0, 0};
    DPF_CUDA(cudaMemcpyAsync(h->counters.p, init_counters, sizeof(init_counters), cudaMemcpyHostToDevice, st));
    unsigned long long* stat_dev = reinterpret_cast<unsigned long long*>(h->counters.p + 8);
    DPF_CUDA(cudaMemsetAsync(stat_dev, 0, sizeof(unsigned long long), st));
    const int64_t work_cap = std::max<int64_t>(1, E / (tp.T + 1) + 1);
    DevBuf<WorkItem> workA, workB;
    workA.reserve(work_cap);
    workB.reserve(work_cap);
    int64_t total_splits = 0;
    {
        StageTimer tm(h, DPF_T_SPLIT);
        k_init_depth1<<<(unsigned)((ncodes + 255) / 256), 256, 0, st>>>(cnt1.p, start1.p, h->table_base.p, ncodes, tp,
                                                                         h->child_ptr.p, h->child_cnt.p, h->counters.p,
                                                                         workA.p, h->node_cap, h->node_table.p); DPF_LAUNCHED();
        DPF_CUDA(cudaGetLastError());
        int32_t hc[4];
        DPF_CUDA(cudaMemcpyAsync(hc, h->counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
        DPF_CUDA(cudaStreamSynchronize(st));
        int nwork = hc[1];
        WorkItem* cur = workA.p;
        WorkItem* nxt = workB.p;
        for (int level = tp.MAXL - 1; level >= 0 && nwork > 0; --level) {
            total_splits += nwork;
            DPF_CUDA(cudaMemsetAsync(h->counters.p + 2, 0, sizeof(int32_t), st));
            k_split_level<1024><<<nwork, 1024, 0, st>>>(cur, h->keys.p, ld, tp, level, h->table_base.p, h->ids_sorted.p,
                                                        tmp.p, h->child_ptr.p, h->child_cnt.p, h->counters.p, nxt,
                                                        h->node_cap, stat_dev, h->node_table.p, slot_buf.p); DPF_LAUNCHED();
            k_split_level<256><<<nwork, 256, 0, st>>>(cur, h->keys.p, ld, tp, level, h->table_base.p, h->ids_sorted.p,
                                                      tmp.p, h->child_ptr.p, h->child_cnt.p, h->counters.p, nxt,
                                                      h->node_cap, stat_dev, h->node_table.p, slot_buf.p); DPF_LAUNCHED();
            DPF_CUDA(cudaGetLastError());
            DPF_CUDA(cudaMemcpyAsync(hc, h->counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
            DPF_CUDA(cudaStreamSynchronize(st));
            nwork = hc[2];
            std::swap(cur, nxt);
        }
        DPF_REQUIRE(hc[0] <= h->node_cap, DPF_ERR_NOMEM, "directory node capacity exceeded");
        h->num_nodes = hc[0];
    }
    unsigned long long single = 0;
    DPF_CUDA(cudaMemcpyAsync(&single, stat_dev, sizeof(single), cudaMemcpyDeviceToHost, st));
    DPF_CUDA(cudaStreamSynchronize(st));
    h->stats[DPF_STAT_SINGLETON_SPLITS] = (int64_t)single;
    h->stats[DPF_STAT_SPLITS] = total_splits;
    h->stats[DPF_STAT_DIR_NODES] = h->num_nodes;
    build_leaf_table(h);
    h->fitted = true;
}

}  // namespace dpf
