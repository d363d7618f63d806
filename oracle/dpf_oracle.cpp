/*
 * dpf_oracle.cpp — CPU parity oracle for the Dynamic Partition Forest hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see dpf_oracle.h).  From-scratch restatement of the reference *behaviour*;
 * every function cites the reference file:line it follows (paths relative to /root/reference/).
 * Compile with -ffp-contract=off: the JVM never contracts a*b+c into an FMA, so every product and every sum
 * below is rounded separately.
 *
 * Parity pinning status: pinned on the reference's own known-answer tests (tests/test_oracle_kat.py);
 * end-to-end outputs on data are "parity unpinned" by the reference (no datasets, unseeded functions, no JVM
 * here), so this literal sequential restatement is the definition the CUDA path is checked against.
 */
#include "dpf_oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------------------
// java.lang.Integer helpers
// ---------------------------------------------------------------------------------------------------------
inline int java_nlz(int32_t v) { return v == 0 ? 32 : __builtin_clz((uint32_t)v); }
inline int java_bitcount(int32_t v) { return __builtin_popcount((uint32_t)v); }
inline int32_t lsr(int32_t v, int s) { return (int32_t)((uint32_t)v >> (s & 31)); }   // Java >>>
inline int32_t lsl(int32_t v, int s) { return (int32_t)((uint32_t)v << (s & 31)); }   // Java <<

// Java (int) cast of a double: NaN -> 0, saturating, truncation toward zero (JLS 5.1.3).
inline int32_t java_d2i(double v) {
    if (v != v) return 0;
    if (v >= 2147483647.0) return INT32_MAX;
    if (v <= -2147483648.0) return INT32_MIN;
    return (int32_t)v;
}

// java.util.Random: 48-bit LCG.
struct JavaRandom {
    uint64_t seed;
    explicit JavaRandom(int64_t s) { seed = ((uint64_t)s ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1); }
    int32_t next(int bits) {
        seed = (seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
        return (int32_t)((int64_t)seed >> (48 - bits));
    }
    int32_t nextInt(int32_t n) {
        if ((n & -n) == n) return (int32_t)(((int64_t)n * (int64_t)next(31)) >> 31);
        int32_t bits, val;
        do {
            bits = next(31);
            val = bits % n;
        } while ((int32_t)((uint32_t)bits - (uint32_t)val + (uint32_t)(n - 1)) < 0);
        return val;
    }
};

// ---------------------------------------------------------------------------------------------------------
// SimilarityCalculator.scala
// ---------------------------------------------------------------------------------------------------------
// :29-49  `pair.foreach(x => similarity = similarity + x._1*x._2)` — left to right, mul then add.
inline double dot_dense(const double* a, const double* x, int d) {
    double s = 0.0;
    for (int j = 0; j < d; ++j) s = s + a[j] * x[j];
    return s;
}

// :9-27  BitSet AND, iterate set bits ascending, `similarity += mapA(i) * mapB(i)`.
inline double dot_sparse(const int32_t* ia, const double* va, int na, const int32_t* ib, const double* vb, int nb) {
    double s = 0.0;
    int p = 0, q = 0;
    while (p < na && q < nb) {
        if (ia[p] == ib[q]) { s += va[p] * vb[q]; ++p; ++q; }
        else if (ia[p] < ib[q]) ++p;
        else ++q;
    }
    return s;
}

// ---------------------------------------------------------------------------------------------------------
// hash families
// ---------------------------------------------------------------------------------------------------------
// AngleHashFamily.scala:184  sign(x) = if (x <= 0) 0 else 1  (NaN <= 0 is false in Scala => 1? no:
// `input <= 0` with NaN is false => else branch => 1).  NB: SURVEY A.2 says NaN -> 0; the literal
// code gives 1 for NaN.  The oracle follows the code.
inline int angle_sign(double v) { return (v <= 0) ? 0 : 1; }

// AngleHashFamily.scala:187-195 / :204-219
inline int32_t angle_key_from_dots(const double* dots, int k) {
    int32_t r = 0;
    for (int b = 0; b < k; ++b) r = lsl(r, 1) | angle_sign(dots[b]);
    return k >= 32 ? r : lsl(r, 32 - k);  // Java `result << (32 - chainSize)`; shift by 0 when k == 32
}

// PStableHashFamily.scala:122-143: ints -> big-endian bytes -> ByteArrayWrapper.hashCode = Arrays.hashCode(byte[])
inline int32_t pstable_key_from_dots(const double* dots, const double* b, const int32_t* w, int k) {
    uint32_t h = 1;
    for (int i = 0; i < k; ++i) {
        int32_t q = java_d2i((dots[i] + b[i]) / (double)w[i]);
        for (int s = 24; s >= 0; s -= 8) {
            int8_t byte = (int8_t)((uint32_t)q >> s);
            h = 31u * h + (uint32_t)(int32_t)byte;
        }
    }
    return (int32_t)h;
}

// Sampling.scala:6-11: random.shuffle(0..31) with scala.util.Random(88387) (Scala 2.10 Fisher-Yates:
// for n <- len to 2 by -1 { k = nextInt(n); swap(n-1, k) }).
struct SamplingIndex {
    int32_t sigma[32];
    SamplingIndex() {
        for (int i = 0; i < 32; ++i) sigma[i] = i;
        JavaRandom rnd(88387);
        for (int n = 32; n >= 2; --n) {
            int k = rnd.nextInt(n);
            std::swap(sigma[n - 1], sigma[k]);
        }
    }
};
const SamplingIndex& sampling_index() { static SamplingIndex s; return s; }

// Sampling.scala:32-39
inline int32_t sampling_one_key(int32_t key) {
    const int32_t* sg = sampling_index().sigma;
    int32_t tmp = 0;
    for (int j = 0; j < 32; ++j) tmp += lsl(lsr(key, sg[j]) & 1, 31 - j);
    return tmp;
}

// significantBits.scala:11-67 with numOfBits = Array(6,4,2,1)
inline int32_t continue_bits_count(int32_t key) {
    const int nob[4] = {6, 4, 2, 1};
    int32_t first4 = lsr(key, 28);
    int32_t arr[4] = {0, 0, 0, 0};
    int count = 0;
    auto flush = [&](int c) {
        if (c >= nob[0]) { arr[0]++; arr[1]++; arr[2]++; arr[3]++; }
        else if (c >= nob[1]) { arr[1]++; arr[2]++; arr[3]++; }
        else if (c >= nob[2]) { arr[2]++; arr[3]++; }
        else if (c >= nob[3]) { arr[3]++; }
    };
    for (int i = 0; i < 28; ++i) {
        if ((lsr(key, i) & 1) == 1) {
            count += 1;
            if (i == 27) { flush(count); count = 0; }
        } else {
            flush(count);
            count = 0;
        }
    }
    // `for (i <- newIndexArray.reverse.indices) tmp += newIndexArray.reverse(i) << ((3 - i)*7)`
    int32_t tmp = 0;
    for (int i = 0; i < 4; ++i) tmp += lsl(arr[3 - i], (3 - i) * 7);
    tmp += lsl(first4, 28);
    return tmp;
}

// significantBits.scala:100-110
inline double angle_distance(int32_t key) {
    // v1 . base = popcount of low 28 bits; |base| = sqrt(28); |v1| = sqrt(popcount)
    double dotp = 0.0, n1 = 0.0, nb = 0.0;
    for (int i = 0; i < 28; ++i) {
        int bit = lsr(key, i) & 1;
        dotp += (double)bit * 1.0;
        n1 += (double)bit * (double)bit;
        nb += 1.0;
    }
    double angle = std::acos(dotp / (std::sqrt(nb) * std::sqrt(n1)));
    return angle * 360 / 2 / M_PI;
}
// significantBits.scala:113-127
inline int32_t angle_new_method(int32_t key) {
    const double metric[9] = {16.0, 25.0, 33.0, 39.0, 46.0, 52.0, 58.0, 66.0, 72.0};
    int index = 0;
    while (index < 9 && (angle_distance(key) > metric[index])) index += 1;
    const int32_t mask = 0x7f;
    int32_t first4 = lsr(key, 28) & mask, first7 = lsr(key, 21) & mask;
    int32_t three7 = lsr(key, 7) & mask, last7 = key & mask;
    return last7 + lsl(three7, 7) + lsl(index, 14) + lsl(first7, 21) + lsl(first4, 28);
}

// LSH.scala:135-166 key transform selected by mclab.lsh.typeOfIndex
inline int32_t apply_transform(int32_t key, int kind) {
    switch (kind) {
        case 1: return sampling_one_key(key);
        case 2: return continue_bits_count(key);
        case 3: return angle_new_method(key);
        default: return key;
    }
}

// Partitioner.scala:40-64: bits of h (bit i -> coordinate i) as a 0/1 sparse vector; chain of pb angle
// functions of the table's private LSH; `calculateIndex(v, 0)(0) >>> (32 - partitionBits)`.
//
// The partitioner's LSH is built from the main configuration with vectorDim = 32, tableNum = 1 and chainLength = pb
// (DensevectorRDFInit.scala:63-70), so with mclab.lsh.name = pStable its chain is a PStableHashChain: the pb sums are
// quantised ((sum + b) / w).toInt and packed by Arrays.hashCode like a pStable key (PStableHashFamily.scala:155-177);
// pb_b / pb_w are that chain's b and w, NULL for the angle family.
inline int32_t partition_id(int32_t h, const double* Ap_t, int pb, int key_transform, const double* pb_b = nullptr,
                            const int32_t* pb_w = nullptr) {
    if (pb == 0) return 0;
    double sums[32];
    for (int j = 0; j < pb; ++j) {
        const double* a = Ap_t + (size_t)j * 32;
        double s = 0.0;
        for (int i = 0; i < 32; ++i)
            if ((lsr(h, i) & 1) != 0 && a[i] != 0.0) s += a[i] * 1.0;   // sparse . sparse, ascending i
        sums[j] = s;
    }
    int32_t key;
    if (pb_b) {
        key = pstable_key_from_dots(sums, pb_b, pb_w, pb);
    } else {
        int32_t r = 0;
        for (int j = 0; j < pb; ++j) r = lsl(r, 1) | angle_sign(sums[j]);
        key = pb >= 32 ? r : lsl(r, 32 - pb);
    }
    key = apply_transform(key, key_transform);
    return lsr(key, 32 - pb);
}

// RandomDrawTreeMap.java:435-465
struct TreeParams { int SEG, nb, mask, MAXL, W, bucket_bits; };
inline int ilog2_like_java(int n) {
    // (int)(Math.log(n) / Math.log(2)); exact for the powers of two the reference uses; guard the ulp.
    int r = (int)(std::log((double)n) / std::log(2.0));
    if ((1 << (r + 1)) <= n) r += 1;
    if ((1 << r) > n && r > 0) r -= 1;
    return r;
}
inline TreeParams tree_params(int bucket_bits, int dir_node_size, int chain_length) {
    TreeParams tp;
    tp.bucket_bits = bucket_bits;
    tp.SEG = 1 << (32 - bucket_bits);
    tp.nb = ilog2_like_java(dir_node_size);
    tp.mask = (1 << tp.nb) - 1;
    tp.MAXL = (chain_length - (32 - bucket_bits)) / tp.nb - 1;
    tp.W = dir_node_size;
    return tp;
}

// ---------------------------------------------------------------------------------------------------------
// forest (semantics of RandomDrawTreeMap; storage engine replaced by plain vectors)
// ---------------------------------------------------------------------------------------------------------
// child encoding: 0 empty; >0 bucket index+1; <0 -(dir index+1)
struct Table {
    std::vector<std::vector<int32_t>> dirs;     // each dir: W children
    std::vector<std::vector<int32_t>> buckets;  // ids in insertion order (reference prepends; order is not observable)
    std::vector<int32_t> keys;                  // key of id in this table (reference re-hashes: RandomDrawTreeMap.java:1747)
    std::vector<int32_t> pids;
    std::vector<int64_t> occupancy;             // numberOfObjectsInEachPartition (RandomDrawTreeMap.java:1572-1573)
    int64_t singleton_splits = 0, splits = 0;
};

}  // namespace

struct dpfo {
    dpfo_cfg cfg;
    TreeParams tp;
    std::vector<double> A;            // P x d
    std::vector<double> Anz;          // unused
    std::vector<int32_t> chain;       // L x k
    std::vector<double> fb;           // P
    std::vector<int32_t> fw;          // P
    std::vector<double> Ap;           // L x pb x 32
    std::vector<double> pb_b;         // L x pb: b of the partitioner chains (pStable family only)
    std::vector<int32_t> pb_w;        // L x pb
    const double* part_b(int t) const { return pb_b.empty() ? nullptr : &pb_b[(size_t)t * cfg.pb]; }
    const int32_t* part_w(int t) const { return pb_w.empty() ? nullptr : &pb_w[(size_t)t * cfg.pb]; }
    std::vector<Table> tables;
    bool dense = true;
    int64_t n = 0;
    std::vector<double> X;            // dense store n x d
    std::vector<int64_t> sp_ptr;      // csr store
    std::vector<int32_t> sp_idx;
    std::vector<double> sp_val;
    // last candidate result
    std::vector<int64_t> cand_off;
    std::vector<int32_t> cand_ids;
    std::atomic<int64_t> nlz_gt28{0};
    std::vector<uint8_t> owned;       // explicit ownership of (table, sub-index) cells, L x 2^pb (dpfo_set_owned[_cells]);
                                      // empty = p % world == rank in every table
    bool owns(int t, int p) const {
        if (!owned.empty()) return owned[((size_t)t << cfg.pb) + (size_t)p] != 0;
        return cfg.world <= 1 || (p % cfg.world) == cfg.rank;
    }
};

namespace {

int resolve_threads(int nthreads) {
    if (nthreads > 0) return nthreads;
    unsigned h = std::thread::hardware_concurrency();
    return h ? (int)h : 1;
}

template <class F>
void parallel_for(int64_t n, int nthreads, F f) {
    nthreads = (int)std::min<int64_t>(resolve_threads(nthreads), std::max<int64_t>(n, 1));
    if (nthreads <= 1) { f(0, n, 0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        int64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
        th.emplace_back([=]() { f(lo, hi, t); });
    }
    for (auto& x : th) x.join();
}

// keys of all L tables for one vector given its P distinct projections
// (LSH.scala:135-166 + AngleHashFamily.scala:187-195; permuted tables reuse the same functions,
// AngleHashFamily.scala:143-146, so each distinct dot is evaluated once — identical values).
inline void keys_from_dots(const dpfo* o, const double* dots, int32_t* keys_L, double* scratch) {
    const int L = o->cfg.L, k = o->cfg.k;
    for (int t = 0; t < L; ++t) {
        const int32_t* ch = &o->chain[(size_t)t * k];
        int32_t key;
        if (o->cfg.family_kind == 0) {
            for (int b = 0; b < k; ++b) scratch[b] = dots[ch[b]];
            key = angle_key_from_dots(scratch, k);
        } else {
            std::vector<double> bb(k);
            std::vector<int32_t> ww(k);
            for (int b = 0; b < k; ++b) { scratch[b] = dots[ch[b]]; bb[b] = o->fb[ch[b]]; ww[b] = o->fw[ch[b]]; }
            key = pstable_key_from_dots(scratch, bb.data(), ww.data(), k);
        }
        keys_L[t] = apply_transform(key, o->cfg.key_transform);
    }
}

// P dot products of one dense vector, each accumulated strictly left-to-right (SimilarityCalculator.scala:40-49);
// 8 functions are advanced together only to overlap their independent dependency chains.
inline void project_dense(const dpfo* o, const double* x, double* dots) {
    const int P = o->cfg.P, d = o->cfg.d;
    int p = 0;
    for (; p + 8 <= P; p += 8) {
        const double* a = &o->A[(size_t)p * d];
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
        for (int j = 0; j < d; ++j) {
            const double xv = x[j];
            s0 = s0 + a[j] * xv;             s1 = s1 + a[(size_t)d + j] * xv;
            s2 = s2 + a[(size_t)2 * d + j] * xv; s3 = s3 + a[(size_t)3 * d + j] * xv;
            s4 = s4 + a[(size_t)4 * d + j] * xv; s5 = s5 + a[(size_t)5 * d + j] * xv;
            s6 = s6 + a[(size_t)6 * d + j] * xv; s7 = s7 + a[(size_t)7 * d + j] * xv;
        }
        dots[p] = s0; dots[p + 1] = s1; dots[p + 2] = s2; dots[p + 3] = s3;
        dots[p + 4] = s4; dots[p + 5] = s5; dots[p + 6] = s6; dots[p + 7] = s7;
    }
    for (; p < P; ++p) dots[p] = dot_dense(&o->A[(size_t)p * d], x, d);
}

// sparse data vector against fully dense functions: ascending index over supp(a) ∩ supp(x)
// (SimilarityCalculator.scala:9-27); a zero coefficient of the function is outside its support.
inline void project_csr(const dpfo* o, const int32_t* idx, const double* val, int nnz, double* dots) {
    const int P = o->cfg.P, d = o->cfg.d;
    for (int p = 0; p < P; ++p) {
        const double* a = &o->A[(size_t)p * d];
        double s = 0.0;
        for (int j = 0; j < nnz; ++j) {
            const double av = a[idx[j]];
            if (av != 0.0) s += av * val[j];
        }
        dots[p] = s;
    }
}

inline int slot_of(const TreeParams& tp, int32_t h, int level) { return lsr(h, tp.nb * level) & tp.mask; }
inline int seg_of(const TreeParams& tp, int32_t h) { return tp.bucket_bits >= 32 ? 0 : lsr(h, tp.bucket_bits); }

void table_init(const dpfo* o, Table& T) {
    const int roots = (1 << o->cfg.pb) * o->tp.SEG;       // initPartition: SEG empty root dirs per sub-index
    T.dirs.assign(roots, std::vector<int32_t>(o->tp.W, 0)); // (RandomDrawTreeMap.java:1430-1459, 2775-2785)
    T.buckets.clear();
    T.keys.clear();
    T.pids.clear();
    T.occupancy.assign(1 << o->cfg.pb, 0);
}

// RandomDrawTreeMap.java:1558-1584 (put) + :1662-1790 (putInner), intended bucket-flag semantics (quirk Q1).
void table_insert(const dpfo* o, Table& T, int32_t id, int32_t h, int32_t pid) {
    const TreeParams& tp = o->tp;
    const int Tov = o->cfg.bucket_overflow;
    T.occupancy[pid]++;
    int32_t dir = pid * tp.SEG + seg_of(tp, h);
    int level = tp.MAXL;
    while (true) {
        const int slot = slot_of(tp, h, level);
        const int32_t e = T.dirs[dir][slot];
        if (e < 0) { dir = -e - 1; level--; continue; }     // d-node: descend
        int c = 0;
        if (e > 0) {
            std::vector<int32_t>& b = T.buckets[e - 1];
            for (int32_t y : b) if (y == id) return;         // same key: value replaced, structure unchanged
            c = (int)b.size();                               // bucketConflictCost = length of the list
        }
        if (c >= Tov && level >= 1) {
            // split: new dir one level down; the new id plus every id of the bucket are placed by their
            // next-level slot; no recursive split (a child may hold up to T+1 ids)
            std::vector<int32_t> old;
            old.swap(T.buckets[e - 1]);
            const int32_t nd = (int32_t)T.dirs.size();
            T.dirs.emplace_back(tp.W, 0);
            const int newpos = slot_of(tp, h, level - 1);
            bool shared = false;
            auto place = [&](int32_t y, int pos) {
                int32_t& ce = T.dirs[nd][pos];
                if (ce == 0) {
                    T.buckets.emplace_back();
                    ce = (int32_t)T.buckets.size();
                }
                T.buckets[ce - 1].push_back(y);
            };
            place(id, newpos);
            for (int32_t y : old) {
                const int pos = slot_of(tp, T.keys[y], level - 1);
                if (pos == newpos) shared = true;
                place(y, pos);
            }
            if (!shared) T.singleton_splits++;               // reference would leave a mis-flagged pointer (Q1)
            T.splits++;
            T.dirs[dir][slot] = -(nd + 1);
            return;
        }
        if (e == 0) {
            T.buckets.emplace_back();
            T.dirs[dir][slot] = (int32_t)T.buckets.size();
            T.buckets.back().push_back(id);
        } else {
            T.buckets[e - 1].push_back(id);                  // level-0 buckets (and c < T) just grow
        }
        return;
    }
}

// RandomDrawTreeMap.java:1817-1932 (remove / removeInternal / recursiveDirDelete), with the tree geometry of the
// configuration instead of the hard-coded 4 levels x 7 bits (quirk Q12): the id leaves its bucket (the order of the others
// is kept); a bucket that becomes empty frees its slot; a directory node that becomes empty is deleted from its parent,
// recursively — except a root, which stays (":1913 parent is segment ... just update to null").
bool table_remove(const dpfo* o, Table& T, int32_t id) {
    const TreeParams& tp = o->tp;
    if (id < 0 || (size_t)id >= T.keys.size()) return false;
    const int32_t h = T.keys[id], pid = T.pids[id];
    int32_t dir = pid * tp.SEG + seg_of(tp, h);
    int level = tp.MAXL;
    std::vector<std::pair<int32_t, int>> path;               // (parent dir, slot) of the dirs descended into
    while (true) {
        const int slot = slot_of(tp, h, level);
        const int32_t e = T.dirs[dir][slot];
        if (e < 0) { path.emplace_back(dir, slot); dir = -e - 1; level--; if (level < 0) return false; continue; }
        if (e == 0) return false;
        std::vector<int32_t>& b = T.buckets[e - 1];
        auto it = std::find(b.begin(), b.end(), id);
        if (it == b.end()) return false;
        b.erase(it);
        T.occupancy[pid]--;
        if (b.empty()) {
            T.dirs[dir][slot] = 0;
            while (!path.empty()) {                           // recursiveDirDelete
                bool empty = true;
                for (int32_t v : T.dirs[dir]) if (v != 0) { empty = false; break; }
                if (!empty) break;
                dir = path.back().first;
                T.dirs[dir][path.back().second] = 0;
                path.pop_back();
            }
        }
        return true;
    }
}

// RandomDrawTreeMap.java:940-994 searchWithSimilarity (+ getInnerWithSimilarity :1106-1121)
inline const std::vector<int32_t>* table_lookup(const dpfo* o, const Table& T, int pid, int seg, int32_t probe) {
    const TreeParams& tp = o->tp;
    int32_t dir = pid * tp.SEG + seg;
    for (int level = tp.MAXL; level >= 0; --level) {
        const int32_t e = T.dirs[dir][slot_of(tp, probe, level)];
        if (e == 0) return nullptr;
        if (e > 0) return &T.buckets[e - 1];
        dir = -e - 1;
    }
    return nullptr;
}

// RandomDrawTreeMap.java:742-797 (dense, multi-probe), :686-732 (sparse vector, no probes), :630-675 (by id)
void table_query(dpfo* o, const Table& T, int table, int32_t h, int32_t qid, int steps, int probe_mode,
                 std::vector<int32_t>& out) {
    const TreeParams& tp = o->tp;
    const int seg = seg_of(tp, h);
    const int pid = partition_id(h, &o->Ap[(size_t)table * o->cfg.pb * 32], o->cfg.pb, o->cfg.key_transform, o->part_b(table),
                                 o->part_w(table));
    const int np = 1 << o->cfg.pb;
    int32_t probes[32];
    int nprobes = 0;
    if (probe_mode == DPFO_PROBE_DENSE) {
        const int len = 32 - java_nlz(h) - 4;                // `new int[32-nlz(h)-4]`
        if (len < 0) { o->nlz_gt28++; return; }              // NegativeArraySizeException in the reference (Q4)
        for (int i = 0; i < len; ++i) probes[nprobes++] = h ^ lsl(1, i);
    } else {
        probes[nprobes++] = h;
    }
    for (int pi = 0; pi < nprobes; ++pi) {
        for (int s = 0; s < np; ++s) {                       // findStepWiseSubIndexIDs :613-621
            if (java_bitcount(s ^ pid) > steps) continue;
            if (!o->owns(table, s)) continue;                // another shard owns it
            const std::vector<int32_t>* b = table_lookup(o, T, s, seg, probes[pi]);
            if (!b) continue;
            for (int32_t y : *b) {
                // `ln.key != key` is reference inequality on boxed Integers (:982): effective only inside
                // the Integer cache -128..127 (quirk Q3)
                if (o->cfg.self_exclude_small_ids && y == qid && qid >= -128 && qid <= 127) continue;
                out.push_back(y);
            }
        }
    }
}

void sort_unique(std::vector<int32_t>& v) {
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
}

void store_candidates(dpfo* o, std::vector<std::vector<int32_t>>& per_q) {
    const int64_t nq = (int64_t)per_q.size();
    o->cand_off.assign(nq + 1, 0);
    for (int64_t i = 0; i < nq; ++i) o->cand_off[i + 1] = o->cand_off[i] + (int64_t)per_q[i].size();
    o->cand_ids.resize(o->cand_off[nq]);
    for (int64_t i = 0; i < nq; ++i)
        std::copy(per_q[i].begin(), per_q[i].end(), o->cand_ids.begin() + o->cand_off[i]);
}

// DensevectorRDFInit.scala:335-360,414-432: threads own table slices, loop all queries, union under a lock.
// Here each thread fills a private per-query list for its slice; lists are merged and de-duplicated afterwards
// (set union is order-independent).
int64_t query_with_keys(dpfo* o, const std::vector<int32_t>& qkeys /* L x nq */, int64_t nq, const int32_t* qids,
                        int steps, int probe_mode, int nthreads) {
    const int L = o->cfg.L;
    int T = std::min(resolve_threads(nthreads), L);
    std::vector<std::vector<std::vector<int32_t>>> part(T, std::vector<std::vector<int32_t>>(nq));
    std::vector<std::thread> th;
    for (int w = 0; w < T; ++w) {
        th.emplace_back([&, w]() {
            const int t0 = w * L / T, t1 = (w + 1) * L / T;
            for (int64_t i = 0; i < nq; ++i)
                for (int t = t0; t < t1; ++t)
                    table_query(o, o->tables[t], t, qkeys[(size_t)t * nq + i], qids ? qids[i] : -1000000, steps,
                                probe_mode, part[w][i]);
        });
    }
    for (auto& x : th) x.join();
    std::vector<std::vector<int32_t>> per_q(nq);
    parallel_for(nq, nthreads, [&](int64_t lo, int64_t hi, int) {
        for (int64_t i = lo; i < hi; ++i) {
            size_t tot = 0;
            for (int w = 0; w < T; ++w) tot += part[w][i].size();
            per_q[i].reserve(tot);
            for (int w = 0; w < T; ++w) per_q[i].insert(per_q[i].end(), part[w][i].begin(), part[w][i].end());
            sort_unique(per_q[i]);
        }
    });
    store_candidates(o, per_q);
    return o->cand_off[nq];
}

struct Scored { double s; int32_t id; };

}  // namespace

extern "C" {

dpfo* dpfo_create(const dpfo_cfg* cfg) {
    if (!cfg || cfg->L <= 0 || cfg->k <= 0 || cfg->k > 32 || cfg->pb < 0 || cfg->pb > 8 || cfg->P <= 0) return nullptr;
    if (cfg->dir_node_size < 2 || cfg->bucket_bits < 1 || cfg->bucket_bits > 32) return nullptr;
    dpfo* o = new dpfo();
    o->cfg = *cfg;
    o->tp = tree_params(cfg->bucket_bits, cfg->dir_node_size, cfg->k);
    if (o->tp.MAXL < 0) { delete o; return nullptr; }
    o->tables.resize(cfg->L);
    for (auto& T : o->tables) table_init(o, T);
    o->Ap.assign((size_t)cfg->L * cfg->pb * 32, 0.0);
    o->cand_off.assign(1, 0);
    return o;
}

void dpfo_destroy(dpfo* o) { delete o; }

int dpfo_set_family(dpfo* o, const double* A, const int32_t* chain_idx, const double* b, const int32_t* w) {
    const dpfo_cfg& c = o->cfg;
    o->A.assign(A, A + (size_t)c.P * c.d);
    o->chain.assign(chain_idx, chain_idx + (size_t)c.L * c.k);
    for (int32_t v : o->chain) if (v < 0 || v >= c.P) return -1;
    if (c.family_kind == 1) {
        if (!b || !w) return -2;
        o->fb.assign(b, b + c.P);
        o->fw.assign(w, w + c.P);
    }
    return 0;
}

/* explicit shard: owned[p] != 0 for the 2^pb sub-indexes this instance builds and searches (the GPU library's balanced
 * assignment, dpf_set_balanced_partition); NULL restores p % world == rank */
int dpfo_set_owned(dpfo* o, const uint8_t* owned) {
    if (!o) return 1;
    o->owned.clear();
    if (!owned) return 0;
    for (int t = 0; t < o->cfg.L; ++t) o->owned.insert(o->owned.end(), owned, owned + ((size_t)1 << o->cfg.pb));
    return 0;
}

/* the same per (table, sub-index) cell: owned[t * 2^pb + p].  Each cell is one of the reference's per-partition stores
 * (RandomDrawTreeMap.java:1430-1459); which instance holds it does not change what a search over all instances finds */
int dpfo_set_owned_cells(dpfo* o, const uint8_t* owned) {
    if (!o) return 1;
    o->owned.clear();
    if (!owned) return 0;
    o->owned.assign(owned, owned + ((size_t)o->cfg.L << o->cfg.pb));
    return 0;
}

int dpfo_set_partitioners(dpfo* o, const double* Ap) {
    o->Ap.assign(Ap, Ap + (size_t)o->cfg.L * o->cfg.pb * 32);
    o->pb_b.clear();
    o->pb_w.clear();
    return 0;
}

/* the partitioners of a pStable index (see partition_id): Ap as above, b and w of the L x pb chain functions */
int dpfo_set_partitioners_pstable(dpfo* o, const double* Ap, const double* b, const int32_t* w) {
    if (!o || !Ap || !b || !w) return 1;
    o->Ap.assign(Ap, Ap + (size_t)o->cfg.L * o->cfg.pb * 32);
    o->pb_b.assign(b, b + (size_t)o->cfg.L * o->cfg.pb);
    o->pb_w.assign(w, w + (size_t)o->cfg.L * o->cfg.pb);
    return 0;
}

int dpfo_hash_dense(dpfo* o, const double* X, int64_t n, int32_t* keys_out, int32_t* pids_out, int nthreads) {
    const int L = o->cfg.L, d = o->cfg.d, P = o->cfg.P;
    parallel_for(n, nthreads, [&](int64_t lo, int64_t hi, int) {
        std::vector<double> dots(P), scratch(32);
        std::vector<int32_t> kk(L);
        for (int64_t i = lo; i < hi; ++i) {
            project_dense(o, X + (size_t)i * d, dots.data());
            keys_from_dots(o, dots.data(), kk.data(), scratch.data());
            for (int t = 0; t < L; ++t) {
                keys_out[(size_t)t * n + i] = kk[t];
                if (pids_out)
                    pids_out[(size_t)t * n + i] =
                        partition_id(kk[t], &o->Ap[(size_t)t * o->cfg.pb * 32], o->cfg.pb, o->cfg.key_transform, o->part_b(t),
                                     o->part_w(t));
            }
        }
    });
    return 0;
}

int dpfo_hash_csr(dpfo* o, const int64_t* indptr, const int32_t* indices, const double* values, int64_t n,
                  int32_t* keys_out, int32_t* pids_out, int nthreads) {
    const int L = o->cfg.L, P = o->cfg.P;
    parallel_for(n, nthreads, [&](int64_t lo, int64_t hi, int) {
        std::vector<double> dots(P), scratch(32);
        std::vector<int32_t> kk(L);
        for (int64_t i = lo; i < hi; ++i) {
            project_csr(o, indices + indptr[i], values + indptr[i], (int)(indptr[i + 1] - indptr[i]), dots.data());
            keys_from_dots(o, dots.data(), kk.data(), scratch.data());
            for (int t = 0; t < L; ++t) {
                keys_out[(size_t)t * n + i] = kk[t];
                if (pids_out)
                    pids_out[(size_t)t * n + i] =
                        partition_id(kk[t], &o->Ap[(size_t)t * o->cfg.pb * 32], o->cfg.pb, o->cfg.key_transform, o->part_b(t),
                                     o->part_w(t));
            }
        }
    });
    return 0;
}

static int fit_common(dpfo* o, const std::vector<int32_t>& keys, const std::vector<int32_t>& pids, int64_t n, int nthreads) {
    const int L = o->cfg.L;
    const int64_t base = o->n;
    int T = std::min(resolve_threads(nthreads), L);
    std::vector<std::thread> th;
    for (int w = 0; w < T; ++w) {
        th.emplace_back([&, w]() {
            const int t0 = w * L / T, t1 = (w + 1) * L / T;
            for (int t = t0; t < t1; ++t) {
                Table& Tb = o->tables[t];
                Tb.keys.resize(base + n);
                Tb.pids.resize(base + n);
                for (int64_t i = 0; i < n; ++i) {
                    Tb.keys[base + i] = keys[(size_t)t * n + i];
                    Tb.pids[base + i] = pids[(size_t)t * n + i];
                }
                for (int64_t i = 0; i < n; ++i) {
                    if (!o->owns(t, pids[(size_t)t * n + i])) continue;
                    table_insert(o, Tb, (int32_t)(base + i), keys[(size_t)t * n + i], pids[(size_t)t * n + i]);
                }
            }
        });
    }
    for (auto& x : th) x.join();
    o->n += n;
    return 0;
}

int dpfo_fit_dense(dpfo* o, const double* X, int64_t n, int nthreads) {
    if (o->n > 0 && !o->dense) return -1;
    o->dense = true;
    const int L = o->cfg.L;
    std::vector<int32_t> keys((size_t)L * n), pids((size_t)L * n);
    dpfo_hash_dense(o, X, n, keys.data(), pids.data(), nthreads);
    o->X.insert(o->X.end(), X, X + (size_t)n * o->cfg.d);
    return fit_common(o, keys, pids, n, nthreads);
}

int dpfo_fit_csr(dpfo* o, const int64_t* indptr, const int32_t* indices, const double* values, int64_t n, int nthreads) {
    if (o->n > 0 && o->dense) return -1;
    o->dense = false;
    const int L = o->cfg.L;
    std::vector<int32_t> keys((size_t)L * n), pids((size_t)L * n);
    dpfo_hash_csr(o, indptr, indices, values, n, keys.data(), pids.data(), nthreads);
    if (o->sp_ptr.empty()) o->sp_ptr.push_back(0);
    const int64_t base = o->sp_ptr.back();
    for (int64_t i = 0; i < n; ++i) o->sp_ptr.push_back(base + indptr[i + 1] - indptr[0]);
    o->sp_idx.insert(o->sp_idx.end(), indices + indptr[0], indices + indptr[n]);
    o->sp_val.insert(o->sp_val.end(), values + indptr[0], values + indptr[n]);
    return fit_common(o, keys, pids, n, nthreads);
}

int64_t dpfo_size(dpfo* o) { return o->n; }

// removes the ids from every table (sequentially, in the order given); returns how many (table, id) entries went
int64_t dpfo_remove(dpfo* o, const int32_t* ids, int64_t m) {
    int64_t gone = 0;
    for (int t = 0; t < o->cfg.L; ++t)
        for (int64_t j = 0; j < m; ++j)
            if (o->owns(t, o->tables[t].pids.size() > (size_t)ids[j] && ids[j] >= 0 ? o->tables[t].pids[ids[j]] : 0) &&
                table_remove(o, o->tables[t], ids[j]))
                gone++;
    return gone;
}

int64_t dpfo_query_candidates_dense(dpfo* o, const double* Q, int64_t nq, const int32_t* qids, int steps,
                                    int probe_mode, int nthreads) {
    std::vector<int32_t> qkeys((size_t)o->cfg.L * nq);
    dpfo_hash_dense(o, Q, nq, qkeys.data(), nullptr, nthreads);
    return query_with_keys(o, qkeys, nq, qids, steps, probe_mode, nthreads);
}

int64_t dpfo_query_candidates_csr(dpfo* o, const int64_t* indptr, const int32_t* indices, const double* values,
                                  int64_t nq, const int32_t* qids, int steps, int nthreads) {
    std::vector<int32_t> qkeys((size_t)o->cfg.L * nq);
    dpfo_hash_csr(o, indptr, indices, values, nq, qkeys.data(), nullptr, nthreads);
    return query_with_keys(o, qkeys, nq, qids, steps, DPFO_PROBE_NONE, nthreads);   // sparse overload has no probes (Q5)
}

int64_t dpfo_query_candidates_by_id(dpfo* o, const int32_t* qids, int64_t nq, int steps, int nthreads) {
    const int L = o->cfg.L;
    std::vector<int32_t> qkeys((size_t)L * nq);
    for (int t = 0; t < L; ++t)
        for (int64_t i = 0; i < nq; ++i) {
            if (qids[i] < 0 || qids[i] >= o->n) return -1;   // reference: System.exit(1) (RandomDrawTreeMap.java:1508-1511)
            qkeys[(size_t)t * nq + i] = o->tables[t].keys[qids[i]];
        }
    return query_with_keys(o, qkeys, nq, qids, steps, DPFO_PROBE_NONE, nthreads);
}

int dpfo_get_candidates(dpfo* o, int64_t* offsets_out, int32_t* ids_out) {
    std::copy(o->cand_off.begin(), o->cand_off.end(), offsets_out);
    if (ids_out) std::copy(o->cand_ids.begin(), o->cand_ids.end(), ids_out);
    return 0;
}

// DensevectorRDFInit.scala:472-507: scores = M * q, top-k by descending dot product.  Extensions (north-star):
// angular = dot/(|q||x|) descending, l2 = |q-x|^2 ascending.  Ties: ascending id (stated GPU contract, A.10).
int dpfo_rerank_dense(dpfo* o, const double* Q, int64_t nq, const int64_t* offsets, const int32_t* cand, int topk,
                      int metric, int32_t* ids_out, double* score_out, int nthreads) {
    if (!o->dense) return -1;
    const int d = o->cfg.d;
    parallel_for(nq, nthreads, [&](int64_t lo, int64_t hi, int) {
        std::vector<Scored> sc;
        for (int64_t i = lo; i < hi; ++i) {
            const double* q = Q + (size_t)i * d;
            double qn = 0.0;
            for (int j = 0; j < d; ++j) qn = qn + q[j] * q[j];
            sc.clear();
            for (int64_t c = offsets[i]; c < offsets[i + 1]; ++c) {
                const double* x = &o->X[(size_t)cand[c] * d];
                double s;
                if (metric == DPFO_METRIC_L2) {
                    s = 0.0;
                    for (int j = 0; j < d; ++j) { const double df = q[j] - x[j]; s = s + df * df; }
                } else {
                    s = 0.0;
                    for (int j = 0; j < d; ++j) s = s + x[j] * q[j];
                    if (metric == DPFO_METRIC_ANGULAR) {
                        double xn = 0.0;
                        for (int j = 0; j < d; ++j) xn = xn + x[j] * x[j];
                        s = s / (std::sqrt(qn) * std::sqrt(xn));
                    }
                }
                sc.push_back({s, cand[c]});
            }
            const bool asc = (metric == DPFO_METRIC_L2);
            auto cmp = [asc](const Scored& a, const Scored& b) {
                if (a.s != b.s) return asc ? a.s < b.s : a.s > b.s;
                return a.id < b.id;
            };
            const size_t kk = std::min<size_t>(topk, sc.size());
            std::partial_sort(sc.begin(), sc.begin() + kk, sc.end(), cmp);
            for (int r = 0; r < topk; ++r) {
                ids_out[(size_t)i * topk + r] = r < (int)kk ? sc[r].id : -1;
                score_out[(size_t)i * topk + r] = r < (int)kk ? sc[r].s : std::numeric_limits<double>::quiet_NaN();
            }
        }
    });
    return 0;
}

int dpfo_query_topk_dense(dpfo* o, const double* Q, int64_t nq, const int32_t* qids, int steps, int probe_mode,
                          int topk, int metric, int32_t* ids_out, double* score_out, int nthreads) {
    if (dpfo_query_candidates_dense(o, Q, nq, qids, steps, probe_mode, nthreads) < 0) return -1;
    return dpfo_rerank_dense(o, Q, nq, o->cand_off.data(), o->cand_ids.data(), topk, metric, ids_out, score_out,
                             nthreads);
}

int64_t dpfo_num_dir_nodes(dpfo* o, int table) { return (int64_t)o->tables[table].dirs.size(); }

int64_t dpfo_dump_buckets(dpfo* o, int table, int32_t* desc_out, int64_t* off_out, int32_t* ids_out) {
    const Table& T = o->tables[table];
    const TreeParams& tp = o->tp;
    const int roots = (1 << o->cfg.pb) * tp.SEG;
    int64_t nbk = 0, nid = 0;
    struct Frame { int32_t dir; int level; int64_t path; };
    for (int r = 0; r < roots; ++r) {
        // DFS in ascending slot order => (root, path) lexicographic order
            std::vector<std::pair<Frame, int>> st;   // frame + next slot
        st.push_back({{r, tp.MAXL, 0}, 0});
        while (!st.empty()) {
            auto& top = st.back();
            if (top.second >= tp.W) { st.pop_back(); continue; }
            const int slot = top.second++;
            const Frame f = top.first;
            const int32_t e = T.dirs[f.dir][slot];
            if (e == 0) continue;
            const int64_t path = (f.path << tp.nb) | slot;
            if (e < 0) { st.push_back({{-e - 1, f.level - 1, path}, 0}); continue; }
            const std::vector<int32_t>& b = T.buckets[e - 1];
            if (b.empty()) continue;
            if (desc_out) {
                desc_out[nbk * 3 + 0] = r;
                desc_out[nbk * 3 + 1] = f.level;
                desc_out[nbk * 3 + 2] = (int32_t)path;
            }
            if (off_out) off_out[nbk] = nid;
            if (ids_out) {
                std::vector<int32_t> s(b);
                std::sort(s.begin(), s.end());
                std::copy(s.begin(), s.end(), ids_out + nid);
            }
            nid += (int64_t)b.size();
            nbk++;
        }
    }
    if (off_out) off_out[nbk] = nid;
    return nbk;
}

int dpfo_stats(dpfo* o, int64_t* s, double* occ_out) {
    for (int i = 0; i < 8; ++i) s[i] = 0;
    for (const Table& T : o->tables) { s[0] += T.singleton_splits; s[2] += T.splits; }
    s[1] = o->nlz_gt28.load();
    s[3] = o->tp.MAXL; s[4] = o->tp.nb; s[5] = o->tp.SEG; s[6] = o->n;
    if (occ_out) {
        const int np = 1 << o->cfg.pb;
        for (int p = 0; p < np; ++p) {
            double a = 0;
            for (const Table& T : o->tables) a += (double)T.occupancy[p];
            occ_out[p] = a / o->cfg.L;                       // DensevectorRDFInit.scala:515-530
        }
    }
    return 0;
}

double dpfo_dot_dense(const double* a, const double* x, int d) { return dot_dense(a, x, d); }
double dpfo_dot_sparse(const int32_t* ia, const double* va, int na, const int32_t* ib, const double* vb, int nb) {
    return dot_sparse(ia, va, na, ib, vb, nb);
}
int32_t dpfo_angle_key_from_dots(const double* dots, int k) { return angle_key_from_dots(dots, k); }
int32_t dpfo_pstable_key_from_dots(const double* dots, const double* b, const int32_t* w, int k) {
    return pstable_key_from_dots(dots, b, w, k);
}
int32_t dpfo_sampling_key(int32_t key) { return sampling_one_key(key); }
void dpfo_sampling_index(int32_t* s) { std::memcpy(s, sampling_index().sigma, sizeof(int32_t) * 32); }
int32_t dpfo_continue_bits_count(int32_t key) { return continue_bits_count(key); }
int32_t dpfo_angle_new_method(int32_t key) { return angle_new_method(key); }
int32_t dpfo_partition_id(int32_t h, const double* Ap_t, int pb, int key_transform) {
    return partition_id(h, Ap_t, pb, key_transform);
}
int32_t dpfo_partition_id_pstable(int32_t h, const double* Ap_t, int pb, int key_transform, const double* b, const int32_t* w) {
    return partition_id(h, Ap_t, pb, key_transform, b, w);
}
// Hasher.scala:18-37 (DefaultHasher on Int keys; `>>` is arithmetic in Scala)
int32_t dpfo_default_hasher(int32_t key) {
    int32_t h = (int32_t)((uint32_t)((key >> 16) ^ key) * 0x45d9f3bu);
    h = (int32_t)((uint32_t)((h >> 16) ^ h) * 0x45d9f3bu);
    h = (h >> 16) ^ h;
    return h;
}

// RandomDrawTreeMap.java:1226-1267
int32_t dpfo_dir_offset_from_slot(const int32_t* dir, int bitmap_words, int slot) {
    const int nb = ilog2_like_java(bitmap_words * 32);
    const int range_bits = nb - ilog2_like_java(bitmap_words);
    int range = 0;
    if (bitmap_words > 1) range = lsr(slot, range_bits);
    const int within = slot & ((1 << range_bits) - 1);
    int is_set = (lsr(dir[range], within) & 1) << 1;
    int offset = 0;
    for (int i = 0; i < range; ++i) offset += java_bitcount(dir[i]);
    const int32_t mask = lsl(1, within) - 1;
    offset += bitmap_words + java_bitcount(dir[range] & mask);
    return -offset + is_set * offset;
}

void dpfo_tree_params(int bucket_bits, int dir_node_size, int chain_length, int32_t* out) {
    TreeParams tp = tree_params(bucket_bits, dir_node_size, chain_length);
    out[0] = tp.SEG; out[1] = tp.nb; out[2] = tp.mask; out[3] = tp.MAXL;
}

}  // extern "C"
