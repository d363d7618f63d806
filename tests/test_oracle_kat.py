"""Pins the CPU oracle on every known-answer test the reference's own test tree holds for the hot path
(SURVEY.md §8c).  CPU only."""
import os

import numpy as np
import pytest

from oracle import oracle_py as O


# ---- SimilarityCalculatorSuite.scala:6-34 -----------------------------------------------------------------
def test_dot_sparse_sparse_kats():
    # :7-11 — the reference asserts 1.8, but 0.5*0.2 + 1.2*1.5 = 1.9 under any IEEE evaluation; the reference
    # test tree does not compile as shipped so this assertion was never executed.  We pin the
    # arithmetic (1.9) and record the discrepancy.
    s = O.dot_sparse([0, 2, 3, 9], [1.0, 0.5, 1.2, 2.5], [2, 3], [0.2, 1.5])
    assert s == 0.5 * 0.2 + 1.2 * 1.5 == 1.9
    # :12-16
    assert O.dot_sparse([0, 2, 3, 9], [1.0] * 4, [5, 6], [1.0, 1.0]) == 0.0
    # :19-25 mismatched bit in the middle
    assert O.dot_sparse([0, 2, 3, 9], [1.0, 1.0, 2.0, 1.0], [3, 6, 9], [1.3, 1.0, 1.0]) == 3.6


def test_dot_dense_kats():
    # :26-33
    assert O.dot_dense([0.1, 0.2, 0.3], [0.2, 0.3, 0.4]) == 0.2
    assert O.dot_dense([1.1, 1.2, 1.3], [0.1, 0.2, 0.3]) == 0.74


def test_dot_dense_is_sequential_unfused():
    rng = np.random.default_rng(7)
    for d in (1, 3, 100, 961):
        a, x = rng.standard_normal(d), rng.standard_normal(d)
        s = 0.0
        for j in range(d):
            s = s + a[j] * x[j]          # python floats: separate rounding of product and sum
        assert O.dot_dense(a, x) == s


# ---- AngleHashSuite.scala:10-33 ------------------------------------------------------------------------------
def test_angle_key_kats():
    f1 = np.array([1.0, 1.0, 1.0])
    f2 = np.array([1.5, -1.0, 0.0])
    v = np.array([1.0, 1.0, 1.0])
    assert O.angle_key_from_dots([O.dot_dense(f1, v)]) == -2147483648
    assert O.angle_key_from_dots([O.dot_dense(f1, v), O.dot_dense(f2, v)]) == -1073741824


def test_angle_sign_zero_is_zero_bit():
    assert O.angle_key_from_dots([0.0]) == 0
    assert O.angle_key_from_dots([-0.0]) == 0
    assert O.angle_key_from_dots([-1e-300, 1e-300]) == 1 << 30
    # full 32-bit chain: first function -> bit 31
    dots = [-1.0] * 32
    dots[0] = 1.0
    assert O.angle_key_from_dots(dots) == -2147483648
    dots = [-1.0] * 32
    dots[31] = 1.0
    assert O.angle_key_from_dots(dots) == 1


# ---- PStableHashSuite.scala:14-41 ----------------------------------------------------------------------------
def test_pstable_key_kats():
    v = np.ones(3)
    assert O.pstable_key_from_dots([O.dot_dense(np.ones(3), v)], [3.0], [10]) == 923521
    dots = [O.dot_dense(np.ones(3), v), O.dot_dense(2 * np.ones(3), v)]
    assert O.pstable_key_from_dots(dots, [3.0, 3.0], [4, 4]) == -1806530940


def test_pstable_truncates_toward_zero():
    # quirk Q10: `.toInt` truncates; floor would give -1
    def java_bytes_hash(ints):
        h = 1
        for q in ints:
            for b in int(q).to_bytes(4, "big", signed=True):
                sb = b - 256 if b > 127 else b
                h = (31 * h + sb) & 0xFFFFFFFF
        return h - (1 << 32) if h & 0x80000000 else h

    assert O.pstable_key_from_dots([-1.5], [0.0], [4]) == java_bytes_hash([0])
    assert O.pstable_key_from_dots([-9.0], [0.0], [4]) == java_bytes_hash([-2])
    assert O.pstable_key_from_dots([1e300], [0.0], [1]) == java_bytes_hash([2**31 - 1])
    assert O.pstable_key_from_dots([float("nan")], [0.0], [1]) == java_bytes_hash([0])


# ---- Sampling.scala:6-39 -------------------------------------------------------------------------------------
def _java_random_shuffle_32(seed):
    mask = (1 << 48) - 1
    s = (seed ^ 0x5DEECE66D) & mask

    def nxt(bits):
        nonlocal s
        s = (s * 0x5DEECE66D + 0xB) & mask
        return s >> (48 - bits)          # bits = 31 here: always a non-negative int

    def next_int(n):
        if n & -n == n:
            return (n * nxt(31)) >> 31
        while True:
            bits = nxt(31)
            val = bits % n
            if bits - val + (n - 1) < (1 << 31):
                return val

    buf = list(range(32))
    for n in range(32, 1, -1):
        k = next_int(n)
        buf[n - 1], buf[k] = buf[k], buf[n - 1]
    return buf


def test_sampling_index_matches_java_lcg():
    sg = O.sampling_index()
    assert sorted(sg.tolist()) == list(range(32))
    assert sg.tolist() == _java_random_shuffle_32(88387)


def test_sampling_key_is_that_bit_permutation():
    sg = O.sampling_index()
    rng = np.random.default_rng(3)
    for key in rng.integers(-2**31, 2**31, 50):
        exp = 0
        for j in range(32):
            exp |= ((int(key) >> int(sg[j])) & 1) << (31 - j)
        exp = exp - (1 << 32) if exp & 0x80000000 else exp
        assert O.sampling_key(key) == exp


# ---- significantBits.scala ----------------------------------------------------------------------------------
def test_continue_bits_count_hand_cases():
    # key = 0: no runs -> only the (zero) first four bits
    assert O.continue_bits_count(0) == 0
    # low 28 bits all ones: one run of 28 >= 6 -> all four counters 1; top nibble 0
    k = (1 << 28) - 1
    assert O.continue_bits_count(k) == (1 << 21) | (1 << 14) | (1 << 7) | 1
    # single isolated one: counter[3] only; `reverse(i) << ((3-i)*7)` puts counter[3] at bit 21
    assert O.continue_bits_count(1) == 1 << 21
    # a run of exactly 4 (>= numOfBits(1)): counters 1,2,3 -> bits 7,14,21
    assert O.continue_bits_count(0b1111) == (1 << 21) | (1 << 14) | (1 << 7)
    # top nibble is copied through
    assert O.continue_bits_count(np.int32(-2**31)) == np.int32(-2**31)


def test_angle_new_method_fields():
    key = 0x7ABCDEF
    out = O.angle_new_method(key)
    assert out & 0x7F == key & 0x7F
    assert (out >> 7) & 0x7F == (key >> 7) & 0x7F
    assert (out >> 21) & 0x7F == (key >> 21) & 0x7F
    assert 0 <= (out >> 14) & 0x7F <= 9


# ---- Hasher.scala:18-37 --------------------------------------------------------------------------------------
def test_default_hasher_matches_int_mixer():
    def mix(k):
        def i32(v):
            v &= 0xFFFFFFFF
            return v - (1 << 32) if v & 0x80000000 else v
        h = i32(((k >> 16) ^ k) * 0x45d9f3b)
        h = i32(((h >> 16) ^ h) * 0x45d9f3b)
        return i32((h >> 16) ^ h)
    for k in (0, 1, 12345, -7, 2**31 - 1, -2**31):
        assert O.default_hasher(k) == mix(k)


# ---- Partitioner.scala:40-64 --------------------------------------------------------------------------------
def test_partition_id_matches_definition(golden_dir):
    g = np.load(os.path.join(golden_dir, "partition_family_angle.npz"))
    Ap = g["rows32"]                     # the reference's pinned 2 x 32-d partitioner functions
    assert Ap.shape == (2, 32)
    rng = np.random.default_rng(11)
    for h in list(rng.integers(-2**31, 2**31, 200)) + [0, -1, 1, -2**31]:
        bits = [(int(h) >> i) & 1 for i in range(32)]
        pid = 0
        for j in range(2):
            s = 0.0
            for i in range(32):
                if bits[i]:
                    s += Ap[j][i] * 1.0
            pid = (pid << 1) | (0 if s <= 0 else 1)
        assert O.partition_id(h, Ap) == pid
    assert O.partition_id(0, Ap) == 0    # empty sum


def test_partition_id_pstable_chain_matches_definition():
    """With mclab.lsh.name = pStable the partitioner's LSH is a pStable one (confForPartitioner falls back to the main
    conf, DensevectorRDFInit.scala:63-70): sub-index = PStableHashChain.compute(bits of the key).hashCode >>> (32 - pb)
    (Partitioner.scala:58, PStableHashFamily.scala:155-177), restated here in plain Python."""
    import math
    rng = np.random.default_rng(12)
    pb = 3
    Ap = rng.standard_normal((pb, 32))
    b = rng.random(pb) * 4
    w = np.array([4, 3, 7], np.int32)

    def to_int(x):                       # Scala Double.toInt
        if math.isnan(x):
            return 0
        return max(-2**31, min(2**31 - 1, int(x)))

    seen = set()
    for h in list(rng.integers(-2**31, 2**31, 300)) + [0, -1, 1, -2**31]:
        qs = []
        for j in range(pb):
            s = 0.0
            for i in range(32):
                if (int(h) >> i) & 1:
                    s += Ap[j][i] * 1.0
            qs.append(to_int((s + b[j]) / int(w[j])))
        code = 1
        for q in qs:
            for byte in int(q).to_bytes(4, "big", signed=True):
                code = (31 * code + (byte - 256 if byte > 127 else byte)) & 0xFFFFFFFF
        want = code >> (32 - pb)
        assert O.partition_id(h, Ap, 0, b, w) == want
        seen.add(want)
    assert len(seen) > 2                 # (the top bits of Arrays.hashCode over a dozen small bytes do not reach every value)
    # with the "sampling" key transform the hash code is permuted before the shift (LSH.scala:113-126)
    h = 123456789
    code = O.pstable_key_from_dots([sum(Ap[j][i] for i in range(32) if (h >> i) & 1) for j in range(pb)], b, w)
    assert O.partition_id(h, Ap, 1, b, w) == (O.sampling_key(code) & 0xFFFFFFFF) >> (32 - pb)


# ---- RandomDrawTreeMap.java:435-465 -------------------------------------------------------------------------
def test_tree_params():
    assert O.tree_params(28, 32, 32) == dict(SEG=16, nb=5, mask=31, MAXL=4)       # TestSettings defaults
    assert O.tree_params(28, 128, 32) == dict(SEG=16, nb=7, mask=127, MAXL=3)     # RandomDrawTreeMapTest
    assert O.tree_params(28, 64, 32) == dict(SEG=16, nb=6, mask=63, MAXL=3)       # DirectoryNodeSuite


# ---- RandomDrawTreeMapTest.java:596-685 (dirOffsetFromSlot vs brute force over random bitmaps) --------------
@pytest.mark.parametrize("words", [1, 2, 4])
def test_dir_offset_from_slot_bruteforce(words):
    rng = np.random.default_rng(words)
    for _ in range(200):
        bm = rng.integers(-2**31, 2**31, words).astype(np.int32)
        for slot in range(words * 32):
            w, b = divmod(slot, 32)
            isset = (int(bm[w]) >> b) & 1
            before = sum(bin(int(bm[i]) & 0xFFFFFFFF).count("1") for i in range(w)) + \
                bin(int(bm[w]) & ((1 << b) - 1)).count("1")
            exp = words + before
            assert O.dir_offset_from_slot(bm, slot) == (exp if isset else -exp)


# ---- RandomDrawTreeMapTest.java:185-285 (bucket split structure; constant hash, dirNodeSize=128) ------------
def _const_key_oracle(T, n, dir_node_size=128):
    # one table, one function whose projection is always <= 0  => key 0 for every vector, pid 0
    o = O.Oracle(d=2, L=1, k=32, P=1, pb=0, bucket_bits=28, dir_node_size=dir_node_size, bucket_overflow=T)
    o.set_family(np.array([[-1.0, -1.0]]), np.zeros((1, 32), np.int32))
    o.set_partitioners(np.zeros((1, 0, 32)))
    o.fit_dense(np.ones((n, 2)), nthreads=1)
    return o


def test_hash_dir_expand_like_reference():
    T = 4                                  # RandomDrawTreeMap.BUCKET_OVERFLOW default (RandomDrawTreeMap.java:36)
    o = _const_key_oracle(T, T)
    desc, off, ids = o.dump_buckets(0)
    # "segment should not be expanded": one bucket in root (level MAXL=3) slot 0 holding ids 0..T-1
    assert desc.tolist() == [[0, 3, 0]] and ids.tolist() == list(range(T))
    assert o.num_dir_nodes(0) == 16        # only the 16 segment roots
    o2 = _const_key_oracle(T, T + 1)
    desc, off, ids = o2.dump_buckets(0)
    # "adding one more item should trigger dir expansion to next level": root slot 0 -> dir -> slot 0 bucket, T+1 ids
    assert desc.tolist() == [[0, 2, 0]] and ids.tolist() == list(range(T + 1))
    assert o2.num_dir_nodes(0) == 17
    assert o2.stats()["singleton_splits"] == 0 and o2.stats()["splits"] == 1


def test_split_is_not_recursive_and_level0_unbounded():
    T = 4
    # constant key: every split moves the whole bucket one level down; level-0 bucket then grows unbounded
    o = _const_key_oracle(T, 40, dir_node_size=32)     # MAXL = 4
    desc, off, ids = o.dump_buckets(0)
    assert desc.tolist() == [[0, 0, 0]] and ids.tolist() == list(range(40))
    assert o.stats()["splits"] == 4
