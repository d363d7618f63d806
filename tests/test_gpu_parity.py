"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.
Bit-exact for keys, partition ids, bucket membership and candidate sets; ids exact / scores within 1e-12 relative
for the re-rank (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

from oracle import oracle_py as O
from similaritysearchbyrdf_b200 import _lib as B
from similaritysearchbyrdf_b200 import synth
from tests import util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg1():
    X, Q = synth.config1()
    A, chain, Ap = U.make_functions(100)
    return X, Q, A, chain, Ap


# ---- K1: keys and partition ids -------------------------------------------------------------------------------
def test_hash_dense_config1_bit_exact(cfg1):
    X, Q, A, chain, Ap = cfg1
    o, ix = U.make_oracle(100, A, chain, Ap), U.make_index(100, A, chain, Ap)
    ko, po = o.hash_dense(X)
    kg, pg = ix.hash_dense(X)
    assert np.array_equal(ko, kg)
    assert np.array_equal(po, pg)
    print("near-zero fix-ups:", ix.stats()["near_zero_fixups"])


def test_hash_dense_pinned_reference_family(golden_dir):
    # the reference's own pinned functions: 10 tables x 32 functions, 100-d, 98 distinct ids
    g = np.load(os.path.join(golden_dir, "angle_family_tablenum10.npz"))
    ids, rows = g["ids"], g["rows"]
    uniq, first, inv = np.unique(ids, return_index=True, return_inverse=True)
    A = rows[first]
    chain = inv.reshape(10, 32).astype(np.int32)
    Ap = synth.partitioner_family(10, 3, 5)
    X, _ = synth.config1(n=5000)
    o, ix = U.make_oracle(100, A, chain, Ap), U.make_index(100, A, chain, Ap)
    ko, po = o.hash_dense(X)
    kg, pg = ix.hash_dense(X)
    assert np.array_equal(ko, kg) and np.array_equal(po, pg)


@pytest.mark.parametrize("d,n", [(1, 7), (3, 65), (33, 130), (96, 1000), (128, 2049), (960, 300)])
def test_hash_dense_shapes(d, n):
    A, chain, Ap = U.make_functions(d, family_size=max(40, d), table_num=4, permutation_num=2, seed=d)
    X = np.random.default_rng(d).standard_normal((n, d))
    o, ix = U.make_oracle(d, A, chain, Ap), U.make_index(d, A, chain, Ap)
    ko, po = o.hash_dense(X)
    kg, pg = ix.hash_dense(X)
    assert np.array_equal(ko, kg) and np.array_equal(po, pg)


def test_hash_dense_near_zero_projections_are_fixed_up():
    # vectors built to be (numerically) orthogonal to some functions: the DMMA sum and the reference's sequential
    # unfused sum can land on different sides of zero; the fix-up must make the signs exact
    d = 64
    A, chain, Ap = U.make_functions(d, family_size=64, table_num=3, permutation_num=1, seed=3)
    rng = np.random.default_rng(4)
    X = rng.standard_normal((4000, d))
    for i in range(0, 4000, 2):
        a = A[rng.integers(0, A.shape[0])]
        X[i] -= (X[i] @ a) / (a @ a) * a          # projection ~ 1e-17 .. 1e-16
    X[1] = 0.0                                      # exact zeros -> bit 0 everywhere
    o, ix = U.make_oracle(d, A, chain, Ap), U.make_index(d, A, chain, Ap)
    ko, _ = o.hash_dense(X)
    kg, _ = ix.hash_dense(X)
    assert np.array_equal(ko, kg)
    assert ix.stats()["near_zero_fixups"] > 0


@pytest.mark.parametrize("transform", ["sampling", "continueBitsCount", "angleNewMethod"])
def test_hash_key_transforms(cfg1, transform):
    X, Q, A, chain, Ap = cfg1
    kt = O.TRANSFORMS[transform]
    o = U.make_oracle(100, A, chain, Ap, key_transform=kt)
    ix = U.make_index(100, A, chain, Ap, key_transform=kt)
    ko, po = o.hash_dense(X[:3000])
    kg, pg = ix.hash_dense(X[:3000])
    assert np.array_equal(ko, kg) and np.array_equal(po, pg)


@pytest.mark.parametrize("family_size,k,pb", [(20, 32, 1), (70, 17, 3), (128, 32, 5), (100, 32, 8)])
def test_hash_table_driven_packing_matches_bit_by_bit(family_size, k, pb):
    """k_pack_keys_tab (byte-indexed lookup tables, FP32 partial sums with an error bound, exact redo below it) against
    the bit-by-bit kernel and the oracle, with partitioner rows built so that many sums land inside the bound."""
    d, L = 48, 6
    A, chain, Ap = U.make_functions(d, family_size=family_size, table_num=3, permutation_num=2, k=k, pb=pb, seed=family_size)
    L = chain.shape[0]
    Ap = Ap.copy()
    Ap[0, 0, :] = 0.0                       # all-zero row: every sum is exactly 0 -> bit 0
    if pb > 1:
        Ap[1, pb - 1, 1::2] = -Ap[1, pb - 1, 0::2]   # pairs that cancel: sums are 0 or tiny whenever both bits agree
    Ap[2, 0, :] *= 1e-30                    # tiny coefficients: the bound scales with them
    Ap[3, 0, :16] = 1e20                    # one-sided huge partials in the low half, ordinary values above
    X = np.random.default_rng(1).standard_normal((70001, d))    # (the table-driven kernel serves n >= 65536)
    o = U.make_oracle(d, A, chain, Ap)
    ko, po = o.hash_dense(X)
    ix = U.make_index(d, A, chain, Ap)
    kg, pg = ix.hash_dense(X)
    ix.set_debug_option(B.DBG_HASH_EXACT, 2)
    kb, pb_ = ix.hash_dense(X)
    assert np.array_equal(kg, kb) and np.array_equal(pg, pb_)
    assert np.array_equal(ko, kg) and np.array_equal(po, pg)


def test_hash_pstable_family():
    d, L, k = 24, 3, 32
    rng = np.random.default_rng(9)
    A = rng.standard_normal((L * k, d))
    chain = np.arange(L * k, dtype=np.int32).reshape(L, k)
    b = rng.random(L * k) * 4
    w = np.full(L * k, 4, np.int32)
    # the partitioner chains of a pStable index are pStable functions too (DensevectorRDFInit.scala:63-70)
    Ap, pb_b, pb_w = synth.pstable_partitioner_family(L, 3, 0.0, 1.0, 4, 2)
    X = rng.standard_normal((3000, d)) * 3
    o = U.make_oracle(d, A, chain, Ap, family_kind=1, b=b, w=w, pb_b=pb_b, pb_w=pb_w)
    ix = U.make_index(d, A, chain, Ap, family_kind=1, b=b, w=w, pb_b=pb_b, pb_w=pb_w)
    ko, po = o.hash_dense(X)
    kg, pg = ix.hash_dense(X)
    assert np.array_equal(ko, kg) and np.array_equal(po, pg)
    assert len(np.unique(po)) > 2                                    # (Arrays.hashCode's top bits do not reach all 8)
    # ... and the whole path on top of them: buckets, multi-step candidates, top-k
    o.fit_dense(X)
    ix.fit_dense(X)
    U.assert_buckets_equal(o, ix, L)
    Q = X[:200] + 0.05 * rng.standard_normal((200, d))
    U.assert_csr_equal(o.query_candidates_dense(Q, None, 1), ix.query_candidates_dense(Q, None, 1))
    # the angle-chain entry point is refused on a pStable index
    from similaritysearchbyrdf_b200 import DPFIndex
    bad = DPFIndex(d=d, L=L, k=k, pb=3, family_kind=1)
    bad.set_family(A, chain, b, w)
    with pytest.raises(B.DpfError):
        bad.set_partitioners(Ap)
    bad.close()


def _small_csr(n, D, seed, mean=12):
    rng = np.random.default_rng(seed)
    nnz = np.maximum(rng.poisson(mean, n), 1)
    nnz[0] = 1
    indptr = np.concatenate([[0], np.cumsum(nnz)]).astype(np.int64)
    idx = np.concatenate([np.sort(rng.choice(D, m, replace=False)) for m in nnz]).astype(np.int32)
    val = rng.random(indptr[-1]) + 0.01
    return indptr, idx, val


def test_hash_csr_bit_exact():
    D = 500
    A, chain, Ap = U.make_functions(D, family_size=100, table_num=10, permutation_num=3, seed=21)
    indptr, idx, val = _small_csr(4000, D, 22)
    o, ix = U.make_oracle(D, A, chain, Ap), U.make_index(D, A, chain, Ap)
    ko, po = o.hash_csr(indptr, idx, val)
    kg, pg = ix.hash_csr(indptr, idx, val)
    assert np.array_equal(ko, kg) and np.array_equal(po, pg)


# ---- K3: bucket membership -----------------------------------------------------------------------------------
@pytest.mark.parametrize("T", [500, 40, 4])
def test_build_bucket_membership_config1(cfg1, T):
    X, Q, A, chain, Ap = cfg1
    o = U.make_oracle(100, A, chain, Ap, bucket_overflow=T)
    ix = U.make_index(100, A, chain, Ap, bucket_overflow=T)
    o.fit_dense(X)
    ix.fit_dense(X)
    U.assert_buckets_equal(o, ix, chain.shape[0])
    so, sg = o.stats(), ix.stats()
    assert so["splits"] == sg["splits"] and so["singleton_splits"] == sg["singleton_splits"]
    assert np.allclose(so["occupancy"], sg["occupancy"])


@pytest.mark.parametrize("dir_node_size,T", [(128, 4), (64, 3), (4, 2), (2, 1)])
def test_build_skewed_keys_order_dependent_splits(dir_node_size, T):
    # few distinct directions => heavy key collisions => the c0 bookkeeping of the split rule is exercised
    d = 8
    A, chain, Ap = U.make_functions(d, family_size=16, table_num=3, permutation_num=2, seed=5)
    rng = np.random.default_rng(6)
    C = rng.standard_normal((12, d))
    X = C[rng.integers(0, 12, 6000)] + 0.02 * rng.standard_normal((6000, d))
    o = U.make_oracle(d, A, chain, Ap, bucket_overflow=T, dir_node_size=dir_node_size)
    ix = U.make_index(d, A, chain, Ap, bucket_overflow=T, dir_node_size=dir_node_size)
    o.fit_dense(X)
    ix.fit_dense(X)
    U.assert_buckets_equal(o, ix, chain.shape[0])
    assert o.stats()["splits"] == ix.stats()["splits"] > 0
    assert o.stats()["singleton_splits"] == ix.stats()["singleton_splits"]


def test_build_append_equals_sequential_insertion(cfg1):
    X, Q, A, chain, Ap = cfg1
    o = U.make_oracle(100, A, chain, Ap, bucket_overflow=6)
    ix = U.make_index(100, A, chain, Ap, bucket_overflow=6)
    o.fit_dense(X[:3000]); o.fit_dense(X[3000:5000])
    ix.fit_dense(X[:3000]); ix.fit_dense(X[3000:5000])
    assert len(ix) == 5000
    U.assert_buckets_equal(o, ix, chain.shape[0])


# ---- K4: candidate sets --------------------------------------------------------------------------------------
@pytest.mark.parametrize("steps", [0, 1, 2, 3])
def test_candidates_dense_multiprobe(cfg1, steps):
    X, Q, A, chain, Ap = cfg1
    o = U.make_oracle(100, A, chain, Ap, bucket_overflow=40)
    ix = U.make_index(100, A, chain, Ap, bucket_overflow=40)
    o.fit_dense(X); ix.fit_dense(X)
    Qs = np.concatenate([Q, X[:200]])
    qids = np.concatenate([[0, 1], np.arange(200)]).astype(np.int32)     # README: query ids 0,1
    U.assert_csr_equal(o.query_candidates_dense(Qs, qids, steps), ix.query_candidates_dense(Qs, qids, steps))
    assert o.stats()["nlz_gt28"] == ix.stats()["nlz_gt28"]


def test_candidates_self_exclusion_quirk(cfg1):
    X, Q, A, chain, Ap = cfg1
    o = U.make_oracle(100, A, chain, Ap); ix = U.make_index(100, A, chain, Ap)
    o.fit_dense(X[:5000]); ix.fit_dense(X[:5000])
    qids = np.array([5, 127, 128, 4000], np.int32)
    co = o.query_candidates_dense(X[qids], qids, 0)
    cg = ix.query_candidates_dense(X[qids], qids, 0)
    U.assert_csr_equal(co, cg)
    sets = U.csr_sets(*cg)
    assert 5 not in sets[0] and 127 not in sets[1]          # inside the Integer cache: excluded
    assert 128 in sets[2] and 4000 in sets[3]               # outside: the query finds itself


def test_candidates_by_id_and_no_probe(cfg1):
    X, Q, A, chain, Ap = cfg1
    o = U.make_oracle(100, A, chain, Ap, bucket_overflow=40); ix = U.make_index(100, A, chain, Ap, bucket_overflow=40)
    o.fit_dense(X); ix.fit_dense(X)
    qids = np.arange(0, 20000, 97, dtype=np.int32)
    for steps in (0, 2):
        U.assert_csr_equal(o.query_candidates_by_id(qids, steps), ix.query_candidates_by_id(qids, steps))
        U.assert_csr_equal(o.query_candidates_dense(X[qids], qids, steps, O.PROBE_NONE),
                           ix.query_candidates_dense(X[qids], qids, steps, B.PROBE_NONE))


def test_candidates_csr_path():
    D = 500
    A, chain, Ap = U.make_functions(D, family_size=100, table_num=10, permutation_num=3, seed=21)
    indptr, idx, val = _small_csr(6000, D, 23)
    o = U.make_oracle(D, A, chain, Ap, bucket_overflow=20); ix = U.make_index(D, A, chain, Ap, bucket_overflow=20)
    o.fit_csr(indptr, idx, val); ix.fit_csr(indptr, idx, val)
    U.assert_buckets_equal(o, ix, chain.shape[0])
    qp, qi, qv = _small_csr(64, D, 24)
    for steps in (0, 1):
        U.assert_csr_equal(o.query_candidates_csr(qp, qi, qv, None, steps), ix.query_candidates_csr(qp, qi, qv, None, steps))
    qids = np.arange(0, 6000, 61, dtype=np.int32)
    U.assert_csr_equal(o.query_candidates_by_id(qids, 1), ix.query_candidates_by_id(qids, 1))


# ---- K5: re-rank / top-k -------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric", [B.METRIC_DOT, B.METRIC_ANGULAR, B.METRIC_L2])
@pytest.mark.parametrize("topk", [10, 100])
def test_topk_parity(cfg1, metric, topk):
    X, Q, A, chain, Ap = cfg1
    o = U.make_oracle(100, A, chain, Ap); ix = U.make_index(100, A, chain, Ap)
    o.fit_dense(X); ix.fit_dense(X)
    Qs = np.concatenate([Q, X[1000:1100] + 0.01])
    io, so = o.query_topk_dense(Qs, None, 1, topk, metric)
    ig, sg = ix.query_topk_dense(Qs, None, 1, topk, metric)
    U.assert_topk_close(io, so, ig, sg)


def test_rerank_given_same_candidates_odd_dim():
    d = 33
    A, chain, Ap = U.make_functions(d, family_size=40, table_num=4, permutation_num=1, seed=8)
    rng = np.random.default_rng(8)
    X = rng.standard_normal((3000, d))
    o = U.make_oracle(d, A, chain, Ap); ix = U.make_index(d, A, chain, Ap)
    o.fit_dense(X); ix.fit_dense(X)
    Q = rng.standard_normal((16, d))
    off = np.arange(0, 17 * 150, 150, dtype=np.int64)
    cand = np.concatenate([np.sort(rng.choice(3000, 150, replace=False)) for _ in range(16)]).astype(np.int32)
    off[1] = off[0]                                           # an empty candidate set -> padded row
    for metric in (0, 1, 2):
        io, so = o.rerank_dense(Q, off, cand, 10, metric)
        ig, sg = ix.rerank_dense(Q, off, cand, 10, metric)
        U.assert_topk_close(io, so, ig, sg)
        assert (ig[0] == -1).all() and np.isnan(sg[0]).all()


# ---- K5b: bucket-major re-rank (rerank_bm.cu) ----------------------------------------------------------------
@pytest.mark.parametrize("d", [8, 36, 100, 128])
@pytest.mark.parametrize("metric", [B.METRIC_DOT, B.METRIC_ANGULAR])
def test_topk_bucket_major_many_queries_share_buckets(d, metric):
    """Hundreds of queries probe the same leaf buckets: runs of pairs longer than one unit (16 queries), both DMMA
    n-blocks in use, partial last 8-column window (d = 36, 100), the smallest row a bulk copy can move (d = 8)."""
    rng = np.random.default_rng(100 + d)
    centres = rng.standard_normal((12, d))
    X = np.repeat(centres, 250, axis=0) + 0.05 * rng.standard_normal((3000, d))
    A, chain, Ap = U.make_functions(d, family_size=max(40, d), table_num=3, permutation_num=2, seed=31 + d)
    o = U.make_oracle(d, A, chain, Ap, bucket_overflow=30); ix = U.make_index(d, A, chain, Ap, bucket_overflow=30)
    o.fit_dense(X); ix.fit_dense(X)
    Qs = X[::5] + 0.01 * rng.standard_normal((600, d))
    io, so = o.query_topk_dense(Qs, None, 1, 10, metric)
    ig, sg = ix.query_topk_dense(Qs, None, 1, 10, metric)
    U.assert_topk_close(io, so, ig, sg)
    st = ix.stats()
    assert st["bm_pairs"] > 0 and st["bm_runs"] > 0, "the bucket-major path did not run"
    assert st["bm_pairs"] > 16 * 3 * 2, "expected long runs of pairs"


def test_topk_bucket_major_equals_row_major_bitwise_on_integer_data(monkeypatch):
    """Integer-valued data: every dot product is exact in FP64 whatever the summation order, so the bucket-major
    kernels (DMMA) and the row-major kernel (FMA) must agree bit for bit — ids and scores."""
    X, Q = synth.config2(n=60_000, nq=512, d=128)
    A, chain = synth.angle_family(128, 128, 10, 3, 32, 88389)
    Ap = synth.partitioner_family(30, 3, 88390)
    ix = U.make_index(128, A, chain, Ap, bucket_overflow=100)
    ix.fit_dense(X)
    res = {}
    for name, opt in (("u8", {}), ("stream", {"bm_kernel": 1}), ("u8_dmma", {"u8_imma": 0}), ("rowmajor", {"rerank": 1}),
                      ("u8_tcgen05", {"u8i_kernel": 3}), ("u8_lean", {"u8i_kernel": 1})):
        with ix.debug_options(**opt):
            res[name] = ix.query_topk_dense(Q, None, 0, 10, B.METRIC_DOT)
        bm = ix.stats()["bm_pairs"]
        assert (bm > 0) == (name != "rowmajor")
    for name in ("u8", "stream", "u8_dmma", "u8_tcgen05", "u8_lean"):
        assert np.array_equal(res[name][0], res["rowmajor"][0]), name
        assert np.array_equal(res[name][1], res["rowmajor"][1]), name
    # a survivor pool far too small for the batch: the warps that find it full flag their queries, which are then answered
    # by the exhaustive per-query kernel — same results, no host round trip
    for pool in (256, 4096):
        with ix.debug_options(pool_records=pool):
            i2, s2 = ix.query_topk_dense(Q, None, 0, 10, B.METRIC_DOT)
        assert ix.stats()["bm_direct"] > 0, "expected queries to overflow the pool"
        assert np.array_equal(i2, res["rowmajor"][0]) and np.array_equal(s2, res["rowmajor"][1]), pool
    assert ix.query_topk_dense(Q, None, 0, 10, B.METRIC_DOT)[0].shape == (512, 10) and ix.stats()["bm_direct"] == 0


def test_topk_bucket_major_self_exclusion_and_qids():
    """qids given: the Integer-cache quirk (ids -128..127 never return themselves) also holds on the bucket-major path."""
    rng = np.random.default_rng(5)
    X = rng.standard_normal((2000, 64))
    A, chain, Ap = U.make_functions(64, family_size=64, table_num=4, permutation_num=2, seed=9)
    o = U.make_oracle(64, A, chain, Ap, bucket_overflow=50); ix = U.make_index(64, A, chain, Ap, bucket_overflow=50)
    o.fit_dense(X); ix.fit_dense(X)
    qids = np.concatenate([np.arange(0, 140), np.arange(1000, 1060)]).astype(np.int32)
    io, so = o.query_topk_dense(X[qids], qids, 2, 5, B.METRIC_DOT)
    ig, sg = ix.query_topk_dense(X[qids], qids, 2, 5, B.METRIC_DOT)
    U.assert_topk_close(io, so, ig, sg)
    assert not any(qids[i] in ig[i] for i in range(128))


# ---- K5c: bucket-major re-rank for d > 128 (rerank_wide.cu) ---------------------------------------------------------
@pytest.mark.parametrize("d", [130, 132, 200, 960])
@pytest.mark.parametrize("metric", [B.METRIC_DOT, B.METRIC_ANGULAR])
@pytest.mark.parametrize("topk", [10, 100])
def test_topk_wide_rows_bucket_major(d, metric, topk):
    """d > 128 (GIST shape): rows streamed once per unit of <= 16 queries.  Partial last column group (d = 130, 132),
    long runs of pairs (both n-blocks), buckets longer than one 32-row slab, register (k <= 32) and shared-memory lists;
    against the oracle, and the row-major kernel given the same candidates."""
    rng = np.random.default_rng(300 + d)
    centres = rng.standard_normal((12, d))
    X = np.repeat(centres, 250, axis=0) + 0.05 * rng.standard_normal((3000, d))
    A, chain, Ap = U.make_functions(d, family_size=max(40, min(d, 160)), table_num=3, permutation_num=2, seed=31 + d)
    o = U.make_oracle(d, A, chain, Ap, bucket_overflow=70); ix = U.make_index(d, A, chain, Ap, bucket_overflow=70)
    o.fit_dense(X); ix.fit_dense(X)
    Qs = X[::5] + 0.01 * rng.standard_normal((600, d))
    io, so = o.query_topk_dense(Qs, None, 1, topk, metric)
    ig, sg = ix.query_topk_dense(Qs, None, 1, topk, metric)
    U.assert_topk_close(io, so, ig, sg)
    st = ix.stats()
    assert st["bm_pairs"] > 16 * 3 * 2 and st["bm_rows_staged"] > 0, "the wide bucket-major path did not run"
    with ix.debug_options(wide=1):
        ir, sr = ix.query_topk_dense(Qs, None, 1, topk, metric)
    assert ix.stats()["bm_pairs"] == 0, "expected the row-major kernel"
    U.assert_topk_close(ir, sr, ig, sg)
    ix.close(); o.close()


def test_topk_wide_rows_corner_cases():
    """d > 128: integer-valued data (every dot product exact: bit-equal to the row-major kernel), self-exclusion with
    qids, queries whose samples hold fewer than k rows, a survivor pool too small for the batch (exhaustive fallback)."""
    d = 192
    rng = np.random.default_rng(77)
    X = np.rint(8 * rng.standard_normal((5000, d))) + 0.0
    X[2500:] = X[:2500] + np.rint(rng.standard_normal((2500, d)))
    A, chain, Ap = U.make_functions(d, family_size=96, table_num=4, permutation_num=2, seed=12)
    o = U.make_oracle(d, A, chain, Ap, bucket_overflow=40); ix = U.make_index(d, A, chain, Ap, bucket_overflow=40)
    o.fit_dense(X); ix.fit_dense(X)
    qids = np.concatenate([np.arange(0, 140), np.arange(2500, 2900)]).astype(np.int32)
    Qs = X[qids]
    for steps, k in ((0, 5), (2, 5), (3, 64)):
        io, so = o.query_topk_dense(Qs, qids, steps, k, B.METRIC_DOT)
        ig, sg = ix.query_topk_dense(Qs, qids, steps, k, B.METRIC_DOT)
        U.assert_topk_close(io, so, ig, sg)
        assert ix.stats()["bm_pairs"] > 0
        with ix.debug_options(wide=1):
            ir, sr = ix.query_topk_dense(Qs, qids, steps, k, B.METRIC_DOT)
        assert np.array_equal(ir, ig) and np.array_equal(sr, sg), (steps, k)
    assert not any(qids[i] in ig[i] for i in range(128))
    with ix.debug_options(pool_records=256):
        i2, s2 = ix.query_topk_dense(Qs, qids, 3, 64, B.METRIC_DOT)
    assert ix.stats()["bm_direct"] > 0, "expected queries to overflow the pool"
    assert np.array_equal(i2, ig) and np.array_equal(s2, sg)
    ix.close(); o.close()


# ---- compact store (store.cu): lossless uint8 / float32 copy of the rows for the re-rank ----------------------------
def _store_data(kind, n, d, seed):
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((12, d))
    X = np.repeat(centres, n // 12, axis=0) + 0.05 * rng.standard_normal((n // 12 * 12, d))
    if kind == B.STORE_KIND_U8:
        X = np.clip(np.rint(60 * X + 128), 0, 255) + 0.0          # + 0.0: rint can leave -0.0, which is not a byte
    elif kind == B.STORE_KIND_F32:
        X = X.astype(np.float32).astype(np.float64)
    return X


@pytest.mark.parametrize("d", [36, 100, 128])
@pytest.mark.parametrize("kind", [B.STORE_KIND_F64, B.STORE_KIND_F32, B.STORE_KIND_U8])
@pytest.mark.parametrize("metric", [B.METRIC_DOT, B.METRIC_ANGULAR])
def test_compact_store_kinds_match_oracle(d, kind, metric):
    """The fit picks the narrowest lossless element type; top-k through the bucket-major kernel that widens it back to
    FP64 equals the oracle on the FP64 rows (ids exact, scores 1e-12) — ragged last chunk at d = 36, 100."""
    X = _store_data(kind, 3000, d, 300 + d)
    rng = np.random.default_rng(7)
    A, chain, Ap = U.make_functions(d, family_size=max(40, d), table_num=3, permutation_num=2, seed=41 + d)
    o = U.make_oracle(d, A, chain, Ap, bucket_overflow=30)
    ix = U.make_index(d, A, chain, Ap, bucket_overflow=30, store_mode=B.STORE_NARROWEST)
    o.fit_dense(X); ix.fit_dense(X)
    st = ix.stats()
    assert st["store_kind"] == kind
    assert st["store_row_bytes"] == {B.STORE_KIND_F64: 8 * d, B.STORE_KIND_F32: 4 * d, B.STORE_KIND_U8: (d + 15) // 16 * 16}[kind]
    Qs = X[::5] + 0.01 * rng.standard_normal((600, d))        # FP64 queries whatever the store
    io, so = o.query_topk_dense(Qs, None, 1, 10, metric)
    ig, sg = ix.query_topk_dense(Qs, None, 1, 10, metric)
    U.assert_topk_close(io, so, ig, sg)
    assert ix.stats()["bm_pairs"] > 0, "the bucket-major path did not run"
    if kind == B.STORE_KIND_U8:
        # byte queries against byte rows: the integer tensor pipe computes the same exact dot products
        Qb = np.ascontiguousarray(X[3::7])
        io, so = o.query_topk_dense(Qb, None, 1, 10, metric)
        ig, sg = ix.query_topk_dense(Qb, None, 1, 10, metric)
        U.assert_topk_close(io, so, ig, sg)
        if metric == B.METRIC_DOT:
            assert np.array_equal(so[~np.isnan(so)], sg[~np.isnan(sg)]), "integer dot products must be exact"
            with ix.debug_options(u8i_kernel=3):                  # the tcgen05 / TMEM kernel: rows narrower than 128 bytes too
                it, st_ = ix.query_topk_dense(Qb, None, 1, 10, metric)
            assert np.array_equal(it, ig) and np.array_equal(st_, sg, equal_nan=True)
            assert not ix.tc_diag()[:8].any(), "tcgen05 kernel watchdog fired"


def test_compact_store_bitwise_equal_to_f64_rows_on_integer_data(monkeypatch):
    """uint8 rows widened in registers are the same doubles: with integer-valued queries every product and sum is
    exact, so the narrow store and the FP64 store must agree bit for bit."""
    X, Q = synth.config2(n=60_000, nq=512, d=128)
    A, chain = synth.angle_family(128, 128, 10, 3, 32, 88389)
    Ap = synth.partitioner_family(30, 3, 88390)
    res = {}
    for mode in (B.STORE_AUTO, B.STORE_F64_ONLY):
        ix = U.make_index(128, A, chain, Ap, bucket_overflow=100)
        ix.set_store_mode(mode)
        ix.fit_dense(X)
        assert ix.stats()["store_kind"] == (B.STORE_KIND_U8 if mode == B.STORE_AUTO else B.STORE_KIND_F64)
        res[mode] = [ix.query_topk_dense(Q, None, 0, 10, m) for m in (B.METRIC_DOT, B.METRIC_ANGULAR)]
        if mode == B.STORE_AUTO:
            # the same byte rows on the FP64 tensor pipe (DPF_U8_IMMA=0) and through the TMA ring kernel
            with ix.debug_options(u8_imma=0):
                res["u8_dmma"] = [ix.query_topk_dense(Q, None, 0, 10, m) for m in (B.METRIC_DOT, B.METRIC_ANGULAR)]
            with ix.debug_options(bm_kernel=1):
                res["u8_stream"] = [ix.query_topk_dense(Q, None, 0, 10, m) for m in (B.METRIC_DOT, B.METRIC_ANGULAR)]
        ix.close()
    for name in (B.STORE_AUTO, "u8_dmma", "u8_stream"):
        for a, b in zip(res[name], res[B.STORE_F64_ONLY]):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), name


@pytest.mark.parametrize("poison,expect", [(-0.0, B.STORE_KIND_F32), (256.0, B.STORE_KIND_F32), (0.5, B.STORE_KIND_F32),
                                           (0.1, B.STORE_KIND_F64), (1e300, B.STORE_KIND_F64), (None, B.STORE_KIND_U8)])
def test_compact_store_one_value_decides(poison, expect):
    """Never lossy: a single value that does not survive the round trip (including -0.0 for uint8) widens the store."""
    d = 32
    X = _store_data(B.STORE_KIND_U8, 1200, d, 5)
    if poison is not None:
        X[777, 13] = poison
    A, chain, Ap = U.make_functions(d, family_size=40, table_num=2, permutation_num=1, seed=3)
    ix = U.make_index(d, A, chain, Ap, bucket_overflow=30, store_mode=B.STORE_NARROWEST)
    ix.fit_dense(X)
    assert ix.stats()["store_kind"] == expect
    ix2 = U.make_index(d, A, chain, Ap, bucket_overflow=30)          # default mode: bytes or nothing
    ix2.fit_dense(X)
    assert ix2.stats()["store_kind"] == (B.STORE_KIND_U8 if expect == B.STORE_KIND_U8 else B.STORE_KIND_F64)


@pytest.mark.parametrize("metric", [B.METRIC_DOT, B.METRIC_ANGULAR, B.METRIC_L2])
@pytest.mark.parametrize("byte_queries", [True, False])
def test_threshold_filter_corner_cases(metric, byte_queries, monkeypatch):
    """Byte store -> filtered pipeline (k_threshold, k_score_u8*, k_select_survivors).  topk larger than the sampled
    buckets (tau = -inf: nothing may be filtered), topk larger than all candidates (padded rows), any number of sampled
    tables, qids with the self-exclusion quirk: always the oracle's top k."""
    d = 48
    X = _store_data(B.STORE_KIND_U8, 2400, d, 77)
    A, chain, Ap = U.make_functions(d, family_size=48, table_num=4, permutation_num=2, seed=78)
    o = U.make_oracle(d, A, chain, Ap, bucket_overflow=25); ix = U.make_index(d, A, chain, Ap, bucket_overflow=25)
    o.fit_dense(X); ix.fit_dense(X)
    assert ix.stats()["store_kind"] == B.STORE_KIND_U8
    qids = np.arange(0, 2400, 9, dtype=np.int32)
    Qs = X[qids] if byte_queries else X[qids] + 0.37
    for topk, steps, nt in ((5, 0, 1), (10, 1, 3), (60, 1, 8), (200, 2, 32), (256, 0, 6)):
        ix.set_debug_option(B.DBG_TAU_TABLES, nt)
        io, so = o.query_topk_dense(Qs, qids, steps, topk, metric)
        ig, sg = ix.query_topk_dense(Qs, qids, steps, topk, metric)
        U.assert_topk_close(io, so, ig, sg)
        st = ix.stats()
        if metric == B.METRIC_L2 and not byte_queries:        # squared L2 is exact only on the integer pipeline
            assert st["bm_pairs"] == 0, "expected the row-major kernel"
        else:
            # every query is served either by the filter (its survivors hold its result) or by the exhaustive kernel
            assert st["bm_pairs"] > 0 and st["bm_survivors"] + st["bm_direct"] * topk >= (ig >= 0).sum()
        if metric == B.METRIC_L2 and byte_queries:
            assert np.array_equal(so[~np.isnan(so)], sg[~np.isnan(sg)]), "integer squared distances must be exact"
    assert np.array_equal(ig == -1, io == -1)                 # padded rows where the candidates run out


@pytest.mark.parametrize("store", ["u8", "f64"])
def test_query_batch_cut_into_memory_bounded_chunks(store, monkeypatch):
    """A tiny candidate budget forces the batch through dozens of chunks (expand / re-rank / select per chunk): results
    must not depend on where the batch is cut — candidate sets and top-k, byte pipeline and FP64 pipeline."""
    d = 64
    X = _store_data(B.STORE_KIND_U8 if store == "u8" else B.STORE_KIND_F64, 3600, d, 91)
    A, chain, Ap = U.make_functions(d, family_size=64, table_num=4, permutation_num=2, seed=92)
    ix = U.make_index(d, A, chain, Ap, bucket_overflow=30)
    ix.fit_dense(X)
    Qs = X[::7] if store == "u8" else X[::7] + 0.01
    ref_c = ix.query_candidates_dense(Qs, None, 1)
    ref = [ix.query_topk_dense(Qs, None, 1, 10, m) for m in (B.METRIC_DOT, B.METRIC_ANGULAR, B.METRIC_L2)]
    ix.set_debug_option(B.DBG_CAND_BUDGET, 20000)
    U.assert_csr_equal(ref_c, ix.query_candidates_dense(Qs, None, 1))
    for m, (ri, rs) in zip((B.METRIC_DOT, B.METRIC_ANGULAR, B.METRIC_L2), ref):
        gi, gs = ix.query_topk_dense(Qs, None, 1, 10, m)
        assert np.array_equal(ri, gi) and np.array_equal(rs, gs, equal_nan=True), m
    o = U.make_oracle(d, A, chain, Ap, bucket_overflow=30)
    o.fit_dense(X)
    io, so = o.query_topk_dense(Qs, None, 1, 10, B.METRIC_DOT)
    U.assert_topk_close(io, so, *ix.query_topk_dense(Qs, None, 1, 10, B.METRIC_DOT))


@pytest.mark.parametrize("kind", [B.STORE_KIND_U8, B.STORE_KIND_F64])
def test_tiny_batches_and_tiny_indexes(kind):
    """One query, two queries, an index smaller than k: grids of one CTA, lists that never fill, padded results."""
    d = 32
    A, chain, Ap = U.make_functions(d, family_size=40, table_num=3, permutation_num=1, seed=55)
    for n in (12, 300):
        X = _store_data(kind, n, d, 60 + n)
        o = U.make_oracle(d, A, chain, Ap, bucket_overflow=10); ix = U.make_index(d, A, chain, Ap, bucket_overflow=10)
        o.fit_dense(X); ix.fit_dense(X)
        assert ix.stats()["store_kind"] == kind
        for nq in (1, 2, 5):
            Qs = np.ascontiguousarray(X[:nq]) if kind == B.STORE_KIND_U8 else X[:nq] + 0.125
            for metric in (B.METRIC_DOT, B.METRIC_ANGULAR, B.METRIC_L2):
                for steps in (0, 3):
                    io, so = o.query_topk_dense(Qs, None, steps, 10, metric)
                    ig, sg = ix.query_topk_dense(Qs, None, steps, 10, metric)
                    U.assert_topk_close(io, so, ig, sg)
            U.assert_csr_equal(o.query_candidates_dense(Qs, None, 1), ix.query_candidates_dense(Qs, None, 1))


def test_compact_store_append_widens():
    """Appending vectors that are not bytes re-types the store; results still equal the oracle's."""
    d = 64
    X1 = _store_data(B.STORE_KIND_U8, 1800, d, 11)
    X2 = _store_data(B.STORE_KIND_F64, 600, d, 12) * 60 + 128
    A, chain, Ap = U.make_functions(d, family_size=64, table_num=3, permutation_num=1, seed=13)
    o = U.make_oracle(d, A, chain, Ap, bucket_overflow=30); ix = U.make_index(d, A, chain, Ap, bucket_overflow=30)
    o.fit_dense(X1); ix.fit_dense(X1)
    assert ix.stats()["store_kind"] == B.STORE_KIND_U8
    o.fit_dense(X2); ix.fit_dense(X2)
    assert ix.stats()["store_kind"] == B.STORE_KIND_F64
    Qs = np.concatenate([X1[::30], X2[::10]]) + 0.25
    io, so = o.query_topk_dense(Qs, None, 1, 10, B.METRIC_DOT)
    ig, sg = ix.query_topk_dense(Qs, None, 1, 10, B.METRIC_DOT)
    U.assert_topk_close(io, so, ig, sg)


# ---- error behaviour -----------------------------------------------------------------------------------------
def test_error_codes():
    from similaritysearchbyrdf_b200 import DPFIndex
    ix = DPFIndex(d=4, L=2, k=32)
    with pytest.raises(B.DpfError) as e:
        ix.fit_dense(np.zeros((3, 4)))                        # family not set
    assert e.value.code == B.ERR_STATE
    A, chain, Ap = U.make_functions(4, family_size=40, table_num=2, permutation_num=1)
    ix.set_family(A, chain); ix.set_partitioners(Ap)
    with pytest.raises(B.DpfError) as e:
        ix.query_topk_dense(np.zeros((1, 4)))                 # "need to fit the data first"
    assert e.value.code == B.ERR_STATE
    ix.fit_dense(np.random.default_rng(0).standard_normal((100, 4)))
    with pytest.raises(B.DpfError) as e:
        ix.query_candidates_by_id(np.array([100], np.int32))  # unknown id
    assert e.value.code == B.ERR_INVALID
