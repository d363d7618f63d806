"""Incremental put and remove on the flat forest (incremental.cu) against the oracle's sequential RandomDrawTreeMap.put /
remove (RandomDrawTreeMap.java:1558-1584, 1662-1790, 1817-1932): bucket membership, split counters, candidate sets and
top-k must equal the oracle's after every operation — the tree is the same one the reference's call sequence produces."""
import numpy as np
import pytest

from similaritysearchbyrdf_b200 import _lib as B
from similaritysearchbyrdf_b200 import synth
from tests import util as U

pytestmark = pytest.mark.gpu


def _pair(d, A, chain, Ap, **kw):
    return U.make_oracle(d, A, chain, Ap, **kw), U.make_index(d, A, chain, Ap, **kw)


def _same(o, ix, L):
    U.assert_buckets_equal(o, ix, L)
    so, sg = o.stats(), ix.stats()
    assert so["splits"] == sg["splits"] and so["singleton_splits"] == sg["singleton_splits"]


@pytest.mark.parametrize("T", [40, 6])
def test_small_appends_are_put_one_by_one(T):
    X, Q = synth.config1(n=9000)
    A, chain, Ap = U.make_functions(100)
    o, ix = _pair(100, A, chain, Ap, bucket_overflow=T)
    o.fit_dense(X[:6000]); ix.fit_dense(X[:6000])
    ix.set_debug_option(B.DBG_APPEND, 2)                       # always incremental (the default would already be)
    at = 6000
    for m in (1, 2, 30, 64, 300, 7, 1000):
        o.fit_dense(X[at:at + m]); ix.fit_dense(X[at:at + m])
        at += m
        assert len(ix) == at
        _same(o, ix, chain.shape[0])
    Qs = np.concatenate([Q, X[5990:6010] + 0.01])
    U.assert_csr_equal(o.query_candidates_dense(Qs, None, 1), ix.query_candidates_dense(Qs, None, 1))
    io, so = o.query_topk_dense(Qs, None, 1, 10, B.METRIC_ANGULAR)
    U.assert_topk_close(io, so, *ix.query_topk_dense(Qs, None, 1, 10, B.METRIC_ANGULAR))


@pytest.mark.parametrize("dir_node_size,T", [(128, 4), (4, 2), (2, 1)])
def test_incremental_put_with_skewed_keys_and_deep_splits(dir_node_size, T):
    d = 8
    A, chain, Ap = U.make_functions(d, family_size=16, table_num=3, permutation_num=2, seed=5)
    rng = np.random.default_rng(6)
    C = rng.standard_normal((12, d))
    X = C[rng.integers(0, 12, 4000)] + 0.02 * rng.standard_normal((4000, d))
    o, ix = _pair(d, A, chain, Ap, bucket_overflow=T, dir_node_size=dir_node_size)
    o.fit_dense(X[:2000]); ix.fit_dense(X[:2000])
    ix.set_debug_option(B.DBG_APPEND, 2)
    for lo, hi in ((2000, 2003), (2003, 2500), (2500, 4000)):   # many ids per slot, splits inside one run of a warp
        o.fit_dense(X[lo:hi]); ix.fit_dense(X[lo:hi])
        _same(o, ix, chain.shape[0])


def test_remove_then_put_equals_the_reference_sequence():
    X, Q = synth.config1(n=8000)
    A, chain, Ap = U.make_functions(100)
    L = chain.shape[0]
    o, ix = _pair(100, A, chain, Ap, bucket_overflow=12)
    o.fit_dense(X[:6000]); ix.fit_dense(X[:6000])
    rng = np.random.default_rng(1)
    gone = rng.choice(6000, 1500, replace=False).astype(np.int32)
    assert o.remove(gone) == ix.remove(gone) == 1500 * L
    U.assert_buckets_equal(o, ix, L)
    assert ix.remove(gone[:10]) == 0 == o.remove(gone[:10])    # already gone: ignored
    assert ix.remove(np.array([10 ** 6, -5], np.int32)) == 0     # never inserted: ignored
    Qs = np.concatenate([Q, X[:40] + 0.01])
    U.assert_csr_equal(o.query_candidates_dense(Qs, None, 1), ix.query_candidates_dense(Qs, None, 1))
    io, so = o.query_topk_dense(Qs, None, 1, 10, B.METRIC_DOT)
    ig, sg = ix.query_topk_dense(Qs, None, 1, 10, B.METRIC_DOT)
    U.assert_topk_close(io, so, ig, sg)
    assert not np.isin(ig, gone).any()
    # puts after removes land in the tree the removes left behind (buckets below the split limit again, freed slots)
    ix.set_debug_option(B.DBG_APPEND, 2)
    for lo, hi in ((6000, 6002), (6002, 6400), (6400, 8000)):
        o.fit_dense(X[lo:hi]); ix.fit_dense(X[lo:hi])
        U.assert_buckets_equal(o, ix, L)
    again = np.arange(5000, 7000, dtype=np.int32)
    assert o.remove(again) == ix.remove(again)
    U.assert_buckets_equal(o, ix, L)
    U.assert_csr_equal(o.query_candidates_dense(Qs, None, 1), ix.query_candidates_dense(Qs, None, 1))


@pytest.mark.parametrize("dir_node_size,T", [(4, 2), (2, 1)])
def test_remove_collapses_empty_directories(dir_node_size, T):
    """Deep, narrow trees: removing whole clusters empties directory nodes, which must disappear from their parents
    (recursiveDirDelete) — the puts that follow create their buckets at the level the reference would."""
    d = 8
    A, chain, Ap = U.make_functions(d, family_size=16, table_num=3, permutation_num=2, seed=5)
    rng = np.random.default_rng(8)
    C = rng.standard_normal((10, d))
    lab = rng.integers(0, 10, 3000)
    X = C[lab] + 0.02 * rng.standard_normal((3000, d))
    o, ix = _pair(d, A, chain, Ap, bucket_overflow=T, dir_node_size=dir_node_size)
    o.fit_dense(X[:2400]); ix.fit_dense(X[:2400])
    gone = np.nonzero(lab[:2400] < 6)[0].astype(np.int32)        # six of the ten clusters vanish
    assert o.remove(gone) == ix.remove(gone)
    U.assert_buckets_equal(o, ix, chain.shape[0])
    ix.set_debug_option(B.DBG_APPEND, 2)
    o.fit_dense(X[2400:]); ix.fit_dense(X[2400:])                # clusters 0..5 come back into collapsed regions
    U.assert_buckets_equal(o, ix, chain.shape[0])
    rest = np.arange(0, 3000, dtype=np.int32)
    assert o.remove(rest) == ix.remove(rest)                     # everything gone: only the empty roots are left
    for t in range(chain.shape[0]):
        assert len(ix.dump_buckets(t)[2]) == 0


def test_two_vector_insert_on_a_large_index_is_incremental_by_default():
    """The facade overload that first inserts its query vectors (DensevectorRDFInit.scala:215-251, 372-399)."""
    X, Q = synth.config2(n=200_000, nq=64, d=128)
    A, chain = synth.angle_family(128, 128, 10, 3, 32, 88389)
    Ap = synth.partitioner_family(30, 3, 88390)
    o, ix = _pair(128, A, chain, Ap)
    o.fit_dense(X); ix.fit_dense(X)
    launches = ix.stats()["kernel_launches"]
    o.fit_dense(Q[:2]); ix.fit_dense(Q[:2])
    assert ix.stats()["kernel_launches"] - launches < 40, "a two-vector insert must not rebuild the forest"
    assert ix.stats()["store_kind"] == B.STORE_KIND_U8
    U.assert_buckets_equal(o, ix, chain.shape[0])
    qid = np.array([200_000, 200_001], np.int32)
    U.assert_csr_equal(o.query_candidates_by_id(qid, 0), ix.query_candidates_by_id(qid, 0))
    io, so = o.query_topk_dense(Q[:16], None, 0, 10, B.METRIC_DOT)
    ig, sg = ix.query_topk_dense(Q[:16], None, 0, 10, B.METRIC_DOT)
    U.assert_topk_close(io, so, ig, sg)
