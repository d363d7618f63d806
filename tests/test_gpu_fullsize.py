"""BASELINE configs[1] at full size (1M x 128, 10k-query batch, k = 10) through the C ABI: the oracle needs minutes at
this size, so the checks are size-independent properties of the path — agreement of independent kernels, idempotence,
order and uniqueness of results, self-retrieval, append == one-shot build, save/load round trip."""
import numpy as np
import pytest

from similaritysearchbyrdf_b200 import _lib as B
from similaritysearchbyrdf_b200 import synth
from tests import util as U

pytestmark = pytest.mark.gpu

N, NQ, D, K = 1_000_000, 10_000, 128, 10


@pytest.fixture(scope="module")
def full():
    X, Q = synth.config2(N, NQ, D)
    A, chain = synth.angle_family(D, max(100, D), 10, 3, 32, 88387 + 2)
    Ap = synth.partitioner_family(chain.shape[0], 3, 88387 + 3)
    ix = U.make_index(D, A, chain, Ap)
    ix.fit_dense(X)
    return X, Q, A, chain, Ap, ix


def _bucket_checksum(ix, tables):
    """order-sensitive checksum of the canonical bucket dump (descriptors, sizes, ids) of some tables"""
    acc = 0
    for t in tables:
        desc, off, ids = ix.dump_buckets(t)
        for a in (desc, off, ids):
            a = np.ascontiguousarray(a).view(np.uint8)
            acc = (acc * 1000003 + int(np.frombuffer(a.tobytes(), np.uint8).astype(np.uint64).sum()) + a.size) % (1 << 61)
        w = np.arange(1, len(ids) + 1, dtype=np.uint64)
        acc = (acc + int((ids.astype(np.uint64) * w).sum() % (1 << 61))) % (1 << 61)
    return acc


def test_full_size_topk_properties_and_kernel_agreement(full, monkeypatch):
    X, Q, A, chain, Ap, ix = full
    st = ix.stats()
    assert st["size"] == N and st["store_kind"] == B.STORE_KIND_U8
    ids, sc = ix.query_topk_dense(Q, None, 0, K, B.METRIC_DOT)
    assert ix.stats()["bm_survivors"] >= NQ * K
    # order, uniqueness, padding
    valid = ids >= 0
    assert valid.all(), "every query of this workload has at least k candidates"
    assert (np.diff(sc, axis=1) <= 0).all(), "scores must be non-increasing"
    assert all(len(set(r)) == K for r in ids[:2000]), "an id may appear once per query"
    # the scores are the exact FP64 dot products of the returned rows (integer data: exact in any order)
    ref = np.einsum("qkd,qd->qk", X[ids[:512]], Q[:512])
    assert np.array_equal(ref, sc[:512])
    # idempotence
    ids2, sc2 = ix.query_topk_dense(Q, None, 0, K, B.METRIC_DOT)
    assert np.array_equal(ids, ids2) and np.array_equal(sc, sc2)
    # independent kernels: integer tensor pipe + threshold filter / FP64 tensor pipe on byte rows / TMA ring / row-major
    for env in ({"DPF_U8_IMMA": "0"}, {"DPF_BM_KERNEL": "stream"}, {"DPF_RERANK": "rowmajor"}, {"DPF_U8I_KERNEL": "lean"},
                {"DPF_TAU_TABLES": "1"}, {"DPF_TAU_TABLES": "30"}):
        for k_, v in env.items():
            monkeypatch.setenv(k_, v)
        i3, s3 = ix.query_topk_dense(Q, None, 0, K, B.METRIC_DOT)
        for k_ in env:
            monkeypatch.delenv(k_)
        assert np.array_equal(ids, i3) and np.array_equal(sc, s3), env
    # angular: cosine of the returned rows, within the north-star tolerance, and the two pipelines agree
    ia, sa = ix.query_topk_dense(Q[:2000], None, 0, K, B.METRIC_ANGULAR)
    cos = np.einsum("qkd,qd->qk", X[ia], Q[:2000]) / (np.linalg.norm(X[ia], axis=2) * np.linalg.norm(Q[:2000], axis=1)[:, None])
    assert np.all(np.abs(cos - sa) <= 1e-12 * np.abs(cos))
    monkeypatch.setenv("DPF_RERANK", "rowmajor")
    ir, sr = ix.query_topk_dense(Q[:2000], None, 0, K, B.METRIC_ANGULAR)
    monkeypatch.delenv("DPF_RERANK")
    U.assert_topk_close(ir, sr, ia, sa)
    # squared L2 on the integer pipeline: exact, ascending, equal to the row-major kernel bit for bit
    il, sl = ix.query_topk_dense(Q, None, 0, K, B.METRIC_L2)
    assert ix.stats()["bm_survivors"] >= NQ * K, "L2 on byte rows and byte queries runs the filtered pipeline"
    assert (np.diff(sl, axis=1) >= 0).all()
    assert np.array_equal(((X[il[:512]] - Q[:512, None, :]) ** 2).sum(axis=2), sl[:512])
    monkeypatch.setenv("DPF_RERANK", "rowmajor")
    ir, sr = ix.query_topk_dense(Q, None, 0, K, B.METRIC_L2)
    monkeypatch.delenv("DPF_RERANK")
    assert np.array_equal(il, ir) and np.array_equal(sl, sr)


def test_full_size_self_retrieval_and_candidate_sets(full):
    X, Q, A, chain, Ap, ix = full
    qids = np.arange(500, N, N // 2000, dtype=np.int32)                   # ids > 127: no self-exclusion quirk
    off, cand = ix.query_candidates_dense(X[qids], qids, 0)
    for i in range(0, len(qids), 37):
        c = cand[off[i]:off[i + 1]]
        assert (np.diff(c) > 0).all(), "candidate sets are sorted and unique"
        assert qids[i] in c, "a stored vector probes its own bucket (a flipped bit below the leaf level)"
    ids, sc = ix.query_topk_dense(X[qids], qids, 0, K, B.METRIC_ANGULAR)
    assert (ids[:, 0] == qids).mean() > 0.99 and np.allclose(sc[ids[:, 0] == qids, 0], 1.0, atol=1e-12)


def test_full_size_append_equals_one_shot_and_save_load(full, tmp_path):
    X, Q, A, chain, Ap, ix = full
    from similaritysearchbyrdf_b200 import DPFIndex
    tables = (0, 13, 29)
    ref = _bucket_checksum(ix, tables)
    ix2 = U.make_index(D, A, chain, Ap)
    ix2.fit_dense(X[:400_000]); ix2.fit_dense(X[400_000:])               # sequential insertion order is what defines the forest
    assert ix2.stats()["splits"] == ix.stats()["splits"] and ix2.stats()["dir_nodes"] == ix.stats()["dir_nodes"]
    assert _bucket_checksum(ix2, tables) == ref
    path = tmp_path / "full.dpf"
    ix2.save(path)
    ix2.close()
    ix3 = DPFIndex.load(path)
    assert _bucket_checksum(ix3, tables) == ref
    a = ix.query_topk_dense(Q[:3000], None, 0, K, B.METRIC_DOT)
    b = ix3.query_topk_dense(Q[:3000], None, 0, K, B.METRIC_DOT)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    ix3.close()
