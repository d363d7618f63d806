"""GPU tests written the way the reference's own suites use its facades (TestSingleRDFSuite.scala, README "simple
test"), plus the two-shard emulation of the content-based partition scheme on one GPU."""
import os

import numpy as np
import pytest

from similaritysearchbyrdf_b200 import _lib as B
from similaritysearchbyrdf_b200 import synth
from tests import util as U

pytestmark = pytest.mark.gpu


def _write_dense_file(path, X):
    with open(path, "w") as f:
        for i, row in enumerate(X):
            f.write("[%d, [%s]]\n" % (i + 7, ", ".join(repr(float(v)) for v in row)))   # file ids are ignored (Q14)


def _write_sparse_file(path, indptr, idx, val, D):
    with open(path, "w") as f:
        for i in range(len(indptr) - 1):
            s, e = indptr[i], indptr[i + 1]
            f.write("[%d, %d, [%s], [%s]]\n" % (i, D, ", ".join(str(int(v)) for v in idx[s:e]),
                                                ", ".join(repr(float(v)) for v in val[s:e])))


def test_deploy_dense_simple_test(tmp_path):
    """README.md:32-42 / TestSingleRDFSuite: newMultiThreadFit on a file, NewMultiThreadQueryBatch of 2 queries."""
    from similaritysearchbyrdf_b200 import deploy as D
    X, Q = synth.config1(n=3000)
    path = os.path.join(tmp_path, "dense.txt")
    _write_dense_file(path, X)
    conf = D.Config.parseString("mclab.lshTable.bufferOverflow=40").withFallback(D.testBaseConf)
    D.LSHServer.lshEngine = D.LSH(conf)
    D.LSHServer.isUseDense = True
    all_vectors = D.DensevectorRDFInit.newMultiThreadFit(path, conf)
    assert len(all_vectors) == 3000 and all_vectors[5].vectorId == 5
    queries = [D.DenseVector(0, Q[0]), D.DenseVector(1, Q[1])]
    res = D.DensevectorRDFInit.NewMultiThreadQueryBatch([0, 1], queries, 0, 5)
    assert len(res) == 2 and all(isinstance(r, set) for r in res)
    # multi-thread query == single-thread query (TestSingleRDFSuite.scala:57-60): thread count is immaterial here
    ids = list(range(100))
    vecs = [all_vectors[i] for i in ids]
    a = D.DensevectorRDFInit.NewMultiThreadQueryBatch(ids, vecs, 0, 5)
    b = D.DensevectorRDFInit.queryBatch(ids, vecs, 0)
    assert a == b
    # same functions through the oracle give the same candidate sets
    lsh = D.LSHServer.lshEngine
    o = U.make_oracle(100, lsh.A, lsh.chain, D.DensevectorRDFInit.partitioners, bucket_overflow=40)
    o.fit_dense(X)
    off, cand = o.query_candidates_dense(np.stack([v.values for v in vecs]), np.array(ids, np.int32), 0)
    assert [set(cand[off[i]:off[i + 1]].tolist()) for i in range(100)] == a
    # the overload that takes only vectors inserts them first (DensevectorRDFInit.scala:372-399)
    r2 = D.DensevectorRDFInit.NewMultiThreadQueryBatch(queries, 0, 5)
    assert len(D.DensevectorRDFInit.index) == 3002 and len(r2) == 2
    assert 3000 in r2[0] and 3001 in r2[1]                   # a query now finds itself (ids outside -128..127)
    topk, precision = D.DensevectorRDFInit.topKAndPrecisionScore(all_vectors, [set(range(10))] * 3, conf, 1)
    assert len(topk) == 3 and 0.0 <= precision <= 1.0
    dt, ht = D.DensevectorRDFInit.getDtAndHtNumDistribution()
    assert abs(ht.sum() - 3002) < 1e-6
    D.DensevectorRDFInit.clearAndClose()
    assert D.DensevectorRDFInit.NewMultiThreadQueryBatch([0], [queries[0]], 0, 5) is None   # "need to fit the data first"
    D.LSHServer.lshEngine = None


def test_deploy_dense_pstable_family(tmp_path):
    """mclab.lsh.name = pStable through the facade: pStable keys AND pStable partitioner chains (confForPartitioner falls
    back to the main conf, DensevectorRDFInit.scala:63-70), candidate sets equal to the oracle's with the same functions."""
    from similaritysearchbyrdf_b200 import deploy as D
    X, Q = synth.config1(n=3000)
    path = os.path.join(tmp_path, "dense.txt")
    _write_dense_file(path, X)
    conf = D.Config.parseString("mclab.lsh.name=pStable\nmclab.lshTable.bufferOverflow=40").withFallback(D.testBaseConf)
    D.LSHServer.lshEngine = D.LSH(conf)
    D.LSHServer.isUseDense = True
    all_vectors = D.DensevectorRDFInit.newMultiThreadFit(path, conf)
    ids = list(range(100))
    vecs = [all_vectors[i] for i in ids]
    a = D.DensevectorRDFInit.NewMultiThreadQueryBatch(ids, vecs, 1, 5)
    lsh = D.LSHServer.lshEngine
    L = lsh.chain.shape[0]
    Ap, pb_b, pb_w = synth.pstable_partitioner_family(L, conf.getInt("mclab.lsh.partitionBits"), 0.0, 1.0, 4,
                                                      int(conf.get("mclab.lsh.seed", 88387)) + 1)
    assert np.array_equal(Ap, D.DensevectorRDFInit.partitioners)
    o = U.make_oracle(100, lsh.A, lsh.chain, Ap, bucket_overflow=40, family_kind=1, b=lsh.b, w=lsh.w, pb_b=pb_b, pb_w=pb_w)
    o.fit_dense(X)
    off, cand = o.query_candidates_dense(np.stack([v.values for v in vecs]), np.array(ids, np.int32), 1)
    assert [set(cand[off[i]:off[i + 1]].tolist()) for i in range(100)] == a
    D.LSHServer.lshEngine = None


def test_deploy_sparse_facade(tmp_path):
    from similaritysearchbyrdf_b200 import deploy as D
    Dm = 300
    rng = np.random.default_rng(3)
    nnz = np.maximum(rng.poisson(10, 2000), 1)
    indptr = np.concatenate([[0], np.cumsum(nnz)]).astype(np.int64)
    idx = np.concatenate([np.sort(rng.choice(Dm, m, replace=False)) for m in nnz]).astype(np.int32)
    val = rng.random(indptr[-1]) + 0.01
    path = os.path.join(tmp_path, "sparse.txt")
    _write_sparse_file(path, indptr, idx, val, Dm)
    conf = D.Config.parseString(f"mclab.lsh.vectorDim={Dm}\nmclab.lshTable.bufferOverflow=20").withFallback(D.testBaseConf)
    D.LSHServer.lshEngine = D.LSH(conf)
    vals = D.SparsevectorRDFInit.newMultiThreadFit(path, conf)
    assert len(vals) == 2000 and D.LSHServer.isUseDense is False
    ids = list(range(0, 2000, 40))
    res = D.SparsevectorRDFInit.NewMultiThreadQueryBatch(ids, 1, 5)
    lsh = D.LSHServer.lshEngine
    o = U.make_oracle(Dm, lsh.A, lsh.chain, D.SparsevectorRDFInit.partitioners, bucket_overflow=20)
    o.fit_csr(indptr, idx, val)
    off, cand = o.query_candidates_by_id(np.array(ids, np.int32), 1)
    assert [set(cand[off[i]:off[i + 1]].tolist()) for i in range(len(ids))] == res
    D.SparsevectorRDFInit.clearAndClose()
    D.LSHServer.lshEngine = None


def test_two_rank_emulation_on_one_gpu():
    """Two handles (rank 0/1 of world 2) on the same device + dpf_merge_topk_dev == the single-handle result."""
    import torch
    X, Q = synth.config1(n=8000)
    A, chain, Ap = U.make_functions(100)
    Qs = np.concatenate([Q, X[:100] + 0.01])
    nq, K = len(Qs), 10
    full = U.make_index(100, A, chain, Ap, bucket_overflow=40)
    full.fit_dense(X)
    for metric in (B.METRIC_DOT, B.METRIC_L2):
        f_ids, f_sc = full.query_topk_dense(Qs, None, 1, K, metric)
        shards = [U.make_index(100, A, chain, Ap, bucket_overflow=40, rank=r, world=2) for r in range(2)]
        per = []
        for s in shards:
            s.fit_dense(X)
            per.append(s.query_topk_dense(Qs, None, 1, K, metric))
        g_ids = torch.from_numpy(np.stack([p[0] for p in per])).cuda()
        g_sc = torch.from_numpy(np.stack([p[1] for p in per])).cuda()
        m_ids = torch.empty((nq, K), dtype=torch.int32, device="cuda")
        m_sc = torch.empty((nq, K), dtype=torch.float64, device="cuda")
        shards[0].merge_topk_dev(g_ids.data_ptr(), g_sc.data_ptr(), 2, nq, K, metric, m_ids.data_ptr(), m_sc.data_ptr())
        shards[0].sync()
        torch.cuda.synchronize()
        # (real-valued data: a query may be answered by different kernels on a shard and on the full index — the filter
        # or the exhaustive per-query kernel — so scores agree to the north-star tolerance, not bit for bit)
        U.assert_topk_close(f_ids, f_sc, m_ids.cpu().numpy(), m_sc.cpu().numpy())
        # each shard equals the oracle restricted to the same sub-indexes
        for r, s in enumerate(shards):
            o = U.make_oracle(100, A, chain, Ap, bucket_overflow=40, rank=r, world=2)
            o.fit_dense(X)
            U.assert_buckets_equal(o, s, chain.shape[0])
            U.assert_csr_equal(o.query_candidates_dense(Qs, None, 1), s.query_candidates_dense(Qs, None, 1))


def test_full_size_properties_config2_sample():
    """Size-independent properties at a BASELINE-sized slice: idempotent hashing, every id in exactly one bucket per
    table, candidate lists sorted unique, a query that is an indexed vector finds itself."""
    X, Q = synth.config2(n=200_000, nq=64)
    A, chain = synth.angle_family(128, 128, 10, 3, 32, 88389)
    Ap = synth.partitioner_family(30, 3, 88390)
    ix = U.make_index(128, A, chain, Ap)
    k1, p1 = ix.hash_dense(X[:5000])
    k2, p2 = ix.hash_dense(X[:5000])
    assert np.array_equal(k1, k2) and np.array_equal(p1, p2)
    ix.fit_dense(X)
    for t in (0, 17, 29):
        desc, off, ids = ix.dump_buckets(t)
        assert len(ids) == 200_000 and np.array_equal(np.sort(ids), np.arange(200_000))
        assert all(np.all(np.diff(ids[off[i]:off[i + 1]]) > 0) for i in range(0, len(off) - 1, 97))
        sizes = np.diff(off)
        assert (sizes[desc[:, 1] >= 1] <= 501).all()         # only level-0 buckets may exceed T+1
    # probe list {h}: the own bucket is always visited (the dense multi-probe list omits h itself, quirk Q4)
    off, cand = ix.query_candidates_dense(X[1000:1064], np.arange(1000, 1064, dtype=np.int32), 0, B.PROBE_NONE)
    for i in range(64):
        c = cand[off[i]:off[i + 1]]
        assert np.all(np.diff(c) > 0) and (1000 + i) in c
    off2, cand2 = ix.query_candidates_dense(X[1000:1064], None, 0)
    assert all(np.all(np.diff(cand2[off2[i]:off2[i + 1]]) > 0) for i in range(64))
    ids, sc = ix.query_topk_dense(X[1000:1064], None, 0, 10, B.METRIC_L2, B.PROBE_NONE)
    assert (ids[:, 0] == np.arange(1000, 1064)).all() and (sc[:, 0] == 0).all()
    assert np.all(np.diff(sc, axis=1) >= 0)


def test_deploy_lsh_calculate_index_and_views(tmp_path):
    """LSHServer.lshEngine.calculateIndex (LSH.scala:93-166) and the public vars vectorDatabase / vectorIdToVector."""
    from similaritysearchbyrdf_b200 import deploy
    rng = np.random.default_rng(2)
    X = rng.standard_normal((500, 16))
    path = tmp_path / "dense.txt"
    with open(path, "w") as f:
        for i, row in enumerate(X):
            f.write(f"[{i},[" + ",".join(repr(float(v)) for v in row) + "]]\n")
    conf = deploy.Config.parseString("mclab.lsh.vectorDim=16\nmclab.lsh.tableNum=3\nmclab.lsh.permutationNum=2\n"
                                     "mclab.lsh.familySize=40\nmclab.lshTable.bufferOverflow=20").withFallback(deploy.testBaseConf)
    deploy.LSHServer.lshEngine = None
    vecs = deploy.DensevectorRDFInit.newMultiThreadFit(str(path), conf)
    try:
        lsh = deploy.LSHServer.getLSHEngine()
        o = U.make_oracle(16, lsh.A, lsh.chain, deploy.DensevectorRDFInit.partitioners, bucket_overflow=20)
        ko, _ = o.hash_dense(X[:7])
        for i in range(7):
            assert np.array_equal(lsh.calculateIndex(vecs[i], -1), ko[:, i])
            assert lsh.calculateIndex(vecs[i], 4)[0] == ko[4, i]
        assert len(deploy.DensevectorRDFInit.vectorIdToVector) == 500
        tables = deploy.DensevectorRDFInit.vectorDatabase
        assert len(tables) == 6 and tables[0].size() == 500 and len(tables[2].dump()[2]) == 500
    finally:
        deploy.DensevectorRDFInit.clearAndClose()
        deploy.LSHServer.lshEngine = None


@pytest.mark.parametrize("world", [2, 3, 4])
def test_balanced_partition_ranks_merge_to_the_unsharded_result(world):
    """dpf_set_balanced_partition: sub-indexes dealt to the ranks by occupancy.  Every sub-index is owned by exactly one
    rank (candidate sets are disjoint and their union is the unsharded set), the merged top-k equals the single-handle
    result, and the load is no worse than p % world."""
    import torch
    X, Q = synth.config1(n=8000)
    A, chain, Ap = U.make_functions(100)
    Qs = np.concatenate([Q, X[:100] + 0.01])
    nq, K = len(Qs), 10
    full = U.make_index(100, A, chain, Ap, bucket_overflow=40)
    full.fit_dense(X)
    f_ids, f_sc = full.query_topk_dense(Qs, None, 1, K, B.METRIC_DOT)
    f_off, f_cand = full.query_candidates_dense(Qs, None, 1)
    loads = {}
    for balanced in (False, True):
        shards = [U.make_index(100, A, chain, Ap, bucket_overflow=40, rank=r, world=world) for r in range(world)]
        per, cands, sizes = [], [], []
        for s in shards:
            s.set_balanced_partition(balanced)
            s.fit_dense(X)
            per.append(s.query_topk_dense(Qs, None, 1, K, B.METRIC_DOT))
            cands.append(s.query_candidates_dense(Qs, None, 1))
            sizes.append(sum(len(s.dump_buckets(t)[2]) for t in range(chain.shape[0])))
        assert sum(sizes) == 8000 * chain.shape[0]                   # every (table, id) entry lives on exactly one rank
        owned = np.stack([s.owned_subindexes() for s in shards])
        assert (owned.sum(axis=0) == 1).all()                          # every sub-index has exactly one owner
        if balanced:                                                   # each shard = the oracle restricted to the same sub-indexes
            for r, s in enumerate(shards):
                o = U.make_oracle(100, A, chain, Ap, bucket_overflow=40)
                o.set_owned(owned[r])
                o.fit_dense(X)
                U.assert_buckets_equal(o, s, chain.shape[0])
                U.assert_csr_equal(o.query_candidates_dense(Qs, None, 1), cands[r])
        loads[balanced] = max(sizes)
        g_ids = torch.from_numpy(np.stack([p[0] for p in per])).cuda()
        g_sc = torch.from_numpy(np.stack([p[1] for p in per])).cuda()
        m_ids = torch.empty((nq, K), dtype=torch.int32, device="cuda")
        m_sc = torch.empty((nq, K), dtype=torch.float64, device="cuda")
        shards[0].merge_topk_dev(g_ids.data_ptr(), g_sc.data_ptr(), world, nq, K, B.METRIC_DOT, m_ids.data_ptr(), m_sc.data_ptr())
        shards[0].sync()
        torch.cuda.synchronize()
        U.assert_topk_close(f_ids, f_sc, m_ids.cpu().numpy(), m_sc.cpu().numpy())
        for i in range(nq):                                              # union of the ranks' candidate sets
            u = np.unique(np.concatenate([c[1][c[0][i]:c[0][i + 1]] for c in cands]))
            assert np.array_equal(u, f_cand[f_off[i]:f_off[i + 1]])
    ideal = 8000 * chain.shape[0] / world
    assert loads[True] <= max(1.02 * loads[False], 4 / 3 * ideal)           # greedy largest-first: within 4/3 of the optimum


# ---- persist / reload (dpf_save / dpf_load) --------------------------------------------------------------------------
def test_save_load_dense_roundtrip(tmp_path):
    """A reloaded index answers exactly like the one that was saved: same buckets, candidate sets, top-k, store kind."""
    from similaritysearchbyrdf_b200 import DPFIndex
    rng = np.random.default_rng(3)
    d = 64
    X = np.clip(np.rint(40 * rng.standard_normal((5000, d)) + 128), 0, 255) + 0.0      # bytes -> compact store
    A, chain, Ap = U.make_functions(d, family_size=64, table_num=4, permutation_num=2, seed=17)
    ix = U.make_index(d, A, chain, Ap, bucket_overflow=40)
    ix.fit_dense(X[:3000])
    ix.fit_dense(X[3000:])                                                         # saved after an append
    path = tmp_path / "index.dpf"
    ix.save(path)
    ix2 = DPFIndex.load(path)
    assert len(ix2) == 5000 and ix2.stats()["store_kind"] == ix.stats()["store_kind"] == B.STORE_KIND_U8
    for t in range(chain.shape[0]):
        for a, b in zip(ix.dump_buckets(t), ix2.dump_buckets(t)):
            assert np.array_equal(a, b)
    Q = X[::50] + 0.25
    U.assert_csr_equal(ix.query_candidates_dense(Q, None, 1), ix2.query_candidates_dense(Q, None, 1))
    for metric in (B.METRIC_DOT, B.METRIC_ANGULAR, B.METRIC_L2):
        i1, s1 = ix.query_topk_dense(Q, None, 1, 10, metric)
        i2, s2 = ix2.query_topk_dense(Q, None, 1, 10, metric)
        assert np.array_equal(i1, i2) and np.array_equal(s1, s2)
    ix2.fit_dense(X[:100])                                                         # a reloaded index can grow
    assert len(ix2) == 5100


def test_save_load_sparse_roundtrip(tmp_path):
    from similaritysearchbyrdf_b200 import DPFIndex
    from tests.test_gpu_parity import _small_csr
    D = 400
    A, chain, Ap = U.make_functions(D, family_size=100, table_num=3, permutation_num=2, seed=5)
    indptr, idx, val = _small_csr(3000, D, 6)
    ix = U.make_index(D, A, chain, Ap, bucket_overflow=20)
    ix.fit_csr(indptr, idx, val)
    path = tmp_path / "sparse.dpf"
    ix.save(path)
    ix2 = DPFIndex.load(path)
    qids = np.arange(0, 3000, 37, dtype=np.int32)
    U.assert_csr_equal(ix.query_candidates_by_id(qids, 1), ix2.query_candidates_by_id(qids, 1))
    qp, qi, qv = _small_csr(32, D, 7)
    U.assert_csr_equal(ix.query_candidates_csr(qp, qi, qv, None, 0), ix2.query_candidates_csr(qp, qi, qv, None, 0))


def test_load_rejects_garbage(tmp_path):
    from similaritysearchbyrdf_b200 import DPFIndex
    p = tmp_path / "junk.dpf"
    p.write_bytes(b"not an index")
    with pytest.raises(B.DpfError):
        DPFIndex.load(p)


def test_save_load_keeps_the_shard_the_store_mode_and_the_removed_ids(tmp_path):
    """A file holds the shard of the rank that saved it (ownership mask — also a balanced one —, store mode, removed ids):
    two ranks save, two fresh handles load, and the reloaded shards answer exactly like the saved ones."""
    from similaritysearchbyrdf_b200 import DPFIndex
    X, Q = synth.config2(n=20_000, nq=100, d=128)
    A, chain = synth.angle_family(128, 128, 10, 3, 32, 88389)
    Ap = synth.partitioner_family(30, 3, 88390)
    gone = np.arange(100, 3000, 7, dtype=np.int32)
    shards, loaded = [], []
    for r in range(2):
        s = U.make_index(128, A, chain, Ap, bucket_overflow=60, rank=r, world=2, store_mode=B.STORE_F64_ONLY)
        s.set_balanced_partition(True)
        s.fit_dense(X)
        s.remove(gone)
        s.save(tmp_path / f"shard{r}.dpf")
        shards.append(s)
        loaded.append(DPFIndex.load(tmp_path / f"shard{r}.dpf"))
    owned = np.stack([s.owned_subindexes() for s in loaded])
    assert (owned.sum(axis=0) == 1).all() and np.array_equal(owned, np.stack([s.owned_subindexes() for s in shards]))
    for s, l in zip(shards, loaded):
        assert l.stats()["store_kind"] == s.stats()["store_kind"] == B.STORE_KIND_F64          # the mode came back, not AUTO
        for t in (0, 11, 29):
            for a, b in zip(s.dump_buckets(t), l.dump_buckets(t)):
                assert np.array_equal(a, b)
        i1, s1 = s.query_topk_dense(Q, None, 1, 10, B.METRIC_DOT)
        i2, s2 = l.query_topk_dense(Q, None, 1, 10, B.METRIC_DOT)
        assert np.array_equal(i1, i2) and np.array_equal(s1, s2, equal_nan=True)
        assert not np.isin(i2, gone).any()
