"""The multi-GPU plane inside the library (comm.cu): NCCL communicator per handle, sharded hashing with one key
all-gather, per-rank top k with one all-gather + merge.  On one GPU the world-1 communicator exercises every call; with
two or more GPUs visible, one handle per device in ONE process (threads) must reproduce the unsharded index — which also
covers per-device kernel attributes (two handles on different devices in one process)."""
import threading

import numpy as np
import pytest
import torch  # noqa: F401  (first: the library binds the NCCL that is already in the process)

from similaritysearchbyrdf_b200 import DPFIndex, synth
from similaritysearchbyrdf_b200 import _lib as B
from tests import util as U

pytestmark = pytest.mark.gpu


def _data():
    X, Q = synth.config2(n=40_000, nq=300, d=128)
    A, chain = synth.angle_family(128, 128, 10, 3, 32, 88389)
    Ap = synth.partitioner_family(30, 3, 88390)
    return X, Q, A, chain, Ap


def test_world1_communicator_sharded_fit_and_collective_query():
    X, Q, A, chain, Ap = _data()
    ref = U.make_index(128, A, chain, Ap, bucket_overflow=100)
    ref.fit_dense(X)
    ix = U.make_index(128, A, chain, Ap, bucket_overflow=100)
    ix.comm_init(DPFIndex.comm_unique_id())
    ix.fit_dense_sharded(X[:25_000])
    ix.fit_dense_sharded(X[25_000:])                      # append through the sharded path
    for t in (0, 7, 29):
        for a, b in zip(ref.dump_buckets(t), ix.dump_buckets(t)):
            assert np.array_equal(a, b)
    for metric in (B.METRIC_DOT, B.METRIC_ANGULAR, B.METRIC_L2):
        r = ref.query_topk_dense(Q, None, 1, 10, metric)
        g = ix.query_topk_dense_all(Q, None, 1, 10, metric)
        assert np.array_equal(r[0], g[0]) and np.array_equal(r[1], g[1], equal_nan=True), metric
    assert ix.stage_times_ms is not None
    ix.comm_destroy()
    ix.close(); ref.close()


def test_one_handle_per_device_in_one_process_equals_the_unsharded_index():
    import torch
    G = min(torch.cuda.device_count(), 8)
    if G < 2:
        pytest.skip("needs at least two GPUs")
    X, Q, A, chain, Ap = _data()
    Qr = Q + 0.25                                          # real-valued queries as well
    ref = U.make_index(128, A, chain, Ap, bucket_overflow=100)
    ref.fit_dense(X)
    want = [ref.query_topk_dense(Q, None, 0, 10, B.METRIC_DOT), ref.query_topk_dense(Qr, None, 1, 10, B.METRIC_ANGULAR)]
    uid = DPFIndex.comm_unique_id()
    out, err = [None] * G, []

    def rank(r):
        try:
            ix = U.make_index(128, A, chain, Ap, bucket_overflow=100, device=r, rank=r, world=G)
            ix.set_balanced_partition(True)
            ix.comm_init(uid)
            ix.fit_dense_sharded(X)
            out[r] = [ix.query_topk_dense_all(Q, None, 0, 10, B.METRIC_DOT), ix.query_topk_dense_all(Qr, None, 1, 10, B.METRIC_ANGULAR),
                      ix.owned_subindexes()]
            ix.comm_destroy()
            ix.close()
        except Exception as e:                             # noqa: BLE001
            err.append((r, e))

    ts = [threading.Thread(target=rank, args=(r,)) for r in range(G)]
    [t.start() for t in ts]
    [t.join(300) for t in ts]
    assert not err, err
    owned = np.stack([o[2] for o in out])
    assert (owned.sum(axis=0) == 1).all(), "every sub-index has exactly one owner"
    for r in range(G):
        assert np.array_equal(out[r][0][0], want[0][0]) and np.array_equal(out[r][0][1], want[0][1]), r   # integer data: exact
        U.assert_topk_close(want[1][0], want[1][1], out[r][1][0], out[r][1][1])
