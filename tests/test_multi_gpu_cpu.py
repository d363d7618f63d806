"""N>1 protocol on CPU: two gloo ranks, each holding one shard (sub-indexes p % 2 == rank) of the oracle forest,
all-gather their top-k and merge; the result must equal the unsharded forest's top-k.  This is the host-side
logic of bench.py's multi-GPU step (ownership rule, all-gather layout, merge semantics) without a GPU."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from similaritysearchbyrdf_b200 import synth
    from tests import util as U
    from tests.merge_ref import merge_topk_host
    X, Q = synth.config1(n=6000)
    A, chain, Ap = U.make_functions(100)
    K, steps = 10, 1
    Qs = np.concatenate([Q, X[:40] + 0.01])
    shard = U.make_oracle(100, A, chain, Ap, bucket_overflow=40, rank=rank, world=world)
    shard.fit_dense(X, nthreads=2)
    ids, sc = shard.query_topk_dense(Qs, None, steps, K, 0, nthreads=2)
    t_ids, t_sc = torch.from_numpy(ids), torch.from_numpy(sc)
    g_ids = [torch.empty_like(t_ids) for _ in range(world)]
    g_sc = [torch.empty_like(t_sc) for _ in range(world)]
    dist.all_gather(g_ids, t_ids)
    dist.all_gather(g_sc, t_sc)
    m_ids, m_sc = merge_topk_host(torch.stack(g_ids).numpy(), torch.stack(g_sc).numpy())
    # candidate sets: the union over shards must be the unsharded set
    off, cand = shard.query_candidates_dense(Qs, None, steps)
    counts = torch.from_numpy(np.diff(off))
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    if rank == 0:
        full = U.make_oracle(100, A, chain, Ap, bucket_overflow=40)
        full.fit_dense(X, nthreads=2)
        f_ids, f_sc = full.query_topk_dense(Qs, None, steps, K, 0, nthreads=2)
        f_off, _ = full.query_candidates_dense(Qs, None, steps)
        ok = np.array_equal(f_ids, m_ids) and np.allclose(f_sc, m_sc, rtol=0, atol=0, equal_nan=True)
        # shards partition each table's sub-indexes, so shard candidate counts add up to >= the unsharded count
        tot = sum(c.numpy() for c in all_counts)
        ok = ok and bool(np.all(tot >= np.diff(f_off)))
        open(os.path.join(out_dir, "result"), "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_merge_to_the_unsharded_result(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert open(tmp_path / "result").read() == "ok"


def test_merge_host_dedups_and_orders():
    from tests.merge_ref import merge_topk_host
    gids = np.array([[[5, 3, -1]], [[3, 9, 1]]], np.int32)
    gsc = np.array([[[9.0, 7.0, np.nan]], [[7.0, 6.0, 2.0]]])
    i, s = merge_topk_host(gids, gsc)
    assert i.tolist() == [[5, 3, 9]] and s.tolist() == [[9.0, 7.0, 6.0]]


def test_oracle_explicit_ownership_partitions_the_forest():
    """dpfo_set_owned (the checker of the library's balanced assignment): shards with complementary sub-index sets hold
    disjoint parts of every table and their candidate sets unite to the unsharded ones."""
    from similaritysearchbyrdf_b200 import synth
    from tests import util as U
    X, Q = synth.config1(n=3000)
    A, chain, Ap = U.make_functions(100)
    full = U.make_oracle(100, A, chain, Ap, bucket_overflow=40)
    full.fit_dense(X)
    Qs = np.concatenate([Q, X[:40] + 0.01])
    f_off, f_cand = full.query_candidates_dense(Qs, None, 1)
    masks = [np.array([1, 0, 0, 1, 1, 0, 0, 0], np.uint8), np.array([0, 1, 0, 0, 0, 0, 1, 1], np.uint8),
             np.array([0, 0, 1, 0, 0, 1, 0, 0], np.uint8)]
    _check_shards_unite(X, Qs, A, chain, Ap, masks, f_off, f_cand)
    # the same per (table, sub-index) cell (dpfo_set_owned_cells): a different owner per table
    owner = np.random.default_rng(5).integers(0, 3, (chain.shape[0], 8))
    _check_shards_unite(X, Qs, A, chain, Ap, [(owner == g).astype(np.uint8) for g in range(3)], f_off, f_cand)


def _check_shards_unite(X, Qs, A, chain, Ap, masks, f_off, f_cand):
    from tests import util as U
    cands, sizes = [], []
    for m in masks:
        o = U.make_oracle(100, A, chain, Ap, bucket_overflow=40)
        o.set_owned(m)
        o.fit_dense(X)
        cands.append(o.query_candidates_dense(Qs, None, 1))
        sizes.append(sum(len(o.dump_buckets(t)[2]) for t in range(chain.shape[0])))
    assert sum(sizes) == 3000 * chain.shape[0]
    for i in range(len(Qs)):
        u = np.unique(np.concatenate([c[1][c[0][i]:c[0][i + 1]] for c in cands]))
        assert np.array_equal(u, f_cand[f_off[i]:f_off[i + 1]])
