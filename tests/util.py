"""Shared helpers: build the CPU oracle and the CUDA index with identical configuration and hash functions."""
import numpy as np

from oracle import oracle_py as O
from similaritysearchbyrdf_b200 import synth


def make_functions(d, family_size=100, table_num=10, permutation_num=3, k=32, pb=3, seed=88387):
    A, chain = synth.angle_family(d, family_size, table_num, permutation_num, k, seed)
    Ap = synth.partitioner_family(chain.shape[0], pb, seed + 1)
    return A, chain, Ap


def make_oracle(d, A, chain, Ap, **kw):
    L, k = chain.shape
    pb = Ap.shape[1]
    b, w = kw.pop("b", None), kw.pop("w", None)
    pb_b, pb_w = kw.pop("pb_b", None), kw.pop("pb_w", None)
    o = O.Oracle(d=d, L=L, k=k, P=A.shape[0], pb=pb, **kw)
    o.set_family(A, chain, b, w)
    o.set_partitioners(Ap, pb_b, pb_w)
    return o


def make_index(d, A, chain, Ap, **kw):
    from similaritysearchbyrdf_b200 import DPFIndex
    L, k = chain.shape
    pb = Ap.shape[1]
    b, w = kw.pop("b", None), kw.pop("w", None)
    pb_b, pb_w = kw.pop("pb_b", None), kw.pop("pb_w", None)
    store_mode = kw.pop("store_mode", None)
    ix = DPFIndex(d=d, L=L, k=k, pb=pb, **kw)
    if store_mode is not None:
        ix.set_store_mode(store_mode)
    ix.set_family(A, chain, b, w)
    ix.set_partitioners(Ap, pb_b, pb_w)
    return ix


def csr_sets(off, ids):
    return [ids[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def assert_csr_equal(a, b):
    (off_a, ids_a), (off_b, ids_b) = a, b
    assert np.array_equal(off_a, off_b), "candidate counts differ"
    assert np.array_equal(ids_a, ids_b), "candidate ids differ"


def assert_buckets_equal(o, ix, L):
    for t in range(L):
        d0, o0, i0 = o.dump_buckets(t)
        d1, o1, i1 = ix.dump_buckets(t)
        assert np.array_equal(d0, d1), f"table {t}: bucket descriptors differ"
        assert np.array_equal(o0, o1), f"table {t}: bucket sizes differ"
        assert np.array_equal(i0, i1), f"table {t}: bucket membership differs"


def assert_topk_close(ids_o, sc_o, ids_g, sc_g, rtol=1e-12):
    """ids exact wherever neighbouring scores are distinct at rtol; scores within rtol (north-star tolerance)."""
    assert ids_o.shape == ids_g.shape
    both_nan = np.isnan(sc_o) & np.isnan(sc_g)
    scale = np.maximum(np.abs(sc_o), 1e-300)
    assert np.all(both_nan | (np.abs(sc_o - sc_g) <= rtol * scale)), "scores differ beyond 1e-12 relative"
    mism = ids_o != ids_g
    if mism.any():
        # a mismatch is only admissible inside a run of scores tied at rtol
        for q, r in zip(*np.nonzero(mism)):
            s = sc_o[q, r]
            tied = np.abs(sc_o[q] - s) <= 4 * rtol * max(abs(s), 1e-300)
            assert tied.sum() > 1, f"query {q} rank {r}: ids differ without a score tie"
            assert set(ids_o[q][tied]) == set(ids_g[q][tied]) or tied[-1], f"query {q}: tie group differs"
