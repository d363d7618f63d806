"""Seeded synthetic inputs of the BASELINE.json configurations and hash-function generation.

Hash functions are *inputs* to the native library (the reference draws them from an unseeded RNG,
AngleHashFamily.scala:29, so they are not reproducible from config; quirk Q7).  `angle_family` is the host-side
analogue of AngleHashFamily.initHashFamily + pick (AngleHashFamily.scala:37-64, 121-149): U[0,1) magnitudes with a
random sign, L2-normalised; chains drawn with replacement from the family; permuted tables are shuffles of the
base chain.
"""
import numpy as np


def angle_family(d, family_size, table_num, permutation_num, k, seed, by_pulling=True):
    """Returns (A[P x d], chain_idx[L x k]) with L = table_num * permutation_num."""
    rng = np.random.default_rng(seed)

    def unit_vectors(m):
        v = rng.random((m, d))
        v[v == 0.0] = 0.5                                   # functions must be fully dense (SimilarityCalculator.scala:43)
        sign = rng.integers(0, 2, (m, d)) > 0
        v = np.where(sign, v, -v)
        return v / np.sqrt((v * v).sum(1, keepdims=True))

    if by_pulling:
        A = unit_vectors(family_size)
        base = rng.integers(0, family_size, (table_num, k))
    else:
        A = unit_vectors(table_num * k)
        base = np.arange(table_num * k).reshape(table_num, k)
    chain = np.empty((table_num * permutation_num, k), np.int32)
    for t in range(table_num):
        for p in range(permutation_num):
            chain[permutation_num * t + p] = rng.permutation(base[t])   # rd.shuffle(hashFunctionChain)
    # keep only the functions that are used, renumbered densely
    used, inv = np.unique(chain, return_inverse=True)
    return np.ascontiguousarray(A[used]), inv.reshape(chain.shape).astype(np.int32)


def partitioner_family(L, pb, seed):
    """L private partitioners of pb 32-d functions each (DensevectorRDFInit.scala:63-77)."""
    rng = np.random.default_rng(seed)
    v = rng.random((L, pb, 32))
    v[v == 0.0] = 0.5
    v = np.where(rng.integers(0, 2, v.shape) > 0, v, -v)
    return v / np.sqrt((v * v).sum(-1, keepdims=True))


def pstable_partitioner_family(L, pb, mu, sigma, w, seed):
    """The partitioners of a pStable index: each table's private LSH(confForPartitioner) draws a PStableHashFamily over
    32 dimensions and picks a chain of pb functions from it (DensevectorRDFInit.scala:63-77, PStableHashFamily.scala:
    37-78).  Returns (Ap[L x pb x 32], b[L x pb], w[L x pb])."""
    rng = np.random.default_rng(seed)
    Ap = mu + sigma * rng.standard_normal((L, pb, 32))
    Ap[Ap == 0.0] = sigma if sigma else 1.0
    return Ap, rng.random((L, pb)) * w, np.full((L, pb), w, np.int32)


def clustered_dense(n, d, seed, centres, spread=0.35, lo=None, hi=None, integer=False, relu=False, chunk=1 << 18):
    """n x d FP64 rows: centre + spread * N(0, I).  With lo/hi the values are mapped to [lo, hi] (and rounded when
    `integer`); with `relu` negative coordinates are clamped to zero first (histogram-like, about half the
    coordinates of a row are zero) so that unrelated rows are far apart in angle, as in real SIFT/GIST data."""
    rng = np.random.default_rng(seed)
    C = rng.standard_normal((centres, d))
    X = np.empty((n, d), np.float64)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        X[s:s + m] = C[rng.integers(0, centres, m)] + spread * rng.standard_normal((m, d))
    if relu:
        np.maximum(X, 0.0, out=X)
        X *= (hi - lo) / 4.0
        X += lo
        np.clip(X, lo, hi, out=X)
    elif lo is not None:
        mn, mx = -4.0, 4.0
        X = np.clip((X - mn) / (mx - mn), 0.0, 1.0) * (hi - lo) + lo
    if integer:
        np.rint(X, out=X)
    return X


def config1(n=20000, d=100):
    """TestSingleRDFSuite 'simple test' shape: 20k x 100-d, 2 queries ~U[0,1)^100 (README.md:32-42)."""
    X = clustered_dense(n, d, 1001, 200)
    Q = np.random.default_rng(1001 + 7).random((2, d))
    return X, Q


def config2(n=1_000_000, nq=10_000, d=128):
    """SIFT shape: non-negative integer-valued doubles in [0, 218], about half of them zero, ~1000 rows per
    cluster; queries are held-out draws from the same mixture."""
    X = clustered_dense(n + nq, d, 1002, max(1, n // 1000), lo=0.0, hi=218.0, integer=True, relu=True)
    return X[:n], X[n:]


def config3(n=1_000_000, nq=1000, d=960):
    """GIST shape: values in [0, 1], ~1000 rows per cluster."""
    X = clustered_dense(n + nq, d, 1003, max(1, n // 1000), lo=0.0, hi=1.0, relu=True)
    return X[:n], X[n:]


def config4_csr(n=2_000_000, D=100_000, mean_nnz=60, seed=1004, chunk=1 << 16):
    """Sparse rows: nnz ~ Poisson(60) clipped >= 1, Zipf(1.1)-distributed distinct indices sorted ascending,
    values ~U(0,1] L2-normalised.  Returns (indptr int64, indices int32, values f64)."""
    rng = np.random.default_rng(seed)
    nnz = np.maximum(rng.poisson(mean_nnz, n), 1).astype(np.int64)
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(nnz, out=indptr[1:])
    indices = np.empty(indptr[-1], np.int32)
    values = np.empty(indptr[-1], np.float64)
    # Zipf over a fixed random relabelling of the features; duplicates inside a row are re-drawn uniformly
    perm = rng.permutation(D)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        for i in range(s, e):
            m = int(nnz[i])
            idx = np.unique(perm[np.minimum(rng.zipf(1.1, m + 8) - 1, D - 1)])
            while len(idx) < m:
                idx = np.unique(np.concatenate([idx, rng.integers(0, D, m - len(idx))]))
            idx = np.sort(rng.choice(idx, m, replace=False)) if len(idx) > m else idx
            v = 1.0 - rng.random(m)
            indices[indptr[i]:indptr[i + 1]] = idx
            values[indptr[i]:indptr[i + 1]] = v / np.sqrt((v * v).sum())
    return indptr, indices, values


def config4_csr_fast(n=2_000_000, D=100_000, mean_nnz=60, seed=1004, chunk=1 << 17):
    """Vectorised generator of the config4 shape for full-size runs (the per-row loop above takes minutes at 2M
    rows): per row, candidate indices are drawn log-uniformly over [0, D) (a Zipf(~1) law over a fixed random
    relabelling of the features), duplicates are dropped and a random subset of min(nnz, distinct) of them is kept,
    sorted ascending; values ~U(0,1], L2-normalised.  Returns (indptr int64, indices int32, values f64)."""
    rng = np.random.default_rng(seed)
    want = np.maximum(rng.poisson(mean_nnz, n), 1).astype(np.int64)
    perm = rng.permutation(D).astype(np.int32)
    width = int(want.max()) + 48
    idx_parts, val_parts, counts = [], [], np.empty(n, np.int64)
    for s0 in range(0, n, chunk):
        e0 = min(n, s0 + chunk)
        m = e0 - s0
        c = perm[np.minimum((np.exp(rng.random((m, width)) * np.log(D)) - 1.0).astype(np.int64), D - 1)]
        c.sort(axis=1)
        dup = np.zeros((m, width), bool)
        dup[:, 1:] = c[:, 1:] == c[:, :-1]
        key = rng.random((m, width))
        key[dup] = 2.0                                        # duplicates are never selected
        distinct = width - dup.sum(1)
        take = np.minimum(want[s0:e0], distinct)
        order = np.argsort(key, axis=1)
        sel = np.arange(width)[None, :] < take[:, None]       # first `take` columns of the random order
        chosen = np.take_along_axis(c, order, axis=1)
        chosen[~sel] = np.iinfo(np.int32).max
        chosen.sort(axis=1)
        idx_parts.append(chosen[sel])                          # row-major: ascending inside every row
        v = 1.0 - rng.random((m, width))
        v[~sel] = 0.0
        v /= np.sqrt((v * v).sum(1, keepdims=True))
        val_parts.append(v[sel])
        counts[s0:e0] = take
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(counts, out=indptr[1:])
    return indptr, np.concatenate(idx_parts).astype(np.int32), np.concatenate(val_parts)
