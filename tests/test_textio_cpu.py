"""Native text loaders (dpf_parse_dense_file / dpf_parse_sparse_file, host code in libdpf_b200.so) against a literal
Python restatement of the reference's line parsers (Vectors.parseDense / fromPythonString, Vector.scala:194-219)."""
import ctypes as C

import numpy as np
import pytest

from similaritysearchbyrdf_b200 import _lib as B
from similaritysearchbyrdf_b200 import deploy


def _write_dense(path, X, ids=None, spaces=False):
    with open(path, "w") as f:
        for i, row in enumerate(X):
            vid = ids[i] if ids is not None else i
            body = (", " if spaces else ",").join(repr(float(v)) for v in row)
            f.write(f"[{vid}, [{body}]]\n" if spaces else f"[{vid},[{body}]]\n")
            if i % 7 == 3:
                f.write("\n")                                   # empty lines are skipped


def test_dense_file_matches_line_parser(tmp_path):
    rng = np.random.default_rng(0)
    X = rng.standard_normal((6000, 17)) * 10.0 ** rng.integers(-8, 8, (6000, 1))
    X[5, 3] = 0.0; X[6, 0] = -0.0; X[7, 1] = 1e-310; X[8, 2] = 123456789012345678.0
    p = tmp_path / "dense.txt"
    _write_dense(p, X, ids=rng.integers(0, 10**6, 6000), spaces=True)
    got = deploy.load_dense_file(p)
    ref = np.stack([deploy.Vectors.parseDense(l)[1] for l in open(p) if l.strip()])
    assert got.shape == (6000, 17)
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64))          # bit for bit, file order, ids ignored
    assert np.array_equal(got.view(np.uint64), X.view(np.uint64))


def test_dense_file_errors(tmp_path):
    lib = B.load()
    n = C.c_int64(0)
    assert lib.dpf_parse_dense_file(str(tmp_path / "missing.txt").encode(), 4, None, 0, C.byref(n)) == B.ERR_INVALID
    p = tmp_path / "ragged.txt"
    p.write_text("[0,[1.0,2.0,3.0]]\n[1,[1.0,2.0]]\n")
    out = np.empty((2, 3))
    assert lib.dpf_parse_dense_file(str(p).encode(), 3, None, 0, C.byref(n)) == B.OK and n.value == 2
    assert lib.dpf_parse_dense_file(str(p).encode(), 3, C.c_void_p(out.ctypes.data), 1, C.byref(n)) == B.ERR_CAPACITY
    assert lib.dpf_parse_dense_file(str(p).encode(), 3, C.c_void_p(out.ctypes.data), 2, C.byref(n)) == B.ERR_INVALID
    with pytest.raises(ValueError):
        deploy.load_dense_file(p)
    e = tmp_path / "empty.txt"
    e.write_text("\n\n")
    assert deploy.load_dense_file(e).size == 0


def test_sparse_file_matches_line_parser(tmp_path):
    rng = np.random.default_rng(1)
    p = tmp_path / "sparse.txt"
    rows = []
    with open(p, "w") as f:
        for i in range(5000):
            m = int(rng.integers(0, 12))
            idx = rng.choice(1000, m, replace=False)               # unsorted in the file
            val = rng.standard_normal(m)
            rows.append((idx, val))
            f.write(f"[{i}, 1000, [{', '.join(str(int(v)) for v in idx)}], [{', '.join(repr(float(v)) for v in val)}]]\n")
    (indptr, idx, val), dim = deploy.load_sparse_file(p)
    assert dim == 1000 and len(indptr) == 5001 and indptr[-1] == sum(len(r[0]) for r in rows)
    for i in (0, 1, 17, 4999):
        _, sz, ri, rv = deploy.Vectors.fromPythonString(open(p).readlines()[i])
        order = np.argsort(ri, kind="stable")
        assert np.array_equal(idx[indptr[i]:indptr[i + 1]], ri[order])
        assert np.array_equal(val[indptr[i]:indptr[i + 1]], rv[order])
    for i, (ri, rv) in enumerate(rows):
        order = np.argsort(ri, kind="stable")
        assert np.array_equal(idx[indptr[i]:indptr[i + 1]], ri[order].astype(np.int32))
        assert np.array_equal(val[indptr[i]:indptr[i + 1]], rv[order])
    bad = tmp_path / "bad.txt"
    bad.write_text("[0, 10, [1, 2, 3], [1.0, 2.0]]\n")
    with pytest.raises(ValueError):
        deploy.load_sparse_file(bad)


def test_dense_parser_number_formats(tmp_path):
    """Java's toDouble and strtod accept the same decimal / scientific spellings; every spelling must give the identical
    double, whatever the spacing around brackets and commas."""
    from hypothesis import given, settings, strategies as st

    fmts = ["{!r}", "{:.17g}", "{:.17e}", "{:+.17e}", "{:.20f}"]

    @settings(max_examples=40, deadline=None)
    @given(st.lists(st.floats(allow_nan=False, allow_infinity=False, width=64), min_size=3, max_size=3), st.integers(0, len(fmts) - 1),
           st.sampled_from(["", " ", "  "]))
    def check(vals, f, sp):
        if fmts[f] == "{:.20f}" and any(v != 0 and not (1e-2 <= abs(v) <= 1e15) for v in vals):
            return                                              # 20 decimals do not round-trip outside this range
        p = tmp_path / "f.txt"
        p.write_text(f"[{sp}7,{sp}[" + f",{sp}".join(fmts[f].format(v) for v in vals) + f"]{sp}]\n")
        got = deploy.load_dense_file(p, d=3)
        assert got.shape == (1, 3)
        assert np.array_equal(got[0].view(np.uint64), np.array(vals, np.float64).view(np.uint64)), (vals, fmts[f])

    check()
