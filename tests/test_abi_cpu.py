"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/dpf.h
declares; without a GPU it fails loudly (no CPU fallback)."""
import os
import re

import pytest

from similaritysearchbyrdf_b200 import _lib as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dpf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dpf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = B.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dpf.h but not exported"
    assert sorted(B.EXPORTS) == declared


def test_python_constants_match_the_header_enums():
    """The ctypes side spells the header's enums by value: debug options, stage timers and store kinds must not drift
    (a renumbered DPF_DBG_* would silently switch another kernel variant in the tests)."""
    text = open(os.path.join(ROOT, "include", "dpf.h")).read()
    enums = {m.group(1): int(m.group(2)) for m in re.finditer(r"\b(DPF_[A-Z0-9_]+)\s*=\s*(\d+)\s*[,}/]", text)}
    dbg = {k[len("DPF_DBG_"):]: v for k, v in enums.items() if k.startswith("DPF_DBG_") and k != "DPF_DBG_COUNT"}
    assert len(dbg) >= 14
    for name, value in dbg.items():
        assert getattr(B, "DBG_" + name) == value, name
    assert all(v < enums["DPF_DBG_COUNT"] for v in dbg.values())
    for name in ("METRIC_DOT", "METRIC_ANGULAR", "METRIC_L2", "STORE_KIND_U8", "STORE_KIND_F32", "PROBE_DENSE"):
        if "DPF_" + name in enums:
            assert getattr(B, name) == enums["DPF_" + name], name


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from similaritysearchbyrdf_b200 import DPFIndex
    with pytest.raises(B.DpfError) as e:
        DPFIndex(d=8, L=2)
    assert e.value.code == B.ERR_CUDA


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "similaritysearchbyrdf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b|dpf_oracle|oracle_py|dpfo_", src, flags=re.M), \
                    f"{f} reaches into oracle/"


def test_jni_shim_compiles_against_stub_header():
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "jni"), "-s", "check"])


def test_deploy_parsers_match_reference_formats():
    """Vector.scala:162-219 formats (VectorSuite.scala:9-53 resources: sparsevectorfile / densevectorfile)."""
    from similaritysearchbyrdf_b200 import deploy as D
    vid, size, idx, val = D.Vectors.fromString("(3,3,[0,1,2],[1.0,2.0,3.0])")
    assert (vid, size, idx.tolist(), val.tolist()) == (3, 3, [0, 1, 2], [1.0, 2.0, 3.0])
    assert str(D.SparseVector(vid, size, idx, val)) == "(3,3,[0,1,2],[1.0,2.0,3.0])"
    vid, size, idx, val = D.Vectors.fromPythonString("[1, 3, [1, 2], [1.0, 2.5]]")
    assert (vid, size, idx.tolist(), val.tolist()) == (1, 3, [1, 2], [1.0, 2.5])
    vid, vals = D.Vectors.parseDense("[1,[0.1, 0.2,0.4,0.9]]")
    assert vid == 1 and vals.tolist() == [0.1, 0.2, 0.4, 0.9]
    assert D.Vectors.fromStringDense("0.3,0.2,0.9").tolist() == [0.3, 0.2, 0.9]
    assert D.Vectors.analysisKNN("[1, 30, 19, 230]", 3).tolist() == [1, 30, 19]
    with pytest.raises(ValueError):
        D.Vectors.fromString("(3,3,[0,1,2])")
    conf = D.Config.parseString("mclab.lsh.tableNum = 4").withFallback(D.testBaseConf)
    assert conf.getInt("mclab.lsh.tableNum") == 4 and conf.getInt("mclab.lsh.permutationNum") == 3
    lsh = D.LSH(conf)
    assert len(lsh.tableIndexGenerators) == 12 and lsh.chain.shape == (12, 32)   # LSHSuite.scala:24-59 (count only)
