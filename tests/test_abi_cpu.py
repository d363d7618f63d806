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
                assert "oracle" not in src.lower().replace("oracle/", "oracle/") or f == "__init__.py" and False, \
                    f"{f} mentions the oracle"
