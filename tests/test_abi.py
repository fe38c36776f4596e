"""The C-ABI shared library loads and exports every symbol include/baofit_b200.h declares, and
the product path fails loudly without a GPU (no CPU fallback). Runs without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "baofit_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built):
    lib = ctypes.CDLL(os.path.join(ROOT, "baofit_b200", "libbaofit_b200.so"))
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export " + s


def test_capi_binds_every_declared_symbol(built):
    from baofit_b200 import capi
    assert sorted(capi.EXPORTS) == declared_symbols()


def test_no_cpu_fallback(built):
    import torch
    from baofit_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.BaofitError) as e:
        capi.Context(0)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """Nothing under baofit_b200/ may import, link or call the oracle."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "baofit_b200")):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "liboracle" not in src and "orc_" not in src, f
