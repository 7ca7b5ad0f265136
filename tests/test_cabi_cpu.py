"""CPU: the C-ABI library builds, loads and exports every symbol include/gpzoo_b200.h declares; the product
path refuses CPU tensors loudly (there is no fallback)."""
import os

import pytest
import torch


@pytest.fixture(scope="module")
def lib():
    from gpzoo_b200 import _cabi, build
    if not os.path.exists(_cabi.LIB_PATH):
        build.build()
    return _cabi.lib()


def test_exports_match_header(lib):
    from gpzoo_b200 import _cabi
    syms = _cabi.exported_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.gpz_abi_version() >= 1
    assert b"success" in lib.gpz_error_string(0)


def test_no_cpu_fallback(lib):
    import gpzoo_b200 as gz
    k = gz.kernels.NSF_RBF(L=2)
    with pytest.raises(gz._cabi.GpzError):
        k(torch.randn(5, 2), torch.randn(3, 2))


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "gpzoo_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
