"""CPU suite: libcolq.so loads without a GPU and exports exactly the symbols include/colq.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "colq.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(colq_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_surface():
    names = declared_symbols()
    for must in ("colq_create", "colq_register", "colq_col_i32", "colq_col_str", "colq_associate_fk", "colq_associate_csr",
                 "colq_query_create", "colq_query_child", "colq_query_criteria_i32_range", "colq_query_criteria_str",
                 "colq_execute", "colq_comm_init"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from colq import _ffi
    lib = _ffi.load()  # raises if lib/libcolq.so is missing
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} is declared in include/colq.h but not exported by libcolq.so"
    assert set(_ffi.SIGNATURES) == set(declared_symbols()), "colq/_ffi.py and include/colq.h disagree"
    assert lib.colq_abi_version() == _ffi.ABI_VERSION


def test_no_cpu_fallback_without_gpu():
    """Without a usable sm_100 device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from colq.engine import ColqContext, ColqError
    with pytest.raises(ColqError):
        ColqContext(0)


def test_product_package_never_imports_the_oracle():
    pkg = ROOT / "java-columnar-query-engine_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")):
        text = path.read_text()
        assert not re.search(r"oracle_system|liboracle|#include\s*[\"<][^\n]*oracle|import\s+oracle|from\s+oracle", text), path
