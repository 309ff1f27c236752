"""CPU suite: the C-ABI library builds/loads without a GPU, exports every symbol include/srk.h declares, and
refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from ml_super_resolution_b200 import _ffi
    from ml_super_resolution_b200.build import build_library
    build_library()
    return _ffi.lib()


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "srk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(srk_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from ml_super_resolution_b200 import _ffi
    names = _declared_functions()
    assert len(names) >= 25
    raw = C.CDLL(os.path.join(ROOT, "ml_super_resolution_b200", "libsrk.so"))
    for n in names:
        assert hasattr(raw, n), f"{n} declared in srk.h but not exported"
        assert n in _ffi.SIGNATURES, f"{n} has no ctypes binding"
    assert sorted(_ffi.SIGNATURES) == names, "bindings and header disagree"


def test_version_and_fpa_geometry(lib):
    assert lib.srk_version() >= 100
    rows = lib.srk_fpa_rows(64, 41, 41)
    assert rows % 128 == 0 and rows >= 64 * 42 * 42 + 43


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.srk_create(0, C.byref(h)) != 0
    assert b"no CPU fallback" in lib.srk_last_error()
    from ml_super_resolution_b200 import ops
    with pytest.raises(Exception):
        ops.handle(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ml_super_resolution_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
