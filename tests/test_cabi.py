"""The C-ABI shared library builds for sm_100a, loads without a GPU, and exports every symbol that
include/icadv.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    from imagecompression_adversarial_b200 import _lib
    return _lib


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "icadv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(icadv_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(built):
    lib = ctypes.CDLL(built.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/icadv.h but not exported"


def test_python_binding_covers_header(built):
    assert set(declared_symbols()) == set(built.exported_symbols())


def test_version_and_error_string(built):
    L = built.lib()
    assert L.icadv_version() >= 100
    assert isinstance(L.icadv_last_error(), bytes)


def test_no_silent_fallback_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from imagecompression_adversarial_b200 import ops
    with pytest.raises(built.IcadvError):
        ops.require_device()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "imagecompression_adversarial_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
