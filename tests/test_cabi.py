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


def test_msssim_workspace_geometry_is_host_only(built):
    """icadv_ssim_vg_workspace_floats is pure host arithmetic (no GPU): two floats per (plane, strip of 118 output
    columns, row segment); segments never shorter than 32 rows, enough blocks for ~2 waves when the image allows it; the
    level launcher and the coefficient launcher derive their block counts from this same number."""
    L = built.lib()
    f = L.icadv_ssim_vg_workspace_floats
    # 32 images x 3 planes, 512 x 768, valid window: 7 strips, 2 segments of 256 rows
    assert f(96, 512, 768, 0) == 96 * 7 * 2 * 2
    # one image: the row axis is split as far as 32-row segments allow (16 of them for 512 rows)
    assert f(3, 512, 768, 0) == 3 * 7 * 16 * 2
    # zero "same" padding adds 5 pixels a side: 778 columns still need 7 strips, 522 rows -> 17 segments at most
    assert f(3, 512, 768, 1) == 3 * 7 * 17 * 2
    # coarsest level of the pyramid (32 x 48): one strip, one segment
    assert f(96, 32, 48, 0) == 96 * 1 * 1 * 2
    # smaller than the 11-tap window, or no planes: nothing to allocate
    assert f(3, 10, 48, 0) == 0 and f(0, 512, 768, 0) == 0
    for planes, h, w in ((1, 161, 161), (192, 177, 203), (7, 1024, 2048)):
        n = f(planes, h, w, 0)
        assert n > 0 and n % (2 * planes) == 0
