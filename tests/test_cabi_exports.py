"""The C-ABI boundary without a GPU: libxparcel.so builds for sm_100a (nvcc cross-compiles), loads, and exports
every entry point that include/xparcel.h declares; the Python binding lists the same symbols; the product
package refuses to compute without a CUDA device instead of falling back to the CPU."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "xparcel.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)              # comments mention functions too
    return sorted(set(re.findall(r"\b(xp_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from xarray_parcel_b200 import _build
    return ctypes.CDLL(_build.build())


def test_header_declares_the_documented_entry_points():
    names = _declared()
    for must in ("xp_create", "xp_destroy", "xp_tables_build", "xp_cape_cin", "xp_suite", "xp_lcl", "xp_moist_lapse",
                 "xp_parcel_profile", "xp_lfc_el", "xp_cape_cin_base", "xp_interp_levels", "xp_level_crossing",
                 "xp_wet_bulb_temperature", "xp_dewpoint_from_specific_humidity", "xp_storm_proxies",
                 "xp_significant_hail_parameter", "xp_last_error"):
        assert must in names, must


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, f"include/xparcel.h declares {missing} but libxparcel.so does not export them"


def test_binding_lists_every_declared_symbol():
    from xarray_parcel_b200._lib import EXPORTS
    assert sorted(EXPORTS) == _declared()


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from xarray_parcel_b200 import _lib
    with pytest.raises(_lib.XparcelError):
        _lib.Context(0)
