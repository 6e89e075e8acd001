import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _gpu_unavailable_reason():
    lib = os.path.join(ROOT, "xarray_parcel_b200", "libxparcel.so")
    if not os.path.exists(lib):
        return "xarray_parcel_b200/libxparcel.so is not built (python -c 'import __graft_entry__ as g; g.build()')"
    try:
        import torch
        if not torch.cuda.is_available():
            return "no CUDA device"
    except Exception as e:  # pragma: no cover
        return f"torch unavailable: {e}"
    return None


def pytest_collection_modifyitems(config, items):
    """Tests marked ``gpu`` are SKIPPED on a machine without a CUDA device or without the built library, so a
    CPU-only run of ``pytest tests`` stays green; XP_REQUIRE_GPU=1 (set it on the B200 box) turns the skip
    into a failure -- there a missing GPU or library must not pass silently."""
    reason = _gpu_unavailable_reason()
    if reason is None:
        return
    require = os.environ.get("XP_REQUIRE_GPU", "") not in ("", "0")
    gpu_items = [item for item in items if "gpu" in item.keywords]
    if require and gpu_items:
        raise pytest.UsageError(f"XP_REQUIRE_GPU=1 but {reason} ({len(gpu_items)} gpu tests selected)")
    for item in gpu_items:
        item.add_marker(pytest.mark.skip(reason=reason))


@pytest.fixture(scope="session")
def soundings():
    """Sounding vectors of the reference's known-answer tests (tests/golden/ut_soundings.json,
    produced by tests/golden/extract_ut_soundings.py from modules/unit_tests.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "ut_soundings.json")) as f:
        raw = json.load(f)

    def conv(v):
        if isinstance(v, list):
            return np.array([np.nan if x is None else x for x in v], dtype=np.float64)
        return np.nan if v is None else float(v)

    return {fn: {k: conv(v) for k, v in d.items()} for fn, d in raw.items()}


@pytest.fixture(scope="session")
def oracle_tables():
    from oracle import tables
    return tables.load_tables()
