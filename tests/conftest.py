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
