"""Device-side bounds checks in place of compute-sanitizer (closed on the B200 pool, profiles/r2_sanitizer_pool_closed.txt).

``libxparcel_check.so`` (built by ``__graft_entry__.build()`` with -DXP_BOUNDS_CHECK) asserts every computed index of
the fast kernels -- shared-memory stash and coefficient table, the uncertain-column list, the 32-bit T/Td offsets --
and traps on a violation, which surfaces as a CUDA error in the child process below.  The child drives the fast
kernels over the awkward inputs: NaN-laden and all-NaN columns, saturated parcels, 3-level and 56-level axes, axes
reaching above the table top, column counts that are not a multiple of the tile, specific-humidity input."""

import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK_LIB = os.path.join(ROOT, "xarray_parcel_b200", "libxparcel_check.so")

CHILD = r"""
import sys
sys.path.insert(0, %r)
import numpy as np, torch
from xarray_parcel_b200 import _lib, synth
ctx = _lib.get_context(0)
ctx.tables_build()
def run(p, t, td, **kw):
    r = ctx.cape_cin(p.cuda(), t.cuda(), td.cuda(), kinds=("sb", "ml", "mu"), **kw)
    torch.cuda.synchronize()
    return r
n_cases = 0
for n in (1, 31, 640, 641, 100003):
    run(*synth.era5_columns(n, seed=n, nan_columns=0.05, allnan_columns=0.01, saturated=0.1)); n_cases += 1
    run(*synth.model_level_columns(n, 70, seed=n, nan_columns=0.05, allnan_columns=0.01, saturated=0.1)); n_cases += 1
# short / long / odd shared axes
for levels in ([1000, 900, 800], [1000 - 17 * k for k in range(56)], [1050, 1000, 950, 700, 400, 150, 60, 20, 4, 2.6, 2.4, 1.0]):
    L, n = len(levels), 5000
    p = torch.tensor(levels, dtype=torch.float32)
    _, t37, td37 = synth.era5_columns(n, seed=L)
    idx = torch.linspace(0, 36, L).long()
    run(p, t37[idx].contiguous(), td37[idx].contiguous()); n_cases += 1
# specific humidity in place of the dewpoint (converted in the load stage)
p, t, td = synth.era5_columns(50000, seed=5)
e = 6.112 * torch.exp(17.67 * (td - 273.15) / (td - 29.65))
w = 0.622 * e / (p[:, None] - e)
run(p, t, (w / (1 + w)).contiguous(), specific_humidity=True); n_cases += 1
# profile rows from the per-column-pressure kernel
p, t, td = synth.model_level_columns(20000, 90, seed=9)
r = ctx.cape_cin(p.cuda(), t.cuda(), td.cuda(), kinds=("mu",), profile=True); torch.cuda.synchronize(); n_cases += 1
print("BOUNDS_CHECK_OK", n_cases)
"""


@pytest.mark.gpu
def test_fast_kernels_run_clean_with_device_bounds_checks():
    assert os.path.exists(CHECK_LIB), "run __graft_entry__.build() first (it builds the XP_BOUNDS_CHECK variant)"
    env = dict(os.environ, XP_LIB_PATH=CHECK_LIB)
    res = subprocess.run([sys.executable, "-c", CHILD % ROOT], env=env, capture_output=True, text=True, timeout=900)
    tail = (res.stdout + res.stderr)[-3000:]
    assert res.returncode == 0, tail
    assert "BOUNDS_CHECK_OK" in res.stdout and "XP_BOUNDS_CHECK failed" not in tail, tail
