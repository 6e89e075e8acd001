"""Committed LUT-mode fixtures (tests/golden/lut_suite_golden.npz, made by tests/golden/make_lut_goldens.py from
the CPU oracle): the oracle must keep reproducing them (CPU), and the CUDA path must match them within the
north_star tolerances with identical LFC/EL existence (GPU)."""

import os

import numpy as np
import pytest

from oracle import parcel as op

HERE = os.path.dirname(os.path.abspath(__file__))
FIELDS = ["cape", "cin", "lcl_pressure", "lcl_temperature", "lcl_virtual_temperature", "lfc_pressure",
          "lfc_temperature", "el_pressure", "el_temperature"]


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "lut_suite_golden.npz"))


@pytest.mark.parametrize("name", ["era5", "model70"])
def test_oracle_reproduces_lut_goldens(golden, oracle_tables, name):
    p, t, td = [golden[f"{name}_{k}"] for k in ("pressure", "temperature", "dewpoint")]
    P = p.astype(np.float64)
    T, D = t.astype(np.float64), td.astype(np.float64)
    P2 = np.broadcast_to(P[:, None], T.shape) if P.ndim == 1 else P
    opts = op.Options(op.MoistLapseLUT(oracle_tables), lcl_mode="converged", metpy_compat="1.4.1")
    res = op.suite(P2, T, D, opts)
    for kind in ("sb", "ml", "mu"):
        for f in FIELDS:
            a, b = res[f"{kind}_{f}"], golden[f"{name}_{kind}_{f}"]
            assert np.array_equal(np.isnan(a), np.isnan(b)), (name, kind, f)
            ok = ~np.isnan(b)
            assert np.allclose(a[ok], b[ok], rtol=1e-11, atol=1e-9), (name, kind, f)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["era5", "model70"])
def test_cuda_path_matches_lut_goldens(golden, name):
    import torch
    from xarray_parcel_b200 import _lib
    ctx = _lib.get_context(0)
    if not ctx.tables_loaded():
        ctx.tables_build()
    p, t, td = [torch.from_numpy(golden[f"{name}_{k}"]).cuda() for k in ("pressure", "temperature", "dewpoint")]
    res = ctx.cape_cin(p, t, td, kinds=("sb", "ml", "mu"))
    for kind in ("sb", "ml", "mu"):
        for f in FIELDS:
            a = res[kind][f].double().cpu().numpy()
            b = golden[f"{name}_{kind}_{f}"]
            assert np.array_equal(np.isnan(a), np.isnan(b)), (name, kind, f)       # LFC/EL existence: bit-exact
            ok = ~np.isnan(b)
            if f in ("cape", "cin"):
                assert np.all(np.abs(a[ok] - b[ok]) <= np.maximum(1.0, 1e-3 * np.abs(b[ok]))), (name, kind, f)   # 0.1 % or 1 J/kg
            else:
                assert np.all(np.abs(a[ok] - b[ok]) <= 1e-3 * np.abs(b[ok])), (name, kind, f)                   # 1e-3 relative
