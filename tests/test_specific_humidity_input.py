"""Specific humidity consumed directly (the q -> Td front end fused into the kernels' load stage).

The reference's callers compute the dewpoint first -- ``dat['dewpoint'] = metpy.calc.dewpoint_from_specific_humidity(
pressure, temperature, specific_humidity)`` (PF:1889, 1969; parcel_test.py:432-436) -- and lift afterwards.  With
``xp_columns.dewpoint_is_specific_humidity`` every kernel converts as a level is loaded, in the MetPy form of
``metpy_compat`` (1.4.1 through the relative humidity, 1.6.2 through the vapour pressure).  Parity: the oracle lifts
the float64 dewpoint computed from the SAME (p, T, q)."""

import numpy as np
import pytest

import hostsim_util as hs
from oracle import parcel as op
from oracle import thermo as th
from xarray_parcel_b200 import synth

FIELDS = ["cape", "cin", "lcl_pressure", "lcl_temperature", "lfc_pressure", "el_pressure", "parcel_dewpoint"]


def _q_columns(kind, n, seed):
    """Synthetic columns with the moisture expressed as specific humidity (float32, as model output holds it)."""
    if kind == "era5":
        p, t, td = synth.era5_columns(n, seed=seed, saturated=0.0)
        P = np.broadcast_to(p.numpy().astype(np.float64)[:, None], t.shape)
    else:
        p, t, td = synth.model_level_columns(n, 70, seed=seed, saturated=0.0, nan_columns=0.0, allnan_columns=0.0)
        P = p.numpy().astype(np.float64)
    T, D = t.numpy().astype(np.float64), td.numpy().astype(np.float64)
    e = th.saturation_vapor_pressure(D)
    w = th.mixing_ratio_from_pressures(e, P)
    q = (w / (1 + w)).astype(np.float32)
    return p.numpy(), t.numpy(), q, P, T


def _oracle(P, T, q, tables, compat):
    D = th.dewpoint_from_specific_humidity(P, T, q.astype(np.float64), metpy_compat=compat)
    return op.suite(P, T, D, op.Options(op.MoistLapseLUT(tables), metpy_compat=compat)), D


def _compare(res, redo, o, n, what, min_kept):
    kept = 0
    for q_, kind in enumerate(("sb", "ml", "mu")):
        keep = (redo & (1 << q_)) == 0
        if kind == "mu":
            keep &= (redo & 8) == 0
        kept += keep.sum()
        for f in FIELDS:
            a = res[kind][f][keep].astype(np.float64)
            if f == "parcel_dewpoint":
                if kind == "sb":
                    continue
                b = o[f"{kind}_parcel_dewpoint"][keep]
            else:
                b = o[f"{kind}_{f}"][keep]
            assert np.array_equal(np.isnan(a), np.isnan(b)), (what, kind, f)
            ok = ~np.isnan(b)
            if f in ("cape", "cin"):
                assert np.all(np.abs(a[ok] - b[ok]) <= np.maximum(1.0, 1e-3 * np.abs(b[ok]))), (what, kind, f)
            else:
                assert np.all(np.abs(a[ok] - b[ok]) <= 1e-3 * np.abs(b[ok])), (what, kind, f)
    assert kept >= min_kept * 3 * n, (what, kept)


@pytest.mark.parametrize("layout,compat", [("era5", "1.4.1"), ("model", "1.4.1"), ("model", "1.6.2"), ("model-table", "1.4.1")])
def test_fast_path_converts_specific_humidity_on_load_hostsim(oracle_tables, layout, compat):
    """The SAME per-column code the kernels inline (tests/hostsim), on the CPU ("model-table": the per-column-pressure
    sweep on the shared-memory adiabat table, suite_fast_ptab_kernel<K, true>)."""
    n = 3000
    table = layout == "model-table"
    layout = "model" if table else layout
    p, t, q, P, T = _q_columns(layout, n, seed=31)
    o, _ = _oracle(P, T, q, oracle_tables, compat)
    code = 141 if compat == "1.4.1" else 162
    hs.set_qmode(code)
    hs.set_pcol_table(table)
    try:
        out = hs.fast_suite(p, t, q, oracle_tables, metpy_compat=code)
    finally:
        hs.set_qmode(0)
        hs.set_pcol_table(False)
    assert out is not None
    res, redo = out
    _compare(res, redo, o, n, (layout, compat), 0.9)


def test_exact_path_converts_specific_humidity_on_load_hostsim(oracle_tables):
    n = 400
    p, t, q, P, T = _q_columns("model", n, seed=33)
    for compat, code in (("1.4.1", 141), ("1.6.2", 162)):
        o, _ = _oracle(P, T, q, oracle_tables, compat)
        hs.set_qmode(code)
        try:
            r = hs.cape_cin(P, T, q.astype(np.float64), oracle_tables, kind="mu", metpy_compat=code)
        finally:
            hs.set_qmode(0)
        for f in ("cape", "cin", "lcl_pressure", "lfc_pressure", "el_pressure"):
            a, b = r[f], o[f"mu_{f}"]
            assert np.array_equal(np.isnan(a), np.isnan(b)), (compat, f)
            ok = ~np.isnan(b)
            assert np.allclose(a[ok], b[ok], rtol=1e-9, atol=1e-7), (compat, f)


@pytest.mark.gpu
@pytest.mark.parametrize("layout,compat,dtype", [("era5", "1.4.1", "f32"), ("model", "1.4.1", "f32"),
                                                 ("model", "1.6.2", "f32"), ("model", "1.4.1", "f64"),
                                                 ("era5", "1.6.2", "f32")])
def test_suite_with_specific_humidity_matches_oracle(oracle_tables, layout, compat, dtype):
    """Through the C ABI: one pass over (p, T, q), against the oracle lifting the float64 dewpoint of the same q;
    and against the two-pass route (xp_dewpoint_from_specific_humidity, then the suite) on the kept columns."""
    import torch
    from xarray_parcel_b200 import _lib
    ctx = _lib.get_context(0)
    if not ctx.tables_loaded():
        ctx.tables_build()
    n = 20000
    p, t, q, P, T = _q_columns(layout, n, seed=37)
    o, D = _oracle(P, T, q, oracle_tables, compat)
    code = 141 if compat == "1.4.1" else 162
    opts = _lib.make_options(metpy_compat=compat)
    tt = torch.float64 if dtype == "f64" else torch.float32
    dp, dt, dq = [torch.from_numpy(np.ascontiguousarray(a)).to(tt).cuda() for a in (p, t, q)]
    res = ctx.cape_cin(dp, dt, dq, kinds=("sb", "ml", "mu"), options=opts, specific_humidity=True)
    torch.cuda.synchronize()
    for kind in ("sb", "ml", "mu"):
        for f in ("cape", "cin", "lcl_pressure", "lcl_temperature", "lfc_pressure", "el_pressure"):
            a, b = res[kind][f].double().cpu().numpy(), o[f"{kind}_{f}"]
            assert np.array_equal(np.isnan(a), np.isnan(b)), (layout, compat, kind, f)          # LFC/EL existence
            ok = ~np.isnan(b)
            if f in ("cape", "cin"):
                assert np.all(np.abs(a[ok] - b[ok]) <= np.maximum(1.0, 1e-3 * np.abs(b[ok]))), (kind, f)
            else:
                assert np.all(np.abs(a[ok] - b[ok]) <= 1e-3 * np.abs(b[ok])), (kind, f)
        if kind != "sb":
            a, b = res[kind]["parcel_dewpoint"].double().cpu().numpy(), o[f"{kind}_parcel_dewpoint"]
            assert np.allclose(a, b, rtol=1e-6, equal_nan=True), kind
    # the level index of the most-unstable parcel is an integer output: bit-exact (PF:102-135 on the converted dewpoint)
    mu = op.most_unstable_parcel({"pressure": P, "temperature": T, "dewpoint": D}, depth=300)
    k_mu = (P > mu["pressure"][None, :]).sum(0)
    assert np.array_equal(res["mu"]["level_shift"].cpu().numpy(), k_mu)
