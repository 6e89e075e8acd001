"""Tie-rich soundings: the exact column code (the device functions of xp_column.cuh / xp_parcels.cuh, host-compiled in
tests/hostsim) against the whole-array oracle on columns drawn from a COARSE LATTICE -- pressures multiples of 20-50
hPa, temperatures and dewpoints multiples of 0.5 K, NaN temperatures / dewpoints anywhere, trailing NaN pressures,
1 to 20 levels.  On such columns the events that a column-serial sweep and the reference's masked whole-array passes
must agree on BY RULE rather than by rounding are common: a level exactly at the mixed-layer top (PF:85-90) or at the
most-unstable bound, two levels equidistant from the bound (PF:224-226), equal neighbouring values (np.sign == 0 in
find_intersections, PF:1019-1022), parcels saturated at their own level, columns that end inside the layer.

Left out of the comparison of LFC / EL / CAPE / CIN (and counted): parcels saturated to the last bit (T == Td), where
the reference's answer hinges on whether ITS exp(log(p)) rounds below p (knife-edge set; e.g.
np.exp(np.log(725.0)) = 724.9999999999999 but libm's is 725.0)."""

import numpy as np
import pytest

import hostsim_util as hs
from oracle import parcel as op

FIELDS = ["cape", "cin", "lcl_pressure", "lcl_temperature", "lcl_virtual_temperature",
          "lfc_pressure", "lfc_temperature", "el_pressure", "el_temperature"]
LCL_FIELDS = FIELDS[2:5]


def lattice_columns(n, L, seed, dp):
    r = np.random.default_rng(seed)
    p0 = 1000.0 - dp * r.integers(0, 3, n)
    steps = r.integers(1, 3, (L, n)); steps[0] = 0
    p = p0[None, :] - dp * np.cumsum(steps, axis=0)
    t0 = 270 + 0.5 * r.integers(0, 70, n)
    lapse = 0.5 * r.integers(-2, 9, (L, n)); lapse[0] = 0
    t = t0[None, :] - np.cumsum(lapse, axis=0)
    td = t - 0.5 * r.integers(0, 30, (L, n))
    t = np.where(r.random((L, n)) < 0.01, np.nan, t)
    td = np.where(r.random((L, n)) < 0.01, np.nan, td)
    cut = r.integers(2, L + 8, n)                                   # trailing NaN pressures in some columns
    p = np.where(np.arange(L)[:, None] >= cut[None, :], np.nan, p)
    p = np.where(np.maximum.accumulate(np.isnan(p) | (p < 60), axis=0), np.nan, p)
    return p, t, td


@pytest.mark.parametrize("L,dp", [(12, 25.0), (9, 40.0), (20, 20.0), (4, 50.0), (2, 25.0), (1, 25.0)])
@pytest.mark.parametrize("o", [dict(vtc=True, li="log", pn=True, c="1.4.1"), dict(vtc=False, li="linear", pn=False, c="1.6.2")],
                         ids=["reference-defaults", "metpy-mode-1.6.2"])
def test_lattice_soundings_match_oracle(oracle_tables, L, dp, o):
    n = 1500
    p, t, td = lattice_columns(n, L, 100 + L, dp)
    opts = op.Options(op.MoistLapseLUT(oracle_tables), lcl_mode="converged", metpy_compat=o["c"])
    ora = op.suite(p, t, td, opts, virtual_temperature_correction=o["vtc"], lcl_interp=o["li"],
                   pos_cape_neg_cin=o["pn"])
    excused = 0
    for kind in ("sb", "ml", "mu"):
        res = hs.cape_cin(p, t, td, oracle_tables, kind=kind, vtc=o["vtc"], lcl_interp=o["li"],
                          pos_cape_neg_cin=o["pn"], metpy_compat=141 if o["c"] == "1.4.1" else 162)
        assert res["flags"] == 0
        if kind == "sb":
            knife = t[0] == td[0]
        else:
            knife = res["parcel_temperature"] == res["parcel_dewpoint"]
            for f in ("pressure", "temperature", "dewpoint"):
                a, b = res["parcel_" + f], ora[f"{kind}_parcel_{f}"]
                assert np.array_equal(np.isnan(a), np.isnan(b)), (kind, f)
                assert np.allclose(a[~np.isnan(b)], b[~np.isnan(b)], rtol=1e-12, atol=0), (kind, f)
        excused += int(knife.sum())
        for f in FIELDS:
            a, b = res[f], ora[kind + "_" + f]
            use = np.ones(n, bool) if f in LCL_FIELDS else ~knife
            assert np.array_equal(np.isnan(a[use]), np.isnan(b[use])), f"{kind} {f}: NaN pattern differs"
            ok = use & ~np.isnan(b)
            err = np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1.0)
            assert err.size == 0 or err.max() < 1e-9, f"{kind} {f}: max rel err {err.max():.3e}"
    assert excused < 0.08 * 3 * n


@pytest.mark.gpu
@pytest.mark.parametrize("L,dp", [(12, 25.0), (9, 40.0), (20, 20.0), (4, 50.0), (3, 25.0)])
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_lattice_soundings_on_the_gpu(L, dp, dtype):
    """The same columns through the C ABI: float32 takes the fast path (per-column pressure; its ties and margins must
    end on the fix-up list), float64 the exact kernel.  Lattice values are exact in float32, so both see the oracle's
    inputs.  Knife-edge columns (see above) are excused from LFC / EL / CAPE / CIN."""
    import torch
    from test_gpu_parity import FAST_BOUND, _lib, otab
    ctx = _lib.get_context(0)
    ctx.tables_build()
    idx, cur = ctx.tables_get()
    pl, tl = otab.default_grids()
    tables = otab.AdiabatTables(pl, tl, idx, cur)
    n = 4000
    p, t, td = lattice_columns(n, L, 300 + L, dp)
    opts = op.Options(op.MoistLapseLUT(tables), lcl_mode="converged")
    ora = op.suite(p, t, td, opts)
    tdt = getattr(torch, dtype)
    P, T, D = [torch.tensor(x, dtype=tdt, device="cuda") for x in (p, t, td)]
    res = ctx.cape_cin(P, T, D, kinds=("sb", "ml", "mu"))
    assert (ctx.last_exact_count() >= 0) == (dtype == "float32")
    excused = 0
    for kind in ("sb", "ml", "mu"):
        r = {f: res[kind][f].double().cpu().numpy() for f in FIELDS}
        if kind == "sb":
            knife = t[0] == td[0]
        else:
            knife = ora[f"{kind}_parcel_temperature"] == ora[f"{kind}_parcel_dewpoint"]
        excused += int(knife.sum())
        for f in FIELDS:
            a, b = r[f], ora[kind + "_" + f]
            use = np.ones(n, bool) if f in LCL_FIELDS else ~knife
            assert np.array_equal(np.isnan(a[use]), np.isnan(b[use])), f"{kind} {f}: NaN pattern differs"
            ok = use & ~np.isnan(b)
            diff = np.abs(a[ok] - b[ok])
            if f in ("cape", "cin"):
                bad = (diff > 1.0) & (diff > 1e-3 * np.abs(b[ok]))          # north_star: 0.1 % or 1 J/kg
            else:
                bad = diff > 1e-3 * np.abs(b[ok])
            assert not bad.any(), f"{kind} {f}: {int(bad.sum())} columns outside the tolerance"
            rel, ab = FAST_BOUND[f] if dtype == "float32" else (1e-9, 1e-9)    # the fast path's own error bound
            over = diff > rel * np.abs(b[ok]) + ab
            assert not over.any(), f"{kind} {f}: {int(over.sum())} columns outside the {dtype} bound, max {diff.max():.3e}"
    assert excused < 0.08 * 3 * n


@pytest.mark.gpu
@pytest.mark.parametrize("L,dp", [(37, 25.0), (16, 50.0), (5, 100.0)])
def test_lattice_soundings_on_a_shared_axis(L, dp):
    """Tie-rich columns on ONE shared pressure axis (the headline kernel: `suite_fast_kernel`, v7 sweep): temperatures
    and dewpoints on the 0.5 K lattice, NaN values anywhere, against the oracle within the fast path's error bound."""
    import torch
    from test_gpu_parity import FAST_BOUND, _lib, otab
    ctx = _lib.get_context(0)
    ctx.tables_build()
    idx, cur = ctx.tables_get()
    pl, tl = otab.default_grids()
    tables = otab.AdiabatTables(pl, tl, idx, cur)
    n = 6000
    _, t, td = lattice_columns(n, L, 500 + L, dp)
    axis = 1000.0 - dp * np.arange(L)
    axis = axis[axis >= 60.0]
    t, td = t[:axis.size], td[:axis.size]
    p = np.broadcast_to(axis[:, None], t.shape)
    ora = op.suite(p, t, td, op.Options(op.MoistLapseLUT(tables), lcl_mode="converged"))
    A, T, D = [torch.tensor(np.ascontiguousarray(x), dtype=torch.float32, device="cuda") for x in (axis, t, td)]
    res = ctx.cape_cin(A, T, D, kinds=("sb", "ml", "mu"))
    assert ctx.last_exact_count() >= 0                                  # the float32 path ran
    for kind in ("sb", "ml", "mu"):
        knife = (t[0] == td[0]) if kind == "sb" else (ora[f"{kind}_parcel_temperature"] == ora[f"{kind}_parcel_dewpoint"])
        for f in FIELDS:
            a, b = res[kind][f].double().cpu().numpy(), ora[kind + "_" + f]
            use = np.ones(n, bool) if f in LCL_FIELDS else ~knife
            assert np.array_equal(np.isnan(a[use]), np.isnan(b[use])), f"{kind} {f}: NaN pattern differs"
            ok = use & ~np.isnan(b)
            rel, ab = FAST_BOUND[f]
            over = np.abs(a[ok] - b[ok]) > rel * np.abs(b[ok]) + ab
            assert not over.any(), f"{kind} {f}: {int(over.sum())} columns outside the float32 bound"


@pytest.mark.parametrize("kind", ["sb", "ml", "mu"])
@pytest.mark.parametrize("L,dp", [(12, 25.0), (5, 50.0), (2, 25.0)])
def test_lattice_profile_rows_match_oracle(oracle_tables, kind, L, dp):
    """parcel_profile_with_lcl rows (PF:806-931) on the lattice: where the LCL row goes when the LCL coincides with a
    level (saturated parcels), which levels are dropped below the lifted parcel, NaN-pressure levels."""
    p, t, td = lattice_columns(700, L, 900 + L, dp)
    opts = op.Options(op.MoistLapseLUT(oracle_tables), lcl_mode="converged")
    fn = {"sb": op.surface_based_cape_cin, "ml": op.mixed_layer_cape_cin, "mu": op.most_unstable_cape_cin}[kind]
    prof = fn(p, t, td, opts)[1]
    res = hs.cape_cin(p, t, td, oracle_tables, kind=kind, profile=True)
    n = prof["pressure"].shape[0]           # the oracle trims with dropna(how='all') like PF:1552/1637
    for k in hs.PROFILE:
        a, b = res["profile"][k][:n], prof[k]
        assert np.array_equal(np.isnan(a), np.isnan(b)), f"{k}: NaN pattern differs"
        ok = ~np.isnan(b)
        assert np.allclose(a[ok], b[ok], rtol=1e-11, atol=0), k
        assert np.isnan(res["profile"][k][n:]).all()
