"""The drop-in boundary with xarray objects (north_star: "accepts and returns the same xarray objects").

xarray is not installable in the build image, so ``tests/xr_double.py`` stands in for it (the subset of the API the
host layer uses, xarray semantics) and is registered as ``sys.modules['xarray']``: the ``is_xr`` branch of
``xarray_parcel_b200.parcel_functions`` -- dims/coords handling, 1-D pressure coordinate, non-leading vertical
dimension, re-labelled vertical coordinate of L+1-level profiles (PF:968-970), variable names and attrs
(PF:669-677, 1188-1196, 1366-1385, 1453-1473), prefix renames (PF:1508-1512) -- then runs for real.
CPU tests cover the (un)wrapping; GPU tests drive the public functions and compare with the oracle.
"""

import numpy as np
import pytest

import xr_double as xd
from oracle import parcel as op
from xarray_parcel_b200 import synth

VD = "model_level_number"


@pytest.fixture()
def pf():
    mod = xd.install()
    try:
        yield mod
    finally:
        xd.uninstall()


def _grid(nt=2, ny=3, nx=4, L=30, seed=5, lead="lev"):
    """(p, T, Td) DataArrays on dims (time, lev, y, x) ['lev' second] or (lev, time, y, x), labels 1..L (UT:142-152)."""
    n = nt * ny * nx
    p, t, td = synth.model_level_columns(n, L, seed=seed, nan_columns=0.0, allnan_columns=0.0)
    shape = (L, nt, ny, nx)
    dims = (VD, "time", "latitude", "longitude")
    coords = {VD: np.arange(1, L + 1), "time": np.arange(nt), "latitude": -30.0 + np.arange(ny),
              "longitude": 140.0 + np.arange(nx)}
    out = []
    for a, name in ((p, "pressure"), (t, "temperature"), (td, "dewpoint")):
        da = xd.DataArray(a.numpy().astype(np.float64).reshape(shape), dims=dims, coords=coords, name=name)
        if lead != "lev":
            da = da.transpose("time", VD, "latitude", "longitude")
        out.append(da)
    return out


# ------------------------------------------------------------------------------------------------ CPU: (un)wrapping
def test_layout_unwraps_non_leading_vertical_dim(pf):
    import torch
    p, t, td = _grid(lead="time")
    lay = pf._Layout(t, VD, 0)
    assert lay.is_xr and lay.vert_axis == 1 and lay.L == 30 and lay.col_shape == (2, 3, 4)
    blk = lay.to_block(t, torch.device("cpu"), torch.float64)
    assert tuple(blk.shape) == (30, 24)
    assert np.array_equal(blk.numpy(), np.moveaxis(t.values, 1, 0).reshape(30, 24))
    # a DataArray with the dims in another order is aligned by NAME, not by position
    blk2 = lay.to_block(td.transpose("latitude", "longitude", VD, "time"), torch.device("cpu"), torch.float64)
    assert np.array_equal(blk2.numpy(), np.moveaxis(td.values, 1, 0).reshape(30, 24))


def test_layout_keeps_a_1d_pressure_coordinate_shared(pf):
    import torch
    _, t, _ = _grid()
    p1 = xd.DataArray(np.linspace(1000.0, 100.0, 30), dims=(VD,), coords={VD: np.arange(1, 31)})
    lay = pf._Layout(t, VD, 0)
    blk = lay.to_block(p1, torch.device("cpu"), torch.float64)
    assert blk.dim() == 1 and blk.shape[0] == 30          # passed to the kernels as ONE shared axis (pressure_is_1d)


def test_layout_wraps_scalars_and_relabelled_profiles(pf):
    import torch
    _, t, _ = _grid(lead="time")
    lay = pf._Layout(t, VD, 0)
    s = lay.wrap_scalar(torch.arange(24.0), "lcl_pressure")
    assert s.dims == ("time", "latitude", "longitude") and s.shape == (2, 3, 4)
    assert set(s.coords) == {"time", "latitude", "longitude"}
    assert s.attrs == {"long_name": "Lifting condensation level pressure", "units": "hPa"}      # PF:669-671
    prof = lay.wrap_profile(torch.zeros(31, 24), "virtual_temperature", 31)
    assert prof.dims == t.dims and prof.shape == (2, 31, 3, 4)
    # PF:968-970: the vertical coordinate of the L+1-level profile is re-labelled from the first input label
    assert np.array_equal(prof[VD].values, np.arange(1, 32))
    ex = lay.scalar_to_block(t.isel({VD: 0}), torch.float64)
    assert np.array_equal(ex.numpy(), t.values[:, 0].reshape(-1))


# ------------------------------------------------------------------------------------------------ GPU: public API
def _oracle(p, t, td, lead):
    ax = 0 if lead == "lev" else 1
    P, T, D = [np.moveaxis(a.values, ax, 0).reshape(a.shape[ax], -1) for a in (p, t, td)]
    return P, T, D


def _close(a, b, rel, what):
    a, b = np.asarray(a, dtype=np.float64).reshape(-1), np.asarray(b, dtype=np.float64).reshape(-1)
    assert np.array_equal(np.isnan(a), np.isnan(b)), what
    ok = ~np.isnan(b)
    assert np.all(np.abs(a[ok] - b[ok]) <= rel * np.maximum(np.abs(b[ok]), 1.0)), what


@pytest.mark.gpu
@pytest.mark.parametrize("lead", ["lev", "time"])
def test_surface_based_cape_cin_returns_reference_shaped_datasets(pf, oracle_tables, lead):
    pf.load_moist_adiabat_lookups()
    p, t, td = _grid(lead=lead)
    res, profile = pf.surface_based_cape_cin(p, t, td, vert_dim=VD, prefix="surface")
    # PF:1508-1512 prefix rename, PF:1506-1507 descriptions, PF:1366-1385 units, PF:1472-1473 correction attr
    assert isinstance(res, xd.Dataset) and set(res.keys()) == {"surface_cape", "surface_cin"}
    assert res["surface_cape"].attrs["description"] == "CAPE for surface-based parcel."
    assert res["surface_cin"].attrs["description"] == "CIN for surface-based parcel."
    assert res["surface_cape"].attrs["units"] == "J kg$^{-1}$"
    assert res.attrs["correction"] == "Virtual temperature correction used in CAPE/CIN calculations."
    cols = ("time", "latitude", "longitude")
    assert res["surface_cape"].dims == cols and res["surface_cape"].shape == (2, 3, 4)
    assert np.array_equal(res["surface_cape"]["latitude"].values, t["latitude"].values)
    # the profile: the six level variables on L+1 re-labelled levels, LCL / LFC / EL scalars (PF:806-856, 1475)
    want = {"pressure", "temperature", "virtual_temperature", "environment_temperature",
            "environment_virtual_temperature", "environment_dewpoint", "lcl_pressure", "lcl_temperature",
            "lcl_virtual_temperature", "lfc_pressure", "lfc_temperature", "el_pressure", "el_temperature"}
    assert set(profile.keys()) == want
    assert profile["pressure"].dims == t.dims
    ax = t.dims.index(VD)
    assert profile["pressure"].shape[ax] == 31
    assert np.array_equal(profile["pressure"][VD].values, np.arange(1, 32))          # PF:968-970
    assert profile["lfc_pressure"].attrs == {"long_name": "Level of free convection pressure", "units": "hPa"}
    assert profile["el_temperature"].attrs == {"long_name": "Equilibrium level temperature", "units": "K"}
    # values: the oracle on the same columns
    P, T, D = _oracle(p, t, td, lead)
    o_cc, o_prof = op.surface_based_cape_cin(P, T, D, op.Options(op.MoistLapseLUT(oracle_tables)))
    _close(res["surface_cape"].values, o_cc["cape"], 1e-7, "cape")
    _close(res["surface_cin"].values, o_cc["cin"], 1e-7, "cin")
    _close(profile["lcl_pressure"].values, o_prof["lcl_pressure"], 1e-10, "lcl_pressure")
    got = np.moveaxis(profile["virtual_temperature"].values, ax, 0).reshape(31, -1)
    _close(got, o_prof["virtual_temperature"], 1e-10, "profile virtual_temperature")


@pytest.mark.gpu
def test_mixed_layer_and_most_unstable_wrappers(pf, oracle_tables):
    pf.load_moist_adiabat_lookups()
    p, t, td = _grid(lead="time", seed=9)
    opts = op.Options(op.MoistLapseLUT(oracle_tables))
    P, T, D = _oracle(p, t, td, "time")
    res, profile, mp = pf.mixed_layer_cape_cin(p, t, td, vert_dim=VD, depth=100, prefix="mixed_100")
    assert set(res.keys()) == {"mixed_100_cape", "mixed_100_cin"}
    assert res["mixed_100_cape"].attrs["description"] == "CAPE for fully-mixed lowest 100 hPa parcel."   # PF:1689-1692
    assert {"pressure", "temperature", "dewpoint"} <= set(mp.keys())
    o_cc, _, o_mp = op.mixed_layer_cape_cin(P, T, D, opts, depth=100)
    _close(res["mixed_100_cape"].values, o_cc["cape"], 1e-7, "ml cape")
    _close(mp["temperature"].values, o_mp["temperature"], 1e-10, "mixed parcel temperature")
    res, profile, ul = pf.most_unstable_cape_cin(p, t, td, vert_dim=VD, depth=300, prefix="max")
    assert set(res.keys()) == {"max_cape", "max_cin"}
    assert res["max_cape"].attrs["description"] == "CAPE for most-unstable parcel in lowest 300 hPa."     # PF:1594-1597
    o_cc, o_prof, _ = op.most_unstable_cape_cin(P, T, D, opts, depth=300)
    _close(res["max_cape"].values, o_cc["cape"], 1e-7, "mu cape")
    _close(res["max_cin"].values, o_cc["cin"], 1e-7, "mu cin")
    # dropna(how='all') (PF:1552): the returned profile is trimmed to the longest lifted column + the LCL row
    n_lev = profile["pressure"].shape[profile["pressure"].dims.index(VD)]
    assert n_lev == o_prof["pressure"].shape[0]
    assert np.array_equal(profile["pressure"][VD].values, np.arange(1, n_lev + 1))


@pytest.mark.gpu
def test_parcel_profile_with_lcl_takes_parcel_dataarrays(pf, oracle_tables):
    pf.load_moist_adiabat_lookups()
    p, t, td = _grid(seed=11)
    # the reference's own call pattern (PF:1502-1504): parcel = level 0 of each field, as DataArrays without VD
    prof = pf.parcel_profile_with_lcl(p, t, td, parcel_pressure=p.isel({VD: 0}), parcel_temperature=t.isel({VD: 0}),
                                      parcel_dewpoint=td.isel({VD: 0}), vert_dim=VD)
    assert "lfc_pressure" not in prof.keys() and "lcl_pressure" in prof.keys()            # PF:806-856
    P, T, D = _oracle(p, t, td, "lev")
    o = op.parcel_profile_with_lcl(P, T, D, P[0], T[0], D[0], op.Options(op.MoistLapseLUT(oracle_tables)))
    for k in ("pressure", "temperature", "environment_dewpoint"):
        _close(prof[k].values.reshape(31, -1), o[k], 1e-10, k)


@pytest.mark.gpu
def test_shared_1d_pressure_coordinate_and_suite(pf, oracle_tables):
    """ERA5 style: pressure is a 1-D DataArray on the vertical dimension; float32 fields -> the fast kernels."""
    pf.load_moist_adiabat_lookups()
    ny, nx = 6, 8
    p1, t, td = synth.era5_columns(ny * nx, seed=21)
    L = p1.shape[0]
    coords = {"level": np.asarray(p1.numpy(), dtype=np.float64), "latitude": np.arange(ny) * 0.25,
              "longitude": np.arange(nx) * 0.25}
    T = xd.DataArray(t.numpy().reshape(L, ny, nx), dims=("level", "latitude", "longitude"), coords=coords)
    D = xd.DataArray(td.numpy().reshape(L, ny, nx), dims=("level", "latitude", "longitude"), coords=coords)
    P = xd.DataArray(p1.numpy(), dims=("level",), coords={"level": coords["level"]})
    ds = pf.parcel_suite(P, T, D, vert_dim="level")
    want = {f"{pre}_{v}" for pre in ("surface", "mixed_100", "max")
            for v in ("cape", "cin", "lcl_pressure", "lcl_temperature", "lcl_virtual_temperature", "lfc_pressure",
                      "lfc_temperature", "el_pressure", "el_temperature")}
    want |= {f"{pre}_parcel_{v}" for pre in ("mixed_100", "max") for v in ("pressure", "temperature", "dewpoint")}
    assert set(ds.keys()) == want
    assert ds["max_cape"].dims == ("latitude", "longitude") and ds["max_cape"].dtype == np.float32
    P2 = np.broadcast_to(p1.numpy().astype(np.float64)[:, None], (L, ny * nx))
    o = op.suite(P2, t.numpy().astype(np.float64), td.numpy().astype(np.float64),
                 op.Options(op.MoistLapseLUT(oracle_tables)))
    for kind, pre in (("sb", "surface"), ("ml", "mixed_100"), ("mu", "max")):
        a, b = ds[f"{pre}_cape"].values.reshape(-1).astype(np.float64), o[f"{kind}_cape"]
        assert np.all(np.abs(a - b) <= np.maximum(1.0, 1e-3 * np.abs(b))), kind            # 0.1 % or 1 J/kg
        _close(ds[f"{pre}_lcl_pressure"].values, o[f"{kind}_lcl_pressure"], 1e-3, kind)
        a, b = ds[f"{pre}_lfc_pressure"].values.reshape(-1), o[f"{kind}_lfc_pressure"]
        assert np.array_equal(np.isnan(a), np.isnan(b)), kind
