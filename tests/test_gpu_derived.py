"""GPU tests of the derived convective indices against the oracle: the interpolation
and level-crossing kernels behind lifted_index, deep_convective_index, isobar_temperature, lapse_rate,
freezing/melting_level_height and wind_shear, called through the reference-facing Python surface."""

import numpy as np
import pytest
import torch

from oracle import parcel as op
from oracle import tables as otab
from xarray_parcel_b200 import _lib, synth
import xarray_parcel_b200.parcel_functions as parcel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = _lib.get_context(0)
    c.tables_build()
    return c


def _cols(n=4000, L=60, seed=3, **kw):
    p, t, td = synth.model_level_columns(n, L, seed=seed, **kw)
    P, T, D = [x.numpy().astype(np.float64) for x in (p, t, td)]
    H = 7500.0 * np.log(P[0][None, :] / P) + 120.0             # a height field [m]
    return P, T, D, H


def _same(a, b, rtol=1e-12):
    a = np.asarray(a, dtype=np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(b)
    assert np.allclose(a[ok], b[ok], rtol=rtol, atol=0)


def test_interp_kernel_semantics(ctx):
    """linear_interp PF:1758-1811: exact hits, duplicated coordinates (mean), NaNs, no extrapolation."""
    c = np.array([[10., 10.], [8., 8.], [8., 6.], [4., np.nan], [2., 2.]])
    x = np.array([[1., 1.], [3., 3.], [5., np.nan], [7., 7.], [9., 9.]])
    for at in (8.0, 7.0, 10.0, 11.0, 1.0, 3.0):
        ora = op.linear_interp({"x": x}, c, np.full(2, at))["x"]
        (got,) = ctx.interp_levels(torch.from_numpy(c).cuda(), [torch.from_numpy(x).cuda()], at)
        _same(got.cpu().numpy(), ora)
    ora = op.log_interp({"x": x}, c, np.array([5.0, 3.0]))["x"]
    (got,) = ctx.interp_levels(torch.from_numpy(c).cuda(), [torch.from_numpy(x).cuda()],
                               torch.tensor([5.0, 3.0], dtype=torch.float64).cuda(), log=True)
    _same(got.cpu().numpy(), ora)


def test_isobar_lapse_dci_against_oracle(ctx):
    P, T, D, H = _cols(nan_columns=0.05)
    _same(parcel.isobar_temperature(P, T, 500), op.isobar_temperature(P, T, 500))
    _same(parcel.lapse_rate(P, T, H), op.lapse_rate(P, T, H), rtol=1e-9)
    li = np.linspace(-8, 6, P.shape[1])
    got = parcel.deep_convective_index(P, T, D, li, prefix="surface")
    _same(got["surface_dci"], op.deep_convective_index(P, T, D, li))
    # float32 inputs and a shared pressure axis
    p1, t, td = synth.era5_columns(3000, seed=5)
    Pb = np.broadcast_to(p1.numpy().astype(np.float64)[:, None], t.shape)
    got = parcel.isobar_temperature(p1.numpy(), t.numpy(), 600)
    assert got.dtype == np.float32
    _same(got, op.isobar_temperature(Pb, t.numpy().astype(np.float64), 600), rtol=2e-7)


def test_freezing_and_melting_level(ctx):
    P, T, D, H = _cols(seed=4)
    _same(parcel.freezing_level_height(T, H), op.freezing_level_height(T, H))
    mlh, wb = parcel.melting_level_height(P, T, D, H)
    _same(wb, op.wet_bulb_temperature_fast(T, D))
    _same(mlh, op.freezing_level_height(op.wet_bulb_temperature_fast(T, D), H))
    # several crossings: the lowest one is returned (PF:2154 .min)
    h = np.linspace(0, 5000, 11)[:, None]
    t = (273.15 + np.array([2, 1, -1, -2, 1, 3, -1, -4, -6, -8, -9.0]))[:, None]
    assert abs(float(parcel.freezing_level_height(t, h)[0]) - 750.0) < 1e-9
    assert np.isnan(parcel.freezing_level_height(t + 50, h)[0])


def test_wind_shear(ctx):
    rng = np.random.default_rng(1)
    N, L = 3000, 40
    H = np.cumsum(rng.uniform(50, 600, (L, N)), axis=0)
    U, V = rng.normal(5, 8, (L, N)), rng.normal(0, 8, (L, N))
    su, sv = rng.normal(2, 3, N), rng.normal(0, 3, N)
    got = parcel.wind_shear(su, sv, U, V, H, shear_height=6000)
    ora = op.wind_shear(su, sv, U, V, H, 6000)
    for k in ("shear_u", "shear_v", "shear_magnitude"):
        _same(got[k], ora[k])
    assert np.array_equal(np.asarray(got["positive_shear"]), ora["positive_shear"])


def test_lifted_index_from_gpu_profile(ctx, soundings):
    """lifted_index on the profile returned by mixed_layer_cape_cin (as min_conv_properties does,
    PF:1900-1911) against the oracle, and the reference's known answer UT:1353-1386 (-7.9176350,
    exact-ODE; the lookup table moves it by < 0.05 K)."""
    idx, cur = ctx.tables_get()
    pl, tl = otab.default_grids()
    tb = otab.AdiabatTables(pl, tl, idx, cur)
    P, T, D, _ = _cols(n=1500, L=50, seed=6, nan_columns=0, allnan_columns=0)
    cc, prof, mp = parcel.mixed_layer_cape_cin(P, T, D, prefix="mixed_100")
    li = parcel.lifted_index(prof, prefix="mixed_100")["mixed_100_lifted_index"]
    opts = op.Options(op.MoistLapseLUT(tb), lcl_mode="converged")
    _, oprof, _ = op.mixed_layer_cape_cin(P, T, D, opts)
    _same(li, op.lifted_index(oprof), rtol=1e-9)
    s = soundings["test_lifted_index"]
    p, t, td = [np.asarray(s[k], dtype=np.float64)[:, None] for k in ("pressure", "temperature", "dewpoint")]
    prof = parcel.parcel_profile(p, p[0], t[0], td[0])               # as UT:1378-1384: no LCL level
    prof["environment_temperature"] = t
    assert abs(float(parcel.lifted_index(prof)["lifted_index"][0]) + 7.9176350) < 0.05


@pytest.mark.parametrize("compat", ["1.4.1", "1.6.2"])
def test_conv_properties_assembly(ctx, compat):
    """conv_properties / min_conv_properties (PF:1872-2100) end to end from specific humidity, against
    the oracle's assembly of the same steps."""
    idx, cur = ctx.tables_get()
    pl, tl = otab.default_grids()
    tb = otab.AdiabatTables(pl, tl, idx, cur)
    P, T, D, H = _cols(n=1200, L=45, seed=8, nan_columns=0.02, allnan_columns=0)
    es_d = 6.112 * np.exp(17.67 * (D - 273.15) / (D - 29.65))
    w = 0.6219569100577033 * es_d / (P - es_d)
    rng = np.random.default_rng(2)
    dat = {"pressure": P, "temperature": T, "specific_humidity": w / (1 + w), "height_asl": H,
           "wind_u": rng.normal(5, 8, P.shape), "wind_v": rng.normal(0, 8, P.shape),
           "wind_height_above_surface": H - H[0][None, :] + 10.0,
           "surface_wind_u": rng.normal(2, 3, P.shape[1]), "surface_wind_v": rng.normal(0, 3, P.shape[1])}
    opts = op.Options(op.MoistLapseLUT(tb), lcl_mode="converged", metpy_compat=compat)
    for min_set in (False, True):
        ora = op.conv_properties(dict(dat), opts, min_set=min_set)
        got = (parcel.min_conv_properties(dict(dat), metpy_compat=compat) if min_set
               else parcel.conv_properties(dict(dat), metpy_compat=compat))
        assert set(ora) == set(got), set(ora) ^ set(got)
        for k, v in ora.items():
            g = np.asarray(got[k])
            if v.dtype == bool:
                assert np.array_equal(g.astype(bool), v), k
            else:
                assert np.array_equal(np.isnan(g), np.isnan(v)), k
                ok = ~np.isnan(v)
                assert np.allclose(g[ok], v[ok], rtol=1e-8, atol=1e-8), k
        if not min_set:
            # the full chain the reference's users run: conv_properties -> storm_proxies (PF:2323-2407)
            prox, prox_ora = parcel.storm_proxies(got), op.storm_proxies(ora)
            for k in _lib.PROXY_FLAGS:
                assert np.array_equal(np.asarray(prox[k]).astype(bool), np.asarray(prox_ora[k]).astype(bool)), k
            _same(prox["ship"], prox_ora["ship"], rtol=1e-7)


# ---- pointwise kernels ------------------------------------------------------------------
def test_pointwise_thermo_against_oracle(ctx):
    """dry_lapse PF:291-316, mixing_ratio PF:684-710, virtual_temperature PF:782-804, q -> Td (PF:1889,
    1969; both MetPy forms) and the saturation mixing ratio, through the reference-facing functions."""
    from oracle import thermo as th
    P, T, D, _ = _cols(n=3000, L=40, nan_columns=0.05)
    for compat in ("1.4.1", "1.6.2"):
        w = th.mixing_ratio_from_t_td(T, D, P, compat)
        _same(parcel.mixing_ratio(T, D, P, metpy_compat=compat), w, rtol=1e-12)
        q = w / (1 + w)
        _same(parcel.dewpoint_from_specific_humidity(P, T, q, metpy_compat=compat),
              th.dewpoint_from_specific_humidity(P, T, q, compat), rtol=1e-12)
    _same(parcel.virtual_temperature(T, w), th.virtual_temperature(T, w), rtol=1e-15)
    _same(parcel.dry_lapse(P, T[0], P[0]), th.dry_lapse(P, T[0][None, :], P[0][None, :]), rtol=1e-13)
    _same(parcel.dry_lapse(P, T[0]), th.dry_lapse(P, T[0][None, :], np.max(P, axis=0, keepdims=True)), rtol=1e-13)
    _same(ctx.saturation_mixing_ratio(torch.from_numpy(P).cuda(), torch.from_numpy(D).cuda()).cpu().numpy(),
          th.saturation_mixing_ratio(P, D), rtol=1e-13)
    # float32 tensors on the device stay float32 tensors on the device
    r = parcel.mixing_ratio(torch.from_numpy(T).float().cuda(), torch.from_numpy(D).float().cuda(),
                            torch.from_numpy(P).float().cuda())
    assert r.is_cuda and r.dtype == torch.float32
    w141 = th.mixing_ratio_from_t_td(T, D, P, "1.4.1")
    ok = ~np.isnan(w141)
    assert np.allclose(r.cpu().numpy()[ok], w141[ok], rtol=2e-5)    # float32 inputs (T, Td rounded to ~2e-5 K)


def test_wet_bulb_temperature_against_oracle(ctx, ):
    """wet_bulb_temperature PF:389-445 (Normand's rule) = lcl + moist_lapse on the lookup tables per point."""
    idx, cur = ctx.tables_get()
    pl, tl = otab.default_grids()
    opts = op.Options(op.MoistLapseLUT(otab.AdiabatTables(pl, tl, idx, cur)), lcl_mode="converged")
    P, T, D, H = _cols(n=600, L=30, seed=9, nan_columns=0.05)
    ora = op.wet_bulb_temperature(P, T, D, opts)
    got = parcel.wet_bulb_temperature(P, T, D)
    a = np.asarray(got)
    assert np.array_equal(np.isnan(a), np.isnan(ora))
    ok = ~np.isnan(ora)
    # the table cell of an LCL within rounding of a cell edge may differ: one 0.02 K adiabat step at most
    assert np.abs(a[ok] - ora[ok]).max() < 0.05
    assert (np.abs(a[ok] - ora[ok]) > 1e-9).mean() < 1e-3
    # between dewpoint and temperature, up to the table resolution (the nearest 0.5 hPa node is coarse aloft)
    low = ok & (P > 400.0)
    assert (a[low] <= T[low] + 0.1).all() and (a[low] >= D[low] - 0.1).all()
    ml, wb = parcel.melting_level_height(P, T, D, H, fast=False)
    _same(np.asarray(wb), a)
    _same(np.asarray(ml), op.freezing_level_height(a, H), rtol=1e-12)


def _proxy_fields(n, seed):
    rng = np.random.default_rng(seed)
    f = {"mixed_100_cape": rng.uniform(-50, 3000, n), "mixed_50_cape": rng.uniform(-50, 3000, n),
         "mu_cape": rng.uniform(-50, 4000, n), "shear_magnitude": rng.uniform(0, 40, n),
         "mixed_100_lifted_index": rng.uniform(-8, 6, n), "mixed_100_dci": rng.uniform(0, 45, n),
         "positive_shear": (rng.uniform(0, 1, n) > 0.3).astype(np.float64),
         "mixed_50_cin": rng.uniform(-100, 0, n), "mixed_100_cin": rng.uniform(-120, 0, n),
         "lapse_rate_700_500": rng.uniform(-9, -4, n), "mu_mixing_ratio": rng.uniform(0.004, 0.018, n),
         "temp_500": rng.uniform(250, 272, n), "freezing_level": rng.uniform(1000, 5000, n)}
    for k in f:                       # NaNs everywhere, threshold values exactly hit
        f[k][rng.uniform(0, 1, n) < 0.03] = np.nan
    f["shear_magnitude"][:6] = [7.0, 27.0, 5.0, 7.5, 6.999, 27.001]
    f["mu_cape"][6:9] = [1300.0, 1474.0, 1299.0]
    f["freezing_level"][9:11] = [2400.0, 2399.0]
    f["mu_mixing_ratio"][11:15] = [0.011, 0.0136, 0.0109, 0.01361]
    return f


def test_significant_hail_parameter_and_storm_proxies(ctx):
    """significant_hail_parameter PF:2261-2306 and storm_proxies PF:2323-2407: flags bit-exact, SHIP to
    rounding, NaN inputs and values exactly on the thresholds included."""
    f = _proxy_fields(20_000, 4)
    ora = op.storm_proxies(f)
    got = parcel.storm_proxies(f)
    for k in _lib.PROXY_FLAGS:
        assert np.array_equal(np.asarray(got[k]).astype(bool), np.asarray(ora[k]).astype(bool)), k
    _same(got["ship"], ora["ship"], rtol=1e-14)
    ship = parcel.significant_hail_parameter(f["mu_cape"], f["mu_mixing_ratio"], f["lapse_rate_700_500"],
                                             f["temp_500"], f["shear_magnitude"], f["freezing_level"])
    _same(ship, op.significant_hail_parameter(f["mu_cape"], f["mu_mixing_ratio"], f["lapse_rate_700_500"],
                                              f["temp_500"], f["shear_magnitude"], f["freezing_level"]), rtol=1e-14)
    # device tensors in, device tensors out
    dev = {k: torch.from_numpy(v).cuda() for k, v in f.items()}
    got = parcel.storm_proxies(dev)
    assert got["proxy_Kunz2007"].is_cuda and got["proxy_Kunz2007"].dtype == torch.bool
    assert np.array_equal(got["proxy_Allen2014"].cpu().numpy(), ora["proxy_Allen2014"])


def test_wet_bulb_reference_known_answers(ctx):
    """The reference's own wet-bulb known answers (unit_tests.py:79-104) through the CUDA path: the lookup
    tables quantise the LCL to 0.5 hPa x 0.02 K cells, so these hold to ~2 decimals (as the reference's LUT
    mode does for its moist-lapse tests, unit_tests.py:106-112), not to the 5-7 of the exact-ODE mode."""
    wb = parcel.wet_bulb_temperature(np.array([1000.0]), np.array([25 + 273.15]), np.array([15 + 273.15]))
    assert abs(float(wb[0]) - (18.3432116 + 273.15)) < 0.03
    wb = parcel.wet_bulb_temperature(np.array([850.0]), np.array([17.6 + 273.15]), np.array([17.6 + 273.15]))
    assert abs(float(wb[0]) - (17.6 + 273.15)) < 0.03
    wb = parcel.wet_bulb_temperature(np.array([1013.0, 1000.0, 990.0]), np.array([25.0, 20.0, 15.0]) + 273.15,
                                     np.array([20.0, 15.0, 10.0]) + 273.15)
    assert np.abs(np.asarray(wb) - (np.array([21.44487, 16.73673, 12.06554]) + 273.15)).max() < 0.03
