"""GPU tests of the level primitives exposed on their own (xp_levels.cu through the C ABI and the reference-facing
Python surface): insert_level (PF:933-990), shift_out_nans (PF:1699-1720), trapz (PF:164-206), valid_data
(PF:2308-2321) and get_layer (PF:63-100), against the oracle."""

import numpy as np
import pytest
import torch

from oracle import parcel as op
from xarray_parcel_b200 import _lib, synth
import xarray_parcel_b200.parcel_functions as parcel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return _lib.get_context(0)


def _same(a, b, rtol=1e-12):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(b)
    assert np.allclose(a[ok], b[ok], rtol=rtol, atol=0)


def _columns(n=4000, L=50, seed=21):
    p, t, td = synth.model_level_columns(n, L, seed=seed, nan_columns=0.1)
    return [x.numpy().astype(np.float64) for x in (p, t, td)]


def test_insert_level_against_oracle(ctx):
    P, T, D = _columns()
    rng = np.random.default_rng(5)
    N = P.shape[1]
    lev = {"pressure": rng.uniform(50.0, 1100.0, N), "temperature": rng.uniform(200.0, 300.0, N),
           "dewpoint": rng.uniform(190.0, 290.0, N)}
    lev["pressure"][:40] = P[rng.integers(0, P.shape[0], 40), np.arange(40)]
    lev["pressure"][40:50] = np.nan
    P[5:9, 60:80] = np.nan
    d = {"pressure": P, "temperature": T, "dewpoint": D}
    ora = op.insert_level(d, lev, "pressure")
    got = parcel.insert_level(d, lev, "pressure")
    assert list(got.keys()) == list(lev.keys())
    for k in lev:
        _same(got[k], ora[k])
    # float32 I/O keeps the dtype; a reference known answer (UT:1388-1411 pattern): existing level kept below
    d32 = {k: v.astype(np.float32) for k, v in d.items()}
    got32 = parcel.insert_level(d32, lev, "pressure")
    assert got32["temperature"].dtype == np.float32 and got32["temperature"].shape[0] == P.shape[0] + 1
    small = parcel.insert_level({"pressure": np.array([[1000.], [900.], [800.]]), "x": np.array([[1.], [2.], [3.]])},
                                {"pressure": np.array([900.]), "x": np.array([9.])}, "pressure")
    assert small["pressure"][:, 0].tolist() == [1000., 900., 900., 800.] and small["x"][:, 0].tolist() == [1., 2., 9., 3.]
    with pytest.raises(AssertionError, match="fill_value"):
        parcel.insert_level({"pressure": np.array([[1000.], [-999.]])}, {"pressure": np.array([900.])}, "pressure")


def test_shift_out_nans_against_oracle(ctx):
    P, T, D = _columns(seed=22)
    N = P.shape[1]
    k = np.random.default_rng(6).integers(0, 6, N)
    Pn = np.where(np.arange(P.shape[0])[:, None] < k[None, :], np.nan, P)
    Pn[:, :5] = np.nan
    x = {"pressure": Pn, "temperature": T, "dewpoint": D}
    ora = op.shift_out_nans(x, "pressure")
    got = parcel.shift_out_nans(x, "pressure")
    for key in x:
        _same(got[key], ora[key])
    _, shift = ctx.shift_out_nans(torch.from_numpy(Pn).cuda(), [])
    assert np.array_equal(shift.cpu().numpy()[5:], np.where(np.isnan(P[0, 5:]), P.shape[0], k[5:]))


@pytest.mark.parametrize("sign", [0, 1, -1])
def test_trapz_against_oracle(ctx, sign):
    P, T, D = _columns(seed=23)
    V = T - D - 8.0
    mask = np.random.default_rng(7).random(P.shape) < 0.7
    kw = dict(only_positive=sign > 0, only_negative=sign < 0)
    for m in (None, mask):
        ora = op.trapz({"pressure": P, "v": V, "t": T}, "pressure", mask=m, **kw)
        got = parcel.trapz({"pressure": P, "v": V, "t": T}, "pressure", mask=m, **kw)
        for key in ora:
            _same(got[key], ora[key])
    # shared 1-D integration variable, float32
    p1, t, _ = synth.era5_columns(3000, seed=9)
    Pb = np.broadcast_to(p1.numpy().astype(np.float64)[:, None], t.shape)
    (g32,) = ctx.trapz(p1.cuda(), [t.cuda()], sign=sign)
    _same(g32.cpu().numpy(), op.trapz({"pressure": Pb, "t": t.numpy().astype(np.float64)}, "pressure", **kw)["t"], rtol=2e-7)


def test_valid_data(ctx):
    P, _, _ = _columns(seed=24)
    P = P[:, ~np.isnan(P[0])]
    assert parcel.valid_data({"pressure": P}) is True
    Q = P.copy()
    Q[[3, 4], 17] = Q[[4, 3], 17]
    with pytest.raises(AssertionError, match="Pressures must decrease"):
        parcel.valid_data({"pressure": Q})
    Q = P.copy()
    Q[7, 3] = Q[6, 3]
    with pytest.raises(AssertionError, match="Pressures must decrease"):
        parcel.valid_data({"pressure": Q})
    assert parcel.valid_data({"pressure": P}) is True           # the flags were consumed by the failed checks


@pytest.mark.parametrize("interpolate", [True, False])
@pytest.mark.parametrize("depth", [100, 300])
def test_get_layer_against_oracle(ctx, interpolate, depth):
    P, T, D = _columns(seed=25)
    d = {"pressure": P, "temperature": T, "dewpoint": D}
    ora = op.get_layer(d, depth=depth, interpolate=interpolate)
    got = parcel.get_layer(d, depth=depth, interpolate=interpolate)
    for key in d:
        _same(got[key], ora[key])


@pytest.mark.parametrize("log_x", [False, True])
def test_find_intersections_against_oracle(ctx, log_x):
    P, T, D = _columns(seed=26)
    A = T - 0.6 * (T - D) + 3.0 * np.sin(np.arange(P.shape[0]))[:, None]
    A[7, :30] = T[7, :30]
    ora = op.find_intersections(P, A, T, log_x=log_x)
    got = parcel.find_intersections(P, A, T, log_x=log_x)
    assert set(got.keys()) == set(ora.keys())
    for k in ora:
        _same(got[k], ora[k])
    # shared 1-D coordinate, float32, crossing with a constant (the freezing level, PF:2153)
    p1, t, _ = synth.era5_columns(2000, seed=10)
    Pb = np.broadcast_to(p1.numpy().astype(np.float64)[:, None], t.shape)
    t64 = t.numpy().astype(np.float64)
    o2 = op.find_intersections(Pb, t64, np.full_like(t64, 273.0), log_x=log_x)
    g2 = ctx.find_intersections(p1.cuda(), t.cuda(), torch.full((1, 1), 273.0, device="cuda"), log_x=log_x)
    _same(g2["all_intersect_x"].cpu().numpy(), o2["all_intersect_x"], rtol=2e-4)


@pytest.mark.parametrize("interpolator", ["log", "linear"])
def test_add_lcl_to_profile_against_oracle(ctx, interpolator):
    """PF:858-931 composed from the stand-alone kernels, against the oracle and against the fused kernel behind
    parcel_profile_with_lcl."""
    from oracle import tables as otab
    from oracle import thermo as th
    ctx.tables_build()
    p, t, td = synth.model_level_columns(1500, 40, seed=27, nan_columns=0, allnan_columns=0)
    P, T, D = [x.numpy().astype(np.float64) for x in (p, t, td)]
    opts = op.Options(op.MoistLapseLUT(otab.load_tables()), lcl_mode="converged")
    prof = op.parcel_profile(P, P[0], T[0], D[0], opts)
    env = {"temperature": T, "virtual_temperature": th.virtual_temperature(T, op.mixing_ratio(T, D, P, opts)),
           "dewpoint": D, "pressure": prof["pressure"]}
    ora = op.add_lcl_to_profile(prof, env, interpolator, opts)
    got = parcel.add_lcl_to_profile(prof, environment=env, interpolator=interpolator)
    assert set(got.keys()) == set(ora.keys())
    for k in ora:
        _same(got[k], ora[k], rtol=1e-11)
    fused = parcel.parcel_profile_with_lcl(P, T, D, P[0], T[0], D[0], lcl_interp=interpolator)
    for k in ("pressure", "environment_temperature", "environment_dewpoint", "environment_virtual_temperature"):
        _same(got[k], fused[k], rtol=1e-9)


def test_interp1d_is_numpy_interp(ctx):
    """interp1d_numba (PF:23-37) through xp_interp1d: rows of moist-adiabat-like curves on a shared and on a per-row
    xp, against numpy.interp row by row (bit-exact in float64)."""
    rng = np.random.default_rng(8)
    R, n, m = 64, 500, 90
    xp1 = np.sort(rng.uniform(2.0, 1100.0, n))
    xpr = np.sort(rng.uniform(2.0, 1100.0, (R, n)), axis=1)
    fp = rng.uniform(180.0, 310.0, (R, n))
    at = rng.uniform(-50.0, 1200.0, (R, m))
    at[:, 0] = xp1[5]
    at[3, 7] = np.nan
    got = parcel.interp1d_numba(at, xp1, fp)
    ref = np.stack([np.interp(at[r], xp1, fp[r]) for r in range(R)])
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(got[~np.isnan(ref)], ref[~np.isnan(ref)])
    got = parcel.interp1d_numba(at, xpr, fp)
    ref = np.stack([np.interp(at[r], xpr[r], fp[r]) for r in range(R)])
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(got[~np.isnan(ref)], ref[~np.isnan(ref)])
    g32 = ctx.interp1d(torch.from_numpy(at).float().cuda(), torch.from_numpy(xp1).float().cuda(),
                       torch.from_numpy(fp).float().cuda())
    assert g32.dtype == torch.float32 and g32.shape == (R, m)


@pytest.mark.parametrize("log_x", [True, False])
def test_trap_around_zeros_against_oracle(ctx, log_x):
    P, T, D = _columns(seed=28)
    Y = 0.6 * (T - D) - 4.0 + 3.0 * np.sin(np.arange(P.shape[0]))[:, None]
    Y[9, :25] = 0.0
    ora, omask = op.trap_around_zeros(P, Y, log_x=log_x)
    got, gmask = parcel.trap_around_zeros(P, Y, log_x=log_x)
    assert np.array_equal(np.asarray(gmask, dtype=bool), omask)
    assert set(got.keys()) == set(ora.keys())
    for k in ora:
        _same(got[k], ora[k])
