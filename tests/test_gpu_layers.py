"""GPU tests of the layer primitives exposed on their own (xp_layers.cu through the C ABI): mixed_layer
(PF:137-162), mixed_parcel with all six returned variables (PF:229-289) and get_layer's bounds (PF:63-100,
bound_pressure PF:208-227), against the oracle and the reference's known answers (UT:1142-1177)."""

import numpy as np
import pytest
import torch

from oracle import parcel as op
from oracle import thermo as th
from xarray_parcel_b200 import _lib, synth
import xarray_parcel_b200.parcel_functions as parcel

pytestmark = pytest.mark.gpu

K = 273.15


@pytest.fixture(scope="module")
def ctx():
    return _lib.get_context(0)


def col(a):
    return np.asarray(a, dtype=np.float64)[:, None]


def _same(a, b, rtol=1e-12):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(b)
    assert np.allclose(a[ok], b[ok], rtol=rtol, atol=0)


def test_mixed_layer_known_answer(soundings):                    # UT:1170-1177
    s = soundings["test_mixed_layer"]
    m = parcel.mixed_layer({"pressure": col(s["pressure"]), "temperature": col(s["temperature"])}, depth=250)
    np.testing.assert_almost_equal(m["temperature"][0], 16.4024930 + K, 6)


def test_mixed_parcel_known_answer(soundings):                   # UT:1142-1153
    s = soundings["test_mixed_parcel"]
    m = parcel.mixed_parcel(col(s["levels"]), col(s["temperatures"]), col(s["dewpoints"]), depth=250)
    np.testing.assert_almost_equal(m["pressure"][0], 959., 6)
    np.testing.assert_almost_equal(m["temperature"][0], 28.7401463 + K, 6)
    np.testing.assert_almost_equal(m["dewpoint"][0], 7.1534658 + K, 6)
    assert set(m.keys()) == {"theta", "mixing_ratio", "temperature", "vapour_pressure", "dewpoint", "pressure"}


@pytest.mark.parametrize("depth", [50, 100, 250])
def test_mixed_layer_against_oracle(ctx, depth):
    p, t, td = synth.model_level_columns(5000, 60, seed=11, nan_columns=0.05)
    P, T, D = [x.numpy().astype(np.float64) for x in (p, t, td)]
    dat = {"pressure": P, "temperature": T, "dewpoint": D, "theta": th.potential_temperature(P, T),
           "mixing_ratio": th.saturation_mixing_ratio(P, D), "height": 7500.0 * np.log(P[0][None, :] / P)}
    ora = op.mixed_layer(dat, depth=depth)
    got = parcel.mixed_layer(dat, depth=depth)                   # six variables: two kernel launches
    assert list(got.keys()) == list(dat.keys())
    for k in dat:
        _same(got[k], ora[k])


def test_mixed_layer_shared_axis_float32(ctx):
    p1, t, td = synth.era5_columns(4000, seed=12)
    P = np.broadcast_to(p1.numpy().astype(np.float64)[:, None], t.shape)
    ora = op.mixed_layer({"pressure": P, "temperature": t.numpy().astype(np.float64)}, depth=100)
    (got,) = ctx.mixed_layer(p1.cuda(), [t.cuda()], depth=100)
    assert got.dtype == torch.float32
    _same(got.cpu().numpy(), ora["temperature"], rtol=2e-7)


def test_mixed_layer_edges(ctx):
    """A level exactly on the layer top (duplicated by insert_level, PF:965), a column shallower than the layer
    (the top level carries NaN values, its trapezoid is skipped), NaN values inside the layer, an all-NaN column,
    a single level."""
    nan = np.nan
    P = np.array([[1000., 1000., 1000., nan, 1000.],
                  [950., 980., 950., nan, 950.],
                  [900., 960., 900., nan, 910.],
                  [850., nan, 850., nan, 890.],
                  [700., nan, 800., nan, 600.]])
    X = np.array([[10., 1., 3., nan, 2.],
                  [12., 2., nan, nan, 4.],
                  [15., 4., 5., nan, 8.],
                  [11., nan, 7., nan, 16.],
                  [5., nan, 9., nan, 32.]])
    for depth in (100, 150, 45):
        ora = op.mixed_layer({"pressure": P, "x": X}, depth=depth)
        (got,) = ctx.mixed_layer(torch.from_numpy(P).cuda(), [torch.from_numpy(X).cuda()], depth=depth)
        _same(got.cpu().numpy(), ora["x"])
    ora = op.mixed_layer({"pressure": P[:1], "x": X[:1]}, depth=100)
    (got,) = ctx.mixed_layer(torch.from_numpy(P[:1].copy()).cuda(), [torch.from_numpy(X[:1].copy()).cuda()], depth=100)
    _same(got.cpu().numpy(), ora["x"])


@pytest.mark.parametrize("depth", [100, 50])
def test_mixed_parcel_against_oracle(ctx, depth):
    p, t, td = synth.model_level_columns(5000, 70, seed=13, nan_columns=0.05)
    P, T, D = [x.numpy().astype(np.float64) for x in (p, t, td)]
    ora = op.mixed_parcel(P, T, D, depth=depth)
    got = parcel.mixed_parcel(P, T, D, depth=depth)
    for k in ("theta", "mixing_ratio", "temperature", "vapour_pressure", "dewpoint", "pressure"):
        _same(got[k], ora[k])
    # the fused lifting kernel selects the same parcel
    _, _, mp = parcel.mixed_layer_cape_cin(P, T, D, depth=depth)
    for k in ("pressure", "temperature", "dewpoint"):
        _same(mp[k], ora[k], rtol=1e-10)
    # float32 I/O
    got32 = parcel.mixed_parcel(p.numpy(), t.numpy(), td.numpy(), depth=depth)
    assert got32["temperature"].dtype == np.float32
    _same(got32["temperature"], ora["temperature"], rtol=3e-7)
    _same(got32["dewpoint"], ora["dewpoint"], rtol=3e-7)


@pytest.mark.parametrize("depth", [100, 300])
def test_layer_bounds(ctx, depth):
    p, _, _ = synth.model_level_columns(3000, 50, seed=14)
    P = p.numpy().astype(np.float64)
    Pg = torch.from_numpy(P).cuda()
    bottom, top = ctx.layer_bounds(Pg, P.shape[1], depth=depth, interpolate=False)
    ob = op.nanmax(P)
    _same(bottom.cpu().numpy(), ob)
    _same(top.cpu().numpy(), op.bound_pressure(P, ob - depth))
    _same(parcel.bound_pressure(P, depth=depth), op.bound_pressure(P, ob - depth))
    _, top_i = ctx.layer_bounds(Pg, P.shape[1], depth=depth, interpolate=True)
    _same(top_i.cpu().numpy(), ob - depth)
    # a tie between two levels goes to the larger pressure (PF:226)
    Pt = np.array([[1000.], [920.], [880.], [700.]])
    _, tt = ctx.layer_bounds(torch.from_numpy(Pt).cuda(), 1, depth=100, interpolate=False)
    assert float(tt[0]) == 920.0


def test_bad_arguments(ctx):
    x = torch.zeros((4, 8), dtype=torch.float64, device="cuda")
    st = ctx.lib.xp_mixed_layer(ctx.handle, x.data_ptr(), 8, 0, None, None, 1, -1, 8, 4, 8, _lib.XP_F64, 100.0, None)
    assert st != _lib.XP_OK
    assert b"mixed_layer" in ctx.lib.xp_last_error(ctx.handle)
