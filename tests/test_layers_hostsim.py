"""CPU check of the layer primitives' per-column code (xp_layers.cuh compiled for the host by tests/hostsim,
test-only) against the oracle and the reference's known answers (UT:1142-1177).  The GPU twin of this file is
tests/test_gpu_layers.py."""

import numpy as np
import pytest

import hostsim_util as hs
from oracle import parcel as op
from oracle import thermo as th
from xarray_parcel_b200 import synth

K = 273.15


def col(a):
    return np.asarray(a, dtype=np.float64)[:, None]


def _same(a, b, rtol=1e-12):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(b)
    assert np.allclose(a[ok], b[ok], rtol=rtol, atol=0)


def test_known_answers(soundings):
    s = soundings["test_mixed_layer"]                             # UT:1170-1177
    (m,) = hs.mixed_layer(col(s["pressure"]), [col(s["temperature"])], depth=250)
    np.testing.assert_almost_equal(m[0], 16.4024930 + K, 6)
    s = soundings["test_mixed_parcel"]                            # UT:1142-1153
    mp = hs.mixed_parcel(col(s["levels"]), col(s["temperatures"]), col(s["dewpoints"]), depth=250)
    np.testing.assert_almost_equal(mp["pressure"][0], 959., 6)
    np.testing.assert_almost_equal(mp["temperature"][0], 28.7401463 + K, 6)
    np.testing.assert_almost_equal(mp["dewpoint"][0], 7.1534658 + K, 6)


@pytest.mark.parametrize("depth", [50, 100, 250])
def test_mixed_layer_against_oracle(depth):
    p, t, td = synth.model_level_columns(3000, 60, seed=11, nan_columns=0.05)
    P, T, D = [x.numpy().astype(np.float64) for x in (p, t, td)]
    dat = {"pressure": P, "temperature": T, "dewpoint": D, "theta": th.potential_temperature(P, T),
           "mixing_ratio": th.saturation_mixing_ratio(P, D), "height": 7500.0 * np.log(P[0][None, :] / P)}
    ora = op.mixed_layer(dat, depth=depth)
    got = hs.mixed_layer(P, list(dat.values()), depth=depth, pressure_field=0)
    for k, g in zip(dat, got):
        _same(g, ora[k])
    mp = hs.mixed_parcel(P, T, D, depth=depth)
    omp = op.mixed_parcel(P, T, D, depth=depth)
    for k in mp:
        _same(mp[k], omp[k])


def test_shared_axis():
    p1, t, td = synth.era5_columns(2000, seed=12)
    T = t.numpy().astype(np.float64)
    P = np.broadcast_to(p1.numpy().astype(np.float64)[:, None], T.shape)
    (got,) = hs.mixed_layer(p1.numpy().astype(np.float64), [T], depth=100)
    _same(got, op.mixed_layer({"pressure": P, "temperature": T}, depth=100)["temperature"])


def test_edges():
    """A level exactly on the layer top, a column shallower than the layer, NaN values inside the layer, an
    all-NaN column, a single level (see tests/test_gpu_layers.py)."""
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
        (got,) = hs.mixed_layer(P, [X], depth=depth)
        _same(got, op.mixed_layer({"pressure": P, "x": X}, depth=depth)["x"])
    (got,) = hs.mixed_layer(P[:1].copy(), [X[:1].copy()], depth=100)
    _same(got, op.mixed_layer({"pressure": P[:1], "x": X[:1]}, depth=100)["x"])


@pytest.mark.parametrize("depth", [100, 300])
def test_layer_bounds(depth):
    p, _, _ = synth.model_level_columns(2000, 50, seed=14)
    P = p.numpy().astype(np.float64)
    ob = op.nanmax(P)
    bottom, top = hs.layer_bounds(P, P.shape[1], depth=depth, interpolate=False)
    _same(bottom, ob)
    _same(top, op.bound_pressure(P, ob - depth))
    _, top_i = hs.layer_bounds(P, P.shape[1], depth=depth, interpolate=True)
    _same(top_i, ob - depth)
    _, tt = hs.layer_bounds(np.array([[1000.], [920.], [880.], [700.]]), 1, depth=100, interpolate=False)
    assert tt[0] == 920.0
