"""CPU check of the level primitives' per-column code (xp_levels.cuh compiled for the host by tests/hostsim,
test-only) against the oracle: insert_level (PF:933-990, incl. the reference's own known answer UT:1388-1411),
shift_out_nans (PF:1699-1720), trapz (PF:164-206) and the pressure check of valid_data (PF:2320).  The GPU twin
is tests/test_gpu_levels.py."""

import numpy as np
import pytest

import hostsim_util as hs
from oracle import parcel as op
from xarray_parcel_b200 import synth


def _same(a, b, rtol=1e-12):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(b)
    assert np.allclose(a[ok], b[ok], rtol=rtol, atol=0)


def _columns(n=1500, L=40, seed=21):
    p, t, td = synth.model_level_columns(n, L, seed=seed, nan_columns=0.1)
    return [x.numpy().astype(np.float64) for x in (p, t, td)]


def test_insert_level_against_oracle():
    P, T, _ = _columns()
    rng = np.random.default_rng(5)
    N = P.shape[1]
    lev_p = rng.uniform(50.0, 1100.0, N)             # inside, below and above the column
    lev_p[:40] = P[rng.integers(0, P.shape[0], 40), np.arange(40)]     # exactly on a level (kept below, PF:955)
    lev_p[40:50] = np.nan
    lev_t = rng.uniform(200.0, 300.0, N)
    P[5:9, 60:80] = np.nan                            # NaN coordinates inside a column (fill-value route)
    ora = op.insert_level({"pressure": P, "temperature": T}, {"pressure": lev_p, "temperature": lev_t}, "pressure")
    _same(hs.insert_level(P, P, lev_p, lev_p), ora["pressure"])
    _same(hs.insert_level(P, T, lev_p, lev_t), ora["temperature"])


def test_shift_out_nans_against_oracle():
    P, T, _ = _columns(seed=22)
    N = P.shape[1]
    k = np.random.default_rng(6).integers(0, 6, N)
    lead = np.arange(P.shape[0])[:, None] < k[None, :]
    Pn = np.where(lead, np.nan, P)
    Pn[:, :5] = np.nan                                # all-NaN columns
    ora = op.shift_out_nans({"pressure": Pn, "temperature": T}, "pressure")
    got_p, shift = hs.shift_out_nans(Pn, Pn)
    got_t, _ = hs.shift_out_nans(Pn, T)
    _same(got_p, ora["pressure"])
    _same(got_t, ora["temperature"])
    assert np.array_equal(shift[5:], np.where(np.isnan(P[0, 5:]), P.shape[0], k[5:]))


@pytest.mark.parametrize("sign", [0, 1, -1])
def test_trapz_against_oracle(sign):
    P, T, D = _columns(seed=23)
    V = T - D - 8.0                                   # changes sign
    mask = np.random.default_rng(7).random(P.shape) < 0.7
    for m in (None, mask):
        ora = op.trapz({"pressure": P, "v": V}, "pressure", mask=m, only_positive=sign > 0, only_negative=sign < 0)
        _same(hs.trapz(P, V, mask=m, sign=sign), ora["v"])
        _same(hs.trapz(P, P, mask=m, sign=sign), ora["pressure"])
    _same(hs.trapz(np.log(P), V, sign=sign),
          op.trapz({"x": np.log(P), "v": V}, "x", only_positive=sign > 0, only_negative=sign < 0)["v"])


def test_pressure_order():
    P, _, _ = _columns(seed=24)
    assert hs.pressure_order(P) == 2                  # valid: differences seen, none >= 0
    Q = P.copy()
    Q[[3, 4], 17] = Q[[4, 3], 17]
    assert hs.pressure_order(Q) == 3
    Q = P.copy()
    Q[7, 3] = Q[6, 3]                                 # equal pressures are not "decreasing" (max < 0 fails)
    assert hs.pressure_order(Q) & 1
    assert hs.pressure_order(np.full((5, 3), np.nan)) == 0


@pytest.mark.parametrize("log_x", [False, True])
def test_find_intersections_against_oracle(log_x, soundings):
    P, T, D = _columns(seed=26)
    A = T - 0.6 * (T - D) + 3.0 * np.sin(np.arange(P.shape[0]))[:, None]          # crosses T several times
    A[7, :30] = T[7, :30]                                                            # exact touches (sign 0)
    ora = op.find_intersections(P, A, T, log_x=log_x)
    got = hs.find_intersections(P, A, T, log_x=log_x)
    for k in ora:
        _same(got[k], ora[k])


def test_interp1d_is_numpy_interp():
    """interp1d_numba (PF:23-37) == numpy.interp: inside, on the nodes, outside (end values), NaN points, NaN values."""
    rng = np.random.default_rng(8)
    xp = np.sort(rng.uniform(2.0, 1100.0, 300))
    fp = rng.uniform(180.0, 310.0, 300)
    at = np.concatenate([rng.uniform(-50.0, 1200.0, 2000), xp[::7], [xp[0], xp[-1], np.nan]])
    got, ref = hs.interp1d(at, xp, fp), np.interp(at, xp, fp)
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(got[~np.isnan(ref)], ref[~np.isnan(ref)])
    fp[40:43] = np.nan
    got, ref = hs.interp1d(at, xp, fp), np.interp(at, xp, fp)
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(got[~np.isnan(ref)], ref[~np.isnan(ref)])
    assert hs.interp1d(np.array([1.0, 5.0, 9.0]), np.array([5.0]), np.array([7.0])).tolist() == [7.0, 7.0, 7.0]


@pytest.mark.parametrize("log_x", [True, False])
def test_trap_around_zeros_against_oracle(log_x):
    P, T, D = _columns(seed=28)
    Y = 0.6 * (T - D) - 4.0 + 3.0 * np.sin(np.arange(P.shape[0]))[:, None]          # crosses zero several times
    Y[9, :25] = 0.0                                                                  # exact zeros
    ora, omask = op.trap_around_zeros(P, Y, log_x=log_x)
    got, gmask = hs.trap_around_zeros(P, Y, log_x=log_x)
    assert np.array_equal(gmask, omask)
    for k in ora:
        _same(got[k], ora[k])


def test_size_independent_properties():
    """Properties that hold whatever the input: trapz is linear in the integrand and additive over the sign split,
    shift_out_nans is idempotent, insert_level keeps every original level in order and adds exactly one, and a
    mixed-layer mean of a constant is that constant."""
    P, T, D = _columns(n=2000, L=45, seed=29)
    ok = ~np.isnan(P[0])
    P, T, D = P[:, ok], np.nan_to_num(T[:, ok], nan=250.0), np.nan_to_num(D[:, ok], nan=240.0)
    V = T - D - 8.0
    # trapz: linearity, and all = positive + negative parts
    a, b = hs.trapz(P, T), hs.trapz(P, V)
    np.testing.assert_allclose(hs.trapz(P, 2.0 * T - 3.0 * V), 2.0 * a - 3.0 * b, rtol=1e-10)
    np.testing.assert_allclose(hs.trapz(P, V, sign=1) + hs.trapz(P, V, sign=-1), b, rtol=1e-10, atol=1e-9)
    # shift_out_nans: idempotent, and the shifted column starts with a value
    lead = np.arange(P.shape[0])[:, None] < np.random.default_rng(3).integers(0, 5, P.shape[1])[None, :]
    Pn = np.where(lead, np.nan, P)
    s1, k1 = hs.shift_out_nans(Pn, Pn)
    s2, k2 = hs.shift_out_nans(s1, s1)
    assert not np.isnan(s1[0]).any() and (k2 == 0).all() and np.array_equal(np.isnan(s1), np.isnan(s2))
    assert np.array_equal(s1[~np.isnan(s1)], s2[~np.isnan(s2)])
    # insert_level: one more level, sorted, the original levels survive in order
    lev_p = P[0] - np.random.default_rng(4).uniform(1.0, 400.0, P.shape[1])
    out_p = hs.insert_level(P, P, lev_p, lev_p)
    out_t = hs.insert_level(P, T, lev_p, np.full(P.shape[1], -1.0))
    assert out_p.shape[0] == P.shape[0] + 1 and (np.diff(out_p, axis=0) <= 0).all()
    assert ((out_t == -1.0).sum(axis=0) == 1).all()
    kept = out_t.T[(out_t != -1.0).T].reshape(P.shape[1], P.shape[0]).T
    assert np.array_equal(kept, T)
    # mixed layer of a constant field
    (m,) = hs.mixed_layer(P, [np.full_like(P, 7.25)], depth=80.0)
    np.testing.assert_allclose(m, 7.25, rtol=1e-12)
