"""CPU check of the KERNEL LOGIC: the per-column device functions (xarray_parcel_b200/csrc/
xp_parcels.cuh, the code the sm_100a kernels inline) compiled for the host with g++
(tests/hostsim, test-only) against the whole-array oracle in lookup-table mode.

The two are structured completely differently (single upward sweep with running sums and
snapshots vs. the reference's masked whole-array passes), so agreement to rounding on
thousands of columns, NaN patterns included, is strong evidence for both.
"""

import numpy as np
import pytest

import hostsim_util as hs
from oracle import parcel as op
from xarray_parcel_b200 import synth

FIELDS = ["cape", "cin", "lcl_pressure", "lcl_temperature", "lcl_virtual_temperature",
          "lfc_pressure", "lfc_temperature", "el_pressure", "el_temperature"]
RTOL = 1e-9      # both sides are float64; differences come from summation order only


def _compare(res, ora, prefix):
    for f in FIELDS:
        a, b = res[f], ora[prefix + f]
        assert np.array_equal(np.isnan(a), np.isnan(b)), f"{prefix}{f}: NaN pattern differs"
        ok = ~np.isnan(b)
        err = np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1.0)
        assert err.size == 0 or err.max() < RTOL, f"{prefix}{f}: max rel err {err.max():.3e}"


OPTION_SETS = [
    dict(vtc=True, lcl_interp="log", pos_cape_neg_cin=True, compat="1.4.1"),     # reference defaults
    dict(vtc=False, lcl_interp="linear", pos_cape_neg_cin=True, compat="1.4.1"),  # "MetPy mode"
    dict(vtc=True, lcl_interp="linear", pos_cape_neg_cin=False, compat="1.6.2"),
    dict(vtc=False, lcl_interp="log", pos_cape_neg_cin=False, compat="1.4.1"),
]


@pytest.mark.parametrize("o", OPTION_SETS, ids=lambda o: f"vtc{int(o['vtc'])}-{o['lcl_interp']}-pn{int(o['pos_cape_neg_cin'])}-{o['compat']}")
@pytest.mark.parametrize("shape", ["model70", "era5"])
def test_suite_matches_oracle(oracle_tables, o, shape):
    if shape == "model70":
        p, t, td = [a.numpy().astype(np.float64) for a in synth.model_level_columns(3000, 70, seed=2)]
        pfull = p
    else:
        p, t, td = [a.numpy().astype(np.float64) for a in synth.era5_columns(3000, seed=3, nan_columns=0.02)]
        pfull = np.broadcast_to(p[:, None], t.shape)
    opts = op.Options(op.MoistLapseLUT(oracle_tables), lcl_mode="converged", metpy_compat=o["compat"])
    ora = op.suite(pfull, t, td, opts, virtual_temperature_correction=o["vtc"],
                   lcl_interp=o["lcl_interp"], pos_cape_neg_cin=o["pos_cape_neg_cin"])
    for kind in ("sb", "ml", "mu"):
        res = hs.cape_cin(p, t, td, oracle_tables, kind=kind, vtc=o["vtc"], lcl_interp=o["lcl_interp"],
                          pos_cape_neg_cin=o["pos_cape_neg_cin"],
                          metpy_compat=141 if o["compat"] == "1.4.1" else 162)
        assert res["flags"] == 0
        _compare(res, ora, kind + "_")
        if kind != "sb":
            for f in ("pressure", "temperature", "dewpoint"):
                a, b = res["parcel_" + f], ora[f"{kind}_parcel_{f}"]
                assert np.array_equal(np.isnan(a), np.isnan(b))
                ok = ~np.isnan(b)
                assert np.allclose(a[ok], b[ok], rtol=1e-12, atol=0)


@pytest.mark.parametrize("kind", ["sb", "ml", "mu"])
def test_profile_rows_match_oracle(oracle_tables, kind):
    """parcel_profile_with_lcl output (PF:806-931): every row of the (L+1)-level profile."""
    p, t, td = [a.numpy().astype(np.float64) for a in synth.model_level_columns(800, 50, seed=5)]
    opts = op.Options(op.MoistLapseLUT(oracle_tables), lcl_mode="converged")
    fn = {"sb": op.surface_based_cape_cin, "ml": op.mixed_layer_cape_cin,
          "mu": op.most_unstable_cape_cin}[kind]
    prof = fn(p, t, td, opts)[1]
    res = hs.cape_cin(p, t, td, oracle_tables, kind=kind, profile=True)
    n = prof["pressure"].shape[0]           # the oracle trims with dropna(how='all') like PF:1552/1637
    for k in hs.PROFILE:
        a, b = res["profile"][k][:n], prof[k]
        assert np.array_equal(np.isnan(a), np.isnan(b)), f"{k}: NaN pattern differs"
        ok = ~np.isnan(b)
        assert np.allclose(a[ok], b[ok], rtol=1e-11, atol=0), k
        assert np.isnan(res["profile"][k][n:]).all()


def test_explicit_parcel_matches_oracle(oracle_tables):
    p, t, td = [a.numpy().astype(np.float64) for a in synth.model_level_columns(500, 40, seed=9)]
    pp, pt, pd_ = p[3], t[3] + 1.0, td[3]
    opts = op.Options(op.MoistLapseLUT(oracle_tables), lcl_mode="converged")
    cc, prof = op.cape_cin(p, t, td, pt, pp, pd_, opts)
    ora = {"x_" + k: v for k, v in {**cc, **prof}.items() if v.ndim == 1}
    res = hs.cape_cin(p, t, td, oracle_tables, kind="explicit", explicit=(pp, pt, pd_))
    _compare(res, ora, "x_")


def test_reference_soundings_lut_mode(oracle_tables, soundings):
    """Every sounding of the reference's unit tests, lookup-table mode, all three parcels."""
    names = [n for n, s in soundings.items()
             if all(k in s for k in ("levels", "temperatures", "dewpoints")) and s["levels"].size > 3]
    L = max(soundings[n]["levels"].size for n in names)

    def pad(a):
        return np.concatenate([a, np.full(L - a.size, np.nan)])

    P = np.stack([pad(soundings[n]["levels"]) for n in names], axis=1)
    T = np.stack([pad(soundings[n]["temperatures"]) for n in names], axis=1)
    D = np.stack([pad(soundings[n]["dewpoints"]) for n in names], axis=1)
    opts = op.Options(op.MoistLapseLUT(oracle_tables), lcl_mode="converged")
    ora = op.suite(P, T, D, opts)
    for kind in ("sb", "ml", "mu"):
        _compare(hs.cape_cin(P, T, D, oracle_tables, kind=kind), ora, kind + "_")


def test_lut_mode_reproduces_reference_pins_loosely(oracle_tables, soundings):
    """The reference's own statement about its lookup table (UT:106-112: moist-lapse tests pass at
    2 decimals; demo notebook: SB CAPE within ~1 % of MetPy): SB CAPE/CIN of UT:940-951."""
    s = soundings["test_surface_based_cape_cin"]
    r = hs.cape_cin(s["levels"][:, None], s["temperatures"][:, None], s["dewpoints"][:, None],
                    oracle_tables, kind="sb")
    assert abs(r["cape"][0] - 230.1982) / 230.1982 < 0.02
    assert abs(r["cin"][0] - (-58.0673)) < 1.0
