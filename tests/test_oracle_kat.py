"""Pin the CPU oracle against every known-answer test the reference ships for the parcel
path (/root/reference/modules/unit_tests.py, "UT"; run list UT:19-77).

As in the reference (parcel_functions_demo.ipynb cell 33: ``parcel.moist_lapse =
tests.metpy_moist_lapse``) the known answers hold for the *exact* moist adiabat, so the
oracle runs in ODE mode here; the lookup-table mode is checked at the reference's own
looser tolerance (UT:106-112) in test_oracle_tables.py.  Expected values and decimals are
the reference's ``assert_almost_equal`` arguments, cited by UT line.
"""

import numpy as np
import pytest
from numpy.testing import assert_almost_equal, assert_array_almost_equal

from oracle import parcel as op
from oracle import thermo as th

# MetPy-1.4.1-faithful numerics (SciPy fixed_point LCL at xtol 1e-5, odeint moist adiabat at
# default tolerances) -- what produced the pins; the most sensitive pin (test_el, 31 hPa/K)
# resolves MetPy's own solver error of ~4e-5 K.
ODE = op.Options(op.MoistLapseODE(solver="odeint-1.4.1"), metpy_compat="1.4.1", lcl_mode="scipy")
MP = dict(virtual_temperature_correction=False, lcl_interp="linear")   # "MetPy mode" kwargs
K = 273.15


def col(a):
    return np.asarray(a, dtype=np.float64)[:, None]


def sfc_profile(s, lcl_interp="linear", parcel=None):
    p, t, td = col(s["levels"]), col(s["temperatures"]), col(s["dewpoints"])
    if parcel is None:
        parcel = (p[0], t[0], td[0])
    return op.parcel_profile_with_lcl(p, t, td, parcel[0], parcel[1], parcel[2], ODE,
                                      lcl_interp=lcl_interp)


def lfc_el_T(profile):
    """lfc_el on real temperature, as the UT tests call it."""
    return op.lfc_el(profile["pressure"], profile["temperature"],
                     profile["environment_temperature"], profile["lcl_pressure"],
                     profile["lcl_temperature"])


def no_lcl_profile(s, parcel=None, add=0.0):
    """UT pattern 'Calculate parcel profile without LCL, as per metpy unit tests'."""
    p, t, td = col(s["levels"]), col(s["temperatures"]), col(s["dewpoints"])
    if parcel is None:
        parcel = (p[0], t[0], td[0])
    prof = op.parcel_profile(p, parcel[0], parcel[1], parcel[2], ODE)
    prof["temperature"] = prof["temperature"] + add
    prof["environment_temperature"] = t
    return prof


# ---- dry / moist lapse, parcel profile, LCL ----------------------------------------------
def test_dry_lapse(soundings):                                   # UT:154-158
    lv = soundings["test_dry_lapse"]["levels"]
    assert_array_almost_equal(th.dry_lapse(lv, 303.15, lv.max()), [303.15, 294.16, 290.83], 2)


def test_dry_lapse_2_levels(soundings):                          # UT:160-164
    lv = soundings["test_dry_lapse_2_levels"]["levels"]
    assert_array_almost_equal(th.dry_lapse(lv, 293., lv.max()), [293., 240.3583], 4)


def test_moist_lapse(soundings):                                 # UT:166-170
    lv = soundings["test_moist_lapse"]["levels"]
    t = ODE.moist_lapse(col(lv), 293., lv[0])[:, 0]
    assert_array_almost_equal(t, [293, 284.64, 272.81, 264.42, 252.91], 2)


def test_moist_lapse_ref_pres(soundings):                        # UT:172-176
    lv = soundings["test_moist_lapse_ref_pres"]["levels"]
    t = ODE.moist_lapse(col(lv), 293., 1000.)[:, 0]
    assert_array_almost_equal(t, [294.76, 284.64, 272.81, 264.42, 252.91], 2)


def test_moist_lapse_scalar(soundings):                          # UT:178-182
    t = ODE.moist_lapse(col([800.]), 293., 1000.)[:, 0]
    assert_array_almost_equal(t, [284.64], 2)


def test_moist_lapse_uniform(soundings):                         # UT:184-188
    t = ODE.moist_lapse(col([900., 900., 900.]), 293.15, 900.)[:, 0]
    assert_almost_equal(t, np.array([293.15, 293.15, 293.15]), 7)


def test_parcel_profile(soundings):                              # UT:190-203
    s = soundings["test_parcel_profile"]
    prof = op.parcel_profile(col(s["levels"]), s["parcel_pressure"], s["parcel_temperature"],
                             s["parcel_dewpoint"], ODE)
    assert_array_almost_equal(prof["temperature"][:, 0], s["true_prof"], 2)


def test_parcel_profile_lcl(soundings):                          # UT:205-230
    s = soundings["test_parcel_profile_lcl"]
    prof = op.parcel_profile(col(s["p"]), s["parcel_pressure"], s["parcel_temperature"],
                             s["parcel_dewpoint"], ODE)
    env = {"temperature": col(s["t"]), "pressure": prof["pressure"]}
    prof = op.add_lcl_to_profile(prof, env, "linear", ODE)
    assert_array_almost_equal(prof["pressure"][:, 0], s["true_p"], 3)
    assert_array_almost_equal(prof["environment_temperature"][:, 0], s["true_t"], 3)
    assert_array_almost_equal(prof["temperature"][:, 0], s["true_prof"], 2)


def test_parcel_profile_saturated(soundings):                    # UT:232-245
    s = soundings["test_parcel_profile_saturated"]
    prof = op.parcel_profile(col(s["levels"]), s["parcel_pressure"], s["parcel_temperature"],
                             s["parcel_dewpoint"], ODE)
    assert_array_almost_equal(prof["temperature"][:, 0], s["true_prof"], 2)


def test_lcl(soundings):                                         # UT:247-256
    s = soundings["test_lcl"]
    r = op.lcl(s["parcel_pressure"], s["parcel_temperature"], s["parcel_dewpoint"], ODE)
    assert_almost_equal(r["lcl_pressure"], 864.806, 2)
    assert_almost_equal(r["lcl_temperature"], 17.676 + K, 2)


@pytest.mark.parametrize("mode", ["scipy", "converged"])
def test_lcl_nans(soundings, mode):                              # UT:258-270 (defined, not run upstream)
    s = soundings["test_lcl_nans"]
    o = op.Options(op.MoistLapseODE(), lcl_mode=mode)
    r = op.lcl(s["p"], s["t"], s["d"], o)
    assert_array_almost_equal(r["lcl_pressure"],
                              [np.nan, 836.4098648012595, np.nan, 836.4098648012595])
    assert_array_almost_equal(r["lcl_temperature"],
                              np.array([np.nan, 18.82281982535794, np.nan, 18.82281982535794]) + K)


def test_lcl_convergence_issue(soundings):                       # UT:1297-1306
    s = soundings["test_lcl_convergence_issue"]
    r = op.lcl(s["pressure"][0], s["temperatures"][0], s["dewpoints"][0], ODE)
    assert_almost_equal(r["lcl_pressure"], 990, 0)


def test_lcl_grid_surface_lcls(soundings):                       # UT:1338-1351
    s = soundings["test_lcl_grid_surface_lcls"]
    r = op.lcl(np.array([1000., 990, 1010]), np.array([15., 14, 13]) + K,
               np.array([15., 10, 13]) + K, ODE)
    assert_array_almost_equal(r["lcl_pressure"], s["pres_truth"], 4)
    assert_array_almost_equal(r["lcl_temperature"], s["temp_truth"], 4)


def test_parcel_profile_below_lcl(soundings):                    # UT:1278-1295
    s = soundings["test_parcel_profile_below_lcl"]
    prof = op.parcel_profile(col(s["pressure"]), s["pressure"][0], s["parcel_temperature"],
                             s["parcel_dewpoint"], ODE)
    assert_array_almost_equal(prof["temperature"][:, 0], s["truth"], 6)


# ---- LFC / EL from a surface parcel (profile with LCL, 'linear', real temperature) -------
NAN = np.nan
LFC_EL_SFC = [
    # UT function, lcl_interp, {key: (value, decimals)}
    ("test_lfc_basic", "linear", {"lfc_pressure": (727.371, 2), "lfc_temperature": (9.705 + K, 2)}),
    ("test_no_lfc", "linear", {"lfc_pressure": (NAN, 0), "lfc_temperature": (NAN, 0)}),
    ("test_lfc_inversion", "linear", {"lfc_pressure": (705.8806, 2),
                                      "lfc_temperature": (10.6232 + K, 2)}),
    ("test_lfc_equals_lcl", "linear", {"lfc_pressure": (777.0786, 2),
                                       "lfc_temperature": (15.8714 + K, 2)}),
    ("test_sensitive_sounding", "linear", {"lfc_pressure": (947.422, 2),
                                           "lfc_temperature": (20.498 + K, 2)}),
    ("test_lfc_sfc_precision", "linear", {"lfc_pressure": (NAN, 0), "lfc_temperature": (NAN, 0)}),
    ("test_lfc_pos_area_below_lcl", "linear", {"lfc_pressure": (NAN, 0),
                                               "lfc_temperature": (NAN, 0)}),
    ("test_el", "linear", {"el_pressure": (471.83286, 3), "el_temperature": (-11.5603 + K, 3)}),
    ("test_no_el", "linear", {"el_pressure": (NAN, 0), "el_temperature": (NAN, 0)}),
    ("test_no_el_multi_crossing", "linear", {"el_pressure": (NAN, 0), "el_temperature": (NAN, 0)}),
    ("test_lfc_and_el_below_lcl", "linear", {"el_pressure": (NAN, 0), "el_temperature": (NAN, 0),
                                             "lfc_pressure": (NAN, 0),
                                             "lfc_temperature": (NAN, 0)}),
    ("test_el_lfc_equals_lcl", "linear", {"el_pressure": (175.7663, 3),
                                          "el_temperature": (-57.03994 + K, 3)}),
    ("test_el_small_surface_instability", "linear", {"el_pressure": (NAN, 0),
                                                     "el_temperature": (NAN, 0)}),
    ("test_no_el_parcel_colder", "linear", {"el_pressure": (NAN, 0), "el_temperature": (NAN, 0)}),
    ("test_el_below_lcl", "linear", {"el_pressure": (NAN, 0), "el_temperature": (NAN, 0)}),
    ("test_lfc_not_below_lcl", "log", {"lfc_pressure": (811.618879, 3),
                                       "lfc_temperature": (6.48644650 + K, 3)}),
    ("multiple_intersections", "linear", {"lfc_pressure": (884.14790, 3),
                                          "lfc_temperature": (13.95707016 + K, 3),
                                          "el_pressure": (228.151466, 3),
                                          "el_temperature": (-56.81015490 + K, 3)}),
]


@pytest.mark.parametrize("name,interp,expect", LFC_EL_SFC, ids=[c[0] for c in LFC_EL_SFC])
def test_lfc_el_surface(soundings, name, interp, expect):        # UT:272-826, 1179-1249
    r = lfc_el_T(sfc_profile(soundings[name], lcl_interp=interp))
    for key, (val, dp) in expect.items():
        if np.isnan(val):
            assert np.isnan(r[key][0]), (key, r[key])
        else:
            assert_almost_equal(r[key][0], val, dp)


# ---- mixed parcel / mixed layer ------------------------------------------------------------
def _mixed(s, depth=100):
    return op.mixed_parcel(col(s["levels"]), col(s["temperatures"]), col(s["dewpoints"]),
                           depth=depth)


@pytest.mark.parametrize("name,expect", [
    ("test_lfc_ml", {"lfc_pressure": (601.225, 2), "lfc_temperature": (-1.90688 + K, 2)}),   # UT:293-313
    ("test_lfc_ml2", {"lfc_pressure": (962.34, 2), "lfc_temperature": (0.767 + K, 2)}),      # UT:315-364
    ("test_el_ml", {"el_pressure": (350.0561, 3), "el_temperature": (-28.36156 + K, 3)}),    # UT:609-630
])
def test_lfc_el_mixed_parcel(soundings, name, expect):
    s = soundings[name]
    m = _mixed(s)
    r = lfc_el_T(sfc_profile(s, parcel=(m["pressure"], m["temperature"], m["dewpoint"])))
    for key, (val, dp) in expect.items():
        assert_almost_equal(r[key][0], val, dp)


def test_lfc_intersection(soundings):                            # UT:366-386
    s = soundings["test_lfc_intersection"]
    m = _mixed(s)
    prof = no_lcl_profile(s, parcel=(m["pressure"], m["temperature"], m["dewpoint"]))
    assert_almost_equal(lfc_el_T(prof)["lfc_pressure"][0], 981.620, 2)


def test_mixed_parcel(soundings):                                # UT:1142-1153
    m = _mixed(soundings["test_mixed_parcel"], depth=250)
    assert_almost_equal(m["pressure"][0], 959., 6)
    assert_almost_equal(m["temperature"][0], 28.7401463 + K, 6)
    assert_almost_equal(m["dewpoint"][0], 7.1534658 + K, 6)


def test_mixed_layer(soundings):                                 # UT:1170-1177
    s = soundings["test_mixed_layer"]
    m = op.mixed_layer({"pressure": col(s["pressure"]), "temperature": col(s["temperature"])},
                       depth=250)
    assert_almost_equal(m["temperature"][0], 16.4024930 + K, 6)


def test_most_unstable_parcel(soundings):                        # UT:909-923
    s = soundings["test_most_unstable_parcel"]
    r = op.most_unstable_parcel({"pressure": col(s["levels"]),
                                 "temperature": col(s["temperatures"]),
                                 "dewpoint": col(s["dewpoints"])}, depth=100)
    assert_almost_equal(r["pressure"][0], 959.0, 6)
    assert_almost_equal(r["temperature"][0], 22.2 + K, 6)
    assert_almost_equal(r["dewpoint"][0], 19.0 + K, 6)


# ---- CAPE / CIN ------------------------------------------------------------------------------
def _base_no_lcl(s, add=0.0):
    prof = no_lcl_profile(s, add=add)
    le = lfc_el_T(prof)
    cc = op.cape_cin_base(col(s["levels"]), col(s["temperatures"]), le["lfc_pressure"],
                          le["el_pressure"], prof["temperature"])
    return le, cc


def test_cape_cin(soundings):                                    # UT:828-854
    _, cc = _base_no_lcl(soundings["test_cape_cin"])
    assert_almost_equal(cc["cape"][0], 75.05354, 2)
    assert_almost_equal(cc["cin"][0], -89.890078, 2)


def test_cape_cin_no_el(soundings):                              # UT:856-881
    _, cc = _base_no_lcl(soundings["test_cape_cin_no_el"])
    assert_almost_equal(cc["cape"][0], 0.08610409, 2)
    assert_almost_equal(cc["cin"][0], -89.8900784, 2)


def test_cape_cin_no_lfc(soundings):                             # UT:883-907
    _, cc = _base_no_lcl(soundings["test_cape_cin_no_lfc"])
    assert_almost_equal(cc["cape"][0], 0.0, 2)
    assert_almost_equal(cc["cin"][0], 0.0, 2)


def test_cape_cin_custom_profile(soundings):                     # UT:1251-1276
    _, cc = _base_no_lcl(soundings["test_cape_cin_custom_profile"], add=5.0)
    assert_almost_equal(cc["cape"][0], 1440.463208696, 2)
    assert_almost_equal(cc["cin"][0], 0.0, 2)


def _sb(s, **kw):
    cc, _ = op.surface_based_cape_cin(col(s["levels"]), col(s["temperatures"]),
                                      col(s["dewpoints"]), ODE, **kw)
    return cc


def _mu(s, **kw):
    cc, prof, ul = op.most_unstable_cape_cin(col(s["levels"]), col(s["temperatures"]),
                                             col(s["dewpoints"]), ODE, **kw)
    return cc


def test_surface_based_cape_cin_mp(soundings):                   # UT:925-938
    cc = _sb(soundings["test_surface_based_cape_cin_mp"], **MP)
    assert_almost_equal(cc["cape"][0], 75.0535446, 2)
    assert_almost_equal(cc["cin"][0], -136.685967, 2)


def test_surface_based_cape_cin(soundings):                      # UT:940-951 (VTC + log: reference-specific)
    cc = _sb(soundings["test_surface_based_cape_cin"])
    assert_almost_equal(cc["cape"][0], 230.1982, 2)
    assert_almost_equal(cc["cin"][0], -58.0673, 2)


def test_sensitive_sounding_cape_mp(soundings):                  # UT:487-493
    cc = _sb(soundings["test_sensitive_sounding_mp"], **MP)
    assert_almost_equal(cc["cape"][0], 0.1115, 3)
    assert_almost_equal(cc["cin"][0], -6.0866, 3)


def test_sensitive_sounding_cape(soundings):                     # UT:525-529
    cc = _sb(soundings["test_sensitive_sounding"])
    assert_almost_equal(cc["cape"][0], 0.5961, 3)
    assert_almost_equal(cc["cin"][0], -5.1399, 3)


def test_profile_with_lcl_in_levels_mp(soundings):               # UT:953-970
    cc = _mu(soundings["test_profile_with_lcl_in_levels_mp"], **MP)
    assert_almost_equal(cc["cape"][0], 75.0535446, 2)
    assert_almost_equal(cc["cin"][0], -136.685967, 2)


def test_profile_with_lcl_in_levels(soundings):                  # UT:972-987
    cc = _mu(soundings["test_profile_with_lcl_in_levels"])
    assert_almost_equal(cc["cape"][0], 230.1982, 2)
    assert_almost_equal(cc["cin"][0], -58.0673, 2)


@pytest.mark.parametrize("name,kw", [("test_profile_with_nans_mp", MP),      # UT:989-1043
                                     ("test_profile_with_nans", {})])        # UT:1045-1095
def test_profile_with_nans(soundings, name, kw):
    s = soundings[name]
    le, cc = _base_no_lcl(s)
    assert np.isnan(le["lfc_pressure"][0])
    assert_almost_equal(cc["cape"][0], 0, 0)
    assert_almost_equal(cc["cin"][0], 0, 0)
    for f in (_sb, _mu):
        r = f(s, **kw)
        assert_almost_equal(r["cape"][0], 0, 0)
        assert_almost_equal(r["cin"][0], 0, 0)


def test_most_unstable_cape_cin_surface_mp(soundings):           # UT:1097-1113
    cc = _mu(soundings["test_most_unstable_cape_cin_surface_mp"], **MP)
    assert_almost_equal(cc["cape"][0], 75.0535446, 2)
    assert_almost_equal(cc["cin"][0], -136.685967, 2)


def test_most_unstable_cape_cin_surface(soundings):              # UT:1115-1129
    cc = _mu(soundings["test_most_unstable_cape_cin_surface"])
    assert_almost_equal(cc["cape"][0], 230.1982, 2)
    assert_almost_equal(cc["cin"][0], -58.0673, 2)


def test_mixed_layer_cape_cin(soundings):                        # UT:1155-1168
    s = soundings["multiple_intersections"]
    cc, _, _ = op.mixed_layer_cape_cin(col(s["levels"]), col(s["temperatures"]),
                                       col(s["dewpoints"]), ODE, **MP)
    assert_almost_equal(cc["cape"][0], 1096.7461, 2)
    assert_almost_equal(cc["cin"][0], -20.6727, 2)


def test_cape_cin_value_error(soundings):                        # UT:1308-1336
    cc = _sb(soundings["test_cape_cin_value_error"], **MP)
    assert_almost_equal(cc["cape"][0], 2007.040698, 3)
    assert_almost_equal(cc["cin"][0], 0.0, 3)


def test_lifted_index(soundings):                                # UT:1353-1386
    s = soundings["test_lifted_index"]
    p, t, td = col(s["pressure"]), col(s["temperature"]), col(s["dewpoint"])
    prof = op.parcel_profile(p, p[0], t[0], td[0], ODE)
    prof["environment_temperature"] = t
    assert_almost_equal(op.lifted_index(prof)[0], -7.9176350, 2)


def test_insert_level(soundings):                                # UT:1388-1411 (defined, not run upstream)
    d = {"pressure": np.array([[1000., 900, 800, 700], [1000., 900, 800, 700]]).T,
         "temperature": np.ones((4, 2))}
    level = {"pressure": np.array([1000., 600.]), "temperature": np.array([1.5, 2.])}
    res = op.insert_level(d, level, "pressure")
    np.testing.assert_array_equal(res["pressure"].T, [[1000, 1000, 900, 800, 700],
                                                      [1000, 900, 800, 700, 600]])
    np.testing.assert_array_equal(res["temperature"].T, [[1, 1.5, 1, 1, 1], [1, 1, 1, 1, 2]])


# ---- whole-array == column-by-column (the oracle is vectorised over columns) ----------------
def test_columns_are_independent(soundings):
    names = ["test_surface_based_cape_cin", "test_sensitive_sounding", "test_lfc_inversion"]
    L = max(soundings[n]["levels"].size for n in names)

    def pad(a):
        return np.concatenate([a, np.full(L - a.size, np.nan)])

    P = np.stack([pad(soundings[n]["levels"]) for n in names], axis=1)
    T = np.stack([pad(soundings[n]["temperatures"]) for n in names], axis=1)
    D = np.stack([pad(soundings[n]["dewpoints"]) for n in names], axis=1)
    o = op.Options(op.MoistLapseODE(), lcl_mode="converged")
    cc, prof = op.surface_based_cape_cin(P, T, D, o)
    for i, n in enumerate(names):
        s = soundings[n]
        c1, p1 = op.surface_based_cape_cin(col(s["levels"]), col(s["temperatures"]),
                                           col(s["dewpoints"]), o)
        assert_almost_equal(cc["cape"][i], c1["cape"][0], 9)
        assert_almost_equal(cc["cin"][i], c1["cin"][0], 9)
        assert_almost_equal(prof["lfc_pressure"][i], p1["lfc_pressure"][0], 9)
        assert_almost_equal(prof["el_pressure"][i], p1["el_pressure"][0], 9)


# ---- derived indices: analytic checks of the restatements ---------------------
def test_derived_index_restatements():
    L, N = 30, 3
    p = np.linspace(1000, 300, L)[:, None] * np.ones((1, N))
    h = 7000.0 * np.log(1000.0 / p)
    t = 290.0 - 6.5e-3 * h                                     # constant lapse rate 6.5 K/km
    td = t - 4.0
    assert_almost_equal(op.lapse_rate(p, t, h), -6.5 * np.ones(N), 9)
    # 600 hPa lies between levels: log-interpolation of a field that is linear in ln p is exact
    assert_almost_equal(op.isobar_temperature(p, t, 600.0), 290.0 - 6.5e-3 * 7000.0 * np.log(1000.0 / 600.0), 9)
    assert_almost_equal(op.freezing_level_height(t, h), (290.0 - 273.15) / 6.5e-3 * np.ones(N), 7)   # PF:2137-2160
    li = np.array([-3.0, 0.0, 2.0])
    t850 = 290.0 - 6.5e-3 * 7000.0 * np.log(1000.0 / 850.0)
    assert_almost_equal(op.deep_convective_index(p, t, td, li), (t850 - 273.15) + (t850 - 4.0 - 273.15) - li, 9)
    u = 0.002 * h; v = -0.001 * h
    ws = op.wind_shear(np.ones(N), np.zeros(N), u, v, h, shear_height=6000)
    assert_almost_equal(ws["shear_u"], 12.0 - 1.0, 9)
    assert_almost_equal(ws["shear_v"], -6.0, 9)
    assert ws["positive_shear"].all()
    assert_almost_equal(op.wet_bulb_temperature_fast(t, td), t - 4.0 / 3.0, 12)                     # PF:364-387


def test_significant_hail_parameter_hand_values():
    """PF:2261-2306 against values worked by hand from the SPC formula the reference cites
    (https://www.spc.noaa.gov/exper/mesoanalysis/help/help_sigh.html)."""
    from oracle import parcel as op
    # all thresholds met: SHIP = 2500 * 12 * 7 * 15 * 20 / 42e6 = 1.5
    one = lambda v: np.array([float(v)])
    s = op.significant_hail_parameter(one(2500), one(0.012), one(-7.0), one(258.15), one(20), one(3000))
    assert abs(s[0] - 1.5) < 1e-12
    # MUCAPE < 1300 scales by MUCAPE/1300; lapse < 5.8 by lapse/5.8; freezing level < 2400 by flh/2400
    s = op.significant_hail_parameter(one(650), one(0.012), one(-5.0), one(258.15), one(20), one(1200))
    base = 650 * 12 * 5.0 * 15 * 20 / 42e6
    assert abs(s[0] - base * 0.5 * (5.0 / 5.8) * 0.5) < 1e-12
    # T500 warmer than -5.5 C is taken as -5.5 C (also when it is NaN); shear / mixing ratio outside their windows -> NaN
    s = op.significant_hail_parameter(one(2500), one(0.012), one(-7.0), one(270.15), one(20), one(3000))
    assert abs(s[0] - 2500 * 12 * 7 * 5.5 * 20 / 42e6) < 1e-12
    s = op.significant_hail_parameter(one(2500), one(0.012), one(-7.0), one(np.nan), one(20), one(3000))
    assert abs(s[0] - 2500 * 12 * 7 * 5.5 * 20 / 42e6) < 1e-12
    for shear, mr in ((6.9, 0.012), (27.1, 0.012), (20, 0.0109), (20, 0.0137)):
        assert np.isnan(op.significant_hail_parameter(one(2500), one(mr), one(-7.0), one(258.15), one(shear), one(3000))[0])


def test_storm_proxies_hand_values():
    """PF:2323-2407: one column per proxy, triggered exactly at its published threshold."""
    from oracle import parcel as op
    base = {"mixed_100_cape": 0.0, "mixed_50_cape": 0.0, "mu_cape": 0.0, "shear_magnitude": 0.0,
            "mixed_100_lifted_index": 5.0, "mixed_100_dci": 0.0, "positive_shear": 1.0, "mixed_50_cin": -100.0,
            "mixed_100_cin": -100.0, "lapse_rate_700_500": -5.0, "mu_mixing_ratio": 0.005, "temp_500": 260.0,
            "freezing_level": 3000.0}

    def run(**kw):
        d = {k: np.array([float(kw.get(k, v))]) for k, v in base.items()}
        return {k: v[0] for k, v in op.storm_proxies(d).items()}

    none = run()
    assert not any(bool(none[k]) for k in none if k.startswith("proxy_"))
    assert run(mixed_100_cape=1000, shear_magnitude=20)["proxy_Craven2004"]
    assert not run(mixed_100_cape=999, shear_magnitude=20)["proxy_Craven2004"]
    assert run(mixed_100_lifted_index=-2.07)["proxy_Kunz2007"] and run(mu_cape=1474)["proxy_Kunz2007"]
    assert run(mixed_100_dci=25.7)["proxy_Kunz2007"] and not run(mixed_100_dci=25.6)["proxy_Kunz2007"]
    assert run(mixed_100_cape=1000, shear_magnitude=10)["proxy_Trapp2007"]
    assert not run(mixed_100_cape=1000, shear_magnitude=10, positive_shear=0)["proxy_Trapp2007"]
    assert run(mixed_100_cape=1000, shear_magnitude=10)["proxy_Marsh2009"]
    assert run(mixed_50_cape=25000 / 10 ** 1.67 + 1e-9, shear_magnitude=10)["proxy_Allen2011"]
    assert run(mixed_50_cape=1000, shear_magnitude=10, mixed_50_cin=-20, lapse_rate_700_500=-7)["proxy_Allen2014"]
    assert not run(mixed_50_cape=1000, shear_magnitude=7.5, mixed_50_cin=-20, lapse_rate_700_500=-7)["proxy_Allen2014"]
    assert run(mixed_100_cape=1001, shear_magnitude=10, mixed_100_cin=-49)["proxy_Eccel2012"]
    assert run(mixed_100_cape=439)["proxy_Mohr2013"] and run(mixed_100_dci=26.4)["proxy_Mohr2013"]
    # negative CAPE is ignored (NaN): no proxy from it
    assert not run(mixed_100_cape=-5, shear_magnitude=-5000)["proxy_Craven2004"]
    assert run(mu_cape=2500, mu_mixing_ratio=0.012, lapse_rate_700_500=-7, temp_500=258.15, shear_magnitude=20)["proxy_SHIP_0.1"]


# ---- wet bulb (Normand's rule), UT:79-104 ------------------------------------------------------------------
def test_wet_bulb_temperature(soundings):                        # UT:79-87
    s = soundings["test_wet_bulb_temperature"]
    val = op.wet_bulb_temperature(col(s["levels"]), col(s["temp"]), col(s["dewp"]), ODE)
    assert_almost_equal(val[0, 0], s["truth"], 5)


def test_wet_bulb_temperature_saturated(soundings):              # UT:89-96
    s = soundings["test_wet_bulb_temperature_saturated"]
    val = op.wet_bulb_temperature(col(s["levels"]), col(s["temp"]), col(s["dewp"]), ODE)
    assert_almost_equal(val[0, 0], 17.6 + K, 7)


def test_wet_bulb_temperature_1d(soundings):                     # UT:98-104 (defined, not run upstream)
    s = soundings["test_wet_bulb_temperature_1d"]
    val = op.wet_bulb_temperature(col(s["pressures"]), col(s["temperatures"]), col(s["dewpoints"]), ODE)
    assert_array_almost_equal(val[:, 0], np.array([21.44487, 16.73673, 12.06554]) + K, 5)
