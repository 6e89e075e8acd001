"""CPU check of the float32 FAST PATH logic (xarray_parcel_b200/csrc/xp_fast.cuh compiled for the
host by tests/hostsim, test-only) against the whole-array oracle.

The fast path may hand any column to the float64 exact kernel (`redo` mask).  For every column it
keeps it must (a) reproduce the oracle's NaN pattern of LFC/EL exactly (i.e. the same crossings
were found), (b) reproduce the integer outputs exactly, (c) meet the north_star tolerances with a
large margin.  The fraction it hands over must stay small.
"""

import numpy as np
import torch
import pytest

import hostsim_util as hs
from oracle import parcel as op
from xarray_parcel_b200 import synth

FIELDS = ["cape", "cin", "lcl_pressure", "lcl_temperature", "lcl_virtual_temperature",
          "lfc_pressure", "lfc_temperature", "el_pressure", "el_temperature"]
# (relative, absolute) regression bounds of the float32 path -- far inside the north_star tolerances
BOUND = {"cape": (2e-5, 0.05), "cin": (2e-5, 0.05), "lcl_pressure": (3e-7, 0), "lcl_temperature": (3e-7, 0),
         "lcl_virtual_temperature": (6e-7, 0), "lfc_pressure": (4e-4, 0), "lfc_temperature": (1e-4, 0),
         "el_pressure": (4e-4, 0), "el_temperature": (1e-4, 0)}

OPTION_SETS = [
    dict(vtc=True, lcl_interp="log", pos_cape_neg_cin=True, compat="1.4.1"),
    dict(vtc=False, lcl_interp="linear", pos_cape_neg_cin=True, compat="1.4.1"),
    dict(vtc=True, lcl_interp="linear", pos_cape_neg_cin=False, compat="1.6.2"),
    dict(vtc=False, lcl_interp="log", pos_cape_neg_cin=False, compat="1.4.1"),
]


def check_fast(res, redo, ora, max_redo, loosen=1.0):
    n = redo.size
    redo = redo | np.where(redo & 8, 4, 0).astype(redo.dtype)     # bit 3: MU == SB, recomputed through SB
    for q, kind in enumerate(("sb", "ml", "mu")):
        keep = ((redo >> q) & 1) == 0
        assert 1.0 - keep.mean() <= max_redo, f"{kind}: {1 - keep.mean():.4f} of the columns handed to the exact path"
        for f in FIELDS:
            a = res[kind][f].astype(np.float64)
            b = ora[f"{kind}_{f}"]
            mis = np.flatnonzero((np.isnan(a) != np.isnan(b)) & keep)
            assert mis.size == 0, f"{kind}_{f}: NaN pattern differs in kept columns {mis[:8]}"
            ok = keep & ~np.isnan(b)
            diff = np.abs(a - b)
            rel, ab = BOUND[f]
            rel, ab = rel * loosen, ab * loosen
            bad = np.flatnonzero(ok & (diff > rel * np.abs(b) + ab))
            assert bad.size == 0, (f"{kind}_{f}: {bad.size}/{n} kept columns outside the float32 bound, e.g. "
                                   f"{[(int(i), float(a[i]), float(b[i])) for i in bad[:4]]}")
        if kind != "sb":
            for f in ("pressure", "temperature", "dewpoint"):
                a = res[kind]["parcel_" + f].astype(np.float64)
                b = ora[f"{kind}_parcel_{f}"]
                ok = keep & ~np.isnan(b)
                assert np.allclose(a[ok], b[ok], rtol=2e-7, atol=0), (kind, f)


@pytest.mark.parametrize("o", OPTION_SETS, ids=lambda o: f"vtc{int(o['vtc'])}-{o['lcl_interp']}-pn{int(o['pos_cape_neg_cin'])}-{o['compat']}")
def test_fast_suite_matches_oracle_era5(oracle_tables, o):
    p, t, td = synth.era5_columns(6000, seed=12, nan_columns=0.01)
    P = np.broadcast_to(p.numpy().astype(np.float64)[:, None], t.shape)
    T, D = t.numpy().astype(np.float64), td.numpy().astype(np.float64)
    opts = op.Options(op.MoistLapseLUT(oracle_tables), lcl_mode="converged", metpy_compat=o["compat"])
    ora = op.suite(P, T, D, opts, virtual_temperature_correction=o["vtc"], lcl_interp=o["lcl_interp"],
                   pos_cape_neg_cin=o["pos_cape_neg_cin"])
    out = hs.fast_suite(p.numpy(), t.numpy(), td.numpy(), oracle_tables, vtc=o["vtc"], lcl_interp=o["lcl_interp"],
                        pos_cape_neg_cin=o["pos_cape_neg_cin"], metpy_compat=141 if o["compat"] == "1.4.1" else 162)
    assert out is not None
    res, redo = out
    check_fast(res, redo, ora, max_redo=0.04)
    # integer outputs: most-unstable level index, number of mixed-layer levels
    mu = op.most_unstable_parcel({"pressure": P, "temperature": T, "dewpoint": D}, depth=300)
    with np.errstate(invalid="ignore"):
        k_mu = (P > mu["pressure"][None, :]).sum(0)
    keep = ((redo >> 2) & 3) == 0
    assert np.array_equal(res["mu"]["level_shift"][keep], k_mu[keep])
    dedup = (redo & 8) != 0                  # "MU == SB" must only be claimed when the MU level is the surface
    assert (k_mu[dedup] == 0).all() and ((redo[dedup] & 1) == 1).all()
    assert (res["ml"]["level_shift"] == 5).all()        # 1000, 975, 950, 925, 900 hPa


@pytest.mark.parametrize("o", OPTION_SETS, ids=lambda o: f"vtc{int(o['vtc'])}-{o['lcl_interp']}-pn{int(o['pos_cape_neg_cin'])}-{o['compat']}")
def test_fast_suite_per_column_pressure(oracle_tables, o):
    """xp_fast_pcol.cuh: model levels with per-column pressure (BASELINE configs[1]/[2] layout)."""
    p, t, td = synth.model_level_columns(5000, 70, seed=21)
    P, T, D = [a.numpy().astype(np.float64) for a in (p, t, td)]
    opts = op.Options(op.MoistLapseLUT(oracle_tables), lcl_mode="converged", metpy_compat=o["compat"])
    ora = op.suite(P, T, D, opts, virtual_temperature_correction=o["vtc"], lcl_interp=o["lcl_interp"],
                   pos_cape_neg_cin=o["pos_cape_neg_cin"])
    res, redo = hs.fast_suite(p.numpy(), t.numpy(), td.numpy(), oracle_tables, vtc=o["vtc"],
                              lcl_interp=o["lcl_interp"], pos_cape_neg_cin=o["pos_cape_neg_cin"],
                              metpy_compat=141 if o["compat"] == "1.4.1" else 162)
    check_fast(res, redo, ora, max_redo=0.07)          # the generator makes 3 % saturated + 1 % NaN columns
    mu = op.most_unstable_parcel({"pressure": P, "temperature": T, "dewpoint": D}, depth=300)
    with np.errstate(invalid="ignore"):
        k_mu = (P > mu["pressure"][None, :]).sum(0)
        k_ml = (P >= (np.nanmax(P, axis=0) - 100.0)[None, :]).sum(0)
    keep = ((redo >> 2) & 3) == 0
    assert np.array_equal(res["mu"]["level_shift"][keep], k_mu[keep])
    keep = ((redo >> 1) & 1) == 0
    assert np.array_equal(res["ml"]["level_shift"][keep], k_ml[keep])


@pytest.mark.parametrize("vtc", [True, False])
def test_fast_profile_rows_per_column_pressure(oracle_tables, vtc):
    """Profile rows (parcel_profile_with_lcl, PF:806-931) written by the per-column fast path, for the
    columns it keeps: every row of every kind against the oracle, NaN padding included."""
    p, t, td = synth.model_level_columns(1500, 50, seed=31, nan_columns=0.05, allnan_columns=0.02)
    P, T, D = [a.numpy().astype(np.float64) for a in (p, t, td)]
    opts = op.Options(op.MoistLapseLUT(oracle_tables), lcl_mode="converged")
    res, redo = hs.fast_suite(p.numpy(), t.numpy(), td.numpy(), oracle_tables, vtc=vtc, profile=True)
    fns = {"sb": op.surface_based_cape_cin, "ml": op.mixed_layer_cape_cin, "mu": op.most_unstable_cape_cin}
    L = T.shape[0]
    seen_rows_ok = [0]
    for q, kind in enumerate(("sb", "ml", "mu")):
        prof = fns[kind](P, T, D, opts, virtual_temperature_correction=vtc)[1]
        keep = ((redo >> q) & 1) == 0
        if kind == "mu":
            keep &= (redo & 8) == 0
        # columns handed over for a crossing decision alone keep their float32 rows (kRedoRowsOk): check them too
        rows_ok = ((redo >> (4 + q)) & 1) == 1
        assert not (rows_ok & keep).any() and rows_ok.sum() <= (~keep).sum()
        seen_rows_ok[0] += int(rows_ok.sum())
        keep = keep | rows_ok
        n = prof["pressure"].shape[0]            # the oracle trims levels that are NaN in every column
        for k in hs.PROFILE:
            a = res[kind]["profile"][k].astype(np.float64)[:, keep]
            b = np.full((L + 1, keep.sum()), np.nan)
            b[:n] = prof[k][:, keep]
            assert not (a == -12345.0).any(), f"{kind} {k}: rows left unwritten"
            if b.shape[1] == 0:
                continue
            assert np.array_equal(np.isnan(a), np.isnan(b)), f"{kind} {k}: NaN pattern differs"
            ok = ~np.isnan(b)
            assert np.allclose(a[ok], b[ok], rtol=3e-6, atol=0), (kind, k, np.abs(a[ok] - b[ok]).max())
    assert seen_rows_ok[0] > 0                   # the case is exercised


def test_ptab_accuracy(oracle_tables):
    """The adiabat family on the two-segment shared-memory grid (fast::PTabView) against the reference's evaluation of
    one adiabat (np.interp on the 0.5 hPa nodes + saturation mixing ratio + virtual temperature): inside the margins
    the kernels decide with (kPTabMarginA 8e-4 / B 5e-3 / Top 0.1 K, half of each for the table itself)."""
    err = hs.ptab_error(oracle_tables, 300000)
    assert err[0] < 4e-4 and err[1] < 2.5e-3 and err[2] < 0.06, err


def test_per_column_pressure_through_the_shared_memory_table(oracle_tables):
    """suite_fast_ptab_kernel's column code (default options, no profile rows): the moist adiabat comes from the
    shared-memory table instead of two gathers per (level, parcel).  Kept columns against the oracle, and no more
    columns handed over than with the gathers (+1 % for the wider margins)."""
    p, t, td = synth.model_level_columns(5000, 70, seed=17)
    P, T, D = [a.numpy().astype(np.float64) for a in (p, t, td)]
    ora = op.suite(P, T, D, op.Options(op.MoistLapseLUT(oracle_tables), lcl_mode="converged"))
    gather, redo_g = hs.fast_suite(p.numpy(), t.numpy(), td.numpy(), oracle_tables)
    hs.set_pcol_table(True)
    try:
        res, redo = hs.fast_suite(p.numpy(), t.numpy(), td.numpy(), oracle_tables)
    finally:
        hs.set_pcol_table(False)
    check_fast(res, redo, ora, max_redo=0.08)
    assert (redo != 0).mean() < (redo_g != 0).mean() + 0.01, ((redo != 0).mean(), (redo_g != 0).mean())
    for kind in ("ml", "mu"):
        assert np.array_equal(res[kind]["level_shift"], gather[kind]["level_shift"])
    # the mixed / above phase boundary is the warp's (here: a stand-in for the other lanes): results must not move
    hs.set_pcol_table(True)
    try:
        for floor in (25, 80):
            hs.set_fast_sweep(7, floor)
            res2, redo2 = hs.fast_suite(p.numpy(), t.numpy(), td.numpy(), oracle_tables)
            assert np.array_equal(redo, redo2), floor
            keep = redo == 0
            for kind in ("sb", "ml", "mu"):
                for f in FIELDS:
                    assert np.array_equal(res[kind][f][keep].view(np.int32), res2[kind][f][keep].view(np.int32)), (floor, kind, f)
    finally:
        hs.set_pcol_table(False)
        hs.set_fast_sweep(7, 0)


@pytest.mark.parametrize("kind,code", [("ml", 2), ("mu", 4)])
def test_rebased_profile_sweep_through_the_ring(oracle_tables, kind, code):
    """ONE lifted kind with profile rows (BASELINE configs[4]): the sweep is re-based per lane (row r at iteration
    r) and reads its levels through a per-thread ring that every lane of a warp fills level-synchronously.  The
    rows must be the oracle's and must not depend on the ring (capacity 48 / too small to be used / none) nor on how
    far the other lanes of the warp lag (the host stand-in for the warp maximum)."""
    p, t, td = synth.model_level_columns(400, 90, seed=21)
    P, T, D = [x.numpy().astype(np.float64) for x in (p, t, td)]
    opts = op.Options(op.MoistLapseLUT(oracle_tables), lcl_mode="converged")
    fn = {"ml": op.mixed_layer_cape_cin, "mu": op.most_unstable_cape_cin}[kind]
    prof = fn(P, T, D, opts)[1]
    n = prof["pressure"].shape[0]
    runs = {}
    try:
        for ring, floor in ((0, 0), (48, 0), (48, 30), (8, 0), (64, 45)):
            hs.set_pcol_kind(code, ring)
            hs.set_fast_sweep(7, floor)
            runs[(ring, floor)] = hs.fast_suite(p.numpy(), t.numpy(), td.numpy(), oracle_tables, profile=True)
    finally:
        hs.set_pcol_kind(0, 0)
        hs.set_fast_sweep(7, 0)
    res0, redo0 = runs[(0, 0)]
    keep = (redo0 & 7) == 0
    assert keep.mean() > 0.8
    for k in ("pressure", "temperature", "virtual_temperature", "environment_temperature",
              "environment_virtual_temperature", "environment_dewpoint"):
        a = res0[kind]["profile"][k].astype(np.float64)[:, keep]
        b = prof[k][:, keep]
        assert np.array_equal(np.isnan(a[:n]), np.isnan(b)), k
        ok = ~np.isnan(b)
        assert np.allclose(a[:n][ok], b[ok], rtol=3e-6, atol=0), k
    for key, (res, redo) in runs.items():
        assert np.array_equal(redo, redo0), key
        for k, v in res0[kind]["profile"].items():
            w = res[kind]["profile"][k]
            assert np.array_equal(v.view(np.int32), w.view(np.int32)), (key, k)
        for f in ("cape", "cin", "lfc_pressure", "el_pressure"):
            assert np.array_equal(res0[kind][f].view(np.int32), res[kind][f].view(np.int32)), (key, f)


def test_fast_suite_per_column_pressure_depths_and_90_levels(oracle_tables):
    p, t, td = synth.model_level_columns(3000, 90, seed=22, nan_columns=0, allnan_columns=0, saturated=0)
    P, T, D = [a.numpy().astype(np.float64) for a in (p, t, td)]
    opts = op.Options(op.MoistLapseLUT(oracle_tables), lcl_mode="converged")
    ora = op.suite(P, T, D, opts, ml_depth=60.0, mu_depth=400.0)
    res, redo = hs.fast_suite(p.numpy(), t.numpy(), td.numpy(), oracle_tables, ml_depth=60.0, mu_depth=400.0)
    check_fast(res, redo, ora, max_redo=0.03)


def test_fast_suite_other_axes(oracle_tables):
    """A shared axis that is not ERA5's: uneven levels, mixed-layer top between levels, depth options."""
    rng = np.random.default_rng(3)
    p = np.sort(np.concatenate([[1013.0], rng.uniform(120, 1010, 30), [80.0, 40.0, 12.0, 4.0, 1.5]]))[::-1]
    p = p.astype(np.float32)
    N = 3000
    z = 7.5 * np.log(1013.0 / p.astype(np.float64))[:, None]
    t0 = rng.uniform(275, 305, N); lapse = rng.uniform(5.5, 9.0, N)
    T = np.maximum(t0[None, :] - lapse[None, :] * z, rng.uniform(200, 220, N)[None, :]).astype(np.float32)
    D = (T - (rng.uniform(1, 15, N)[None, :] + 1.5 * z)).astype(np.float32)
    P = np.broadcast_to(p.astype(np.float64)[:, None], T.shape)
    opts = op.Options(op.MoistLapseLUT(oracle_tables), lcl_mode="converged")
    ora = op.suite(P, T.astype(np.float64), D.astype(np.float64), opts, ml_depth=75.0, mu_depth=250.0)
    res, redo = hs.fast_suite(p, T, D, oracle_tables, ml_depth=75.0, mu_depth=250.0)
    check_fast(res, redo, ora, max_redo=0.04)


def test_fast_path_refuses_bad_axes(oracle_tables):
    t = np.full((5, 4), 280.0, dtype=np.float32)
    for p in ([1000, 900, 900, 700, 500], [900, 1000, 800, 700, 600], [1200, 1000, 800, 700, 600],
              [1000, np.nan, 800, 700, 600]):
        assert hs.fast_suite(np.array(p, dtype=np.float32), t, t - 5, oracle_tables) is None


@pytest.mark.parametrize("which", [0, 1], ids=["lcl_fast", "lcl_fast6"])
def test_fast_lcl_solvers_reach_the_converged_fixed_point(which):
    """float32 Newton + one float64 step (xp_fast.cuh / xp_fast6.cuh) against the converged fixed point of
    metpy.calc.lcl (oracle, PF:644): the agreement must be far below the 0.5 hPa x 0.02 K table cell, so that
    the cell of the LCL is the reference's except within ~1e-8 of a cell edge."""
    import ctypes
    from oracle import thermo as th
    rng = np.random.default_rng(11)
    n = 200_000
    p = rng.uniform(600.0, 1050.0, n)
    t = rng.uniform(240.0, 315.0, n)
    td = t - np.concatenate([rng.uniform(2e-3, 1.0, n // 4), rng.uniform(1.0, 45.0, n - n // 4)])
    lp, lt = np.empty(n), np.empty(n)
    fn = hs.lib().hostsim_lcl_fast
    fn.restype = None
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    fn(ptr(p), ptr(t), ptr(td), ctypes.c_int64(n), ctypes.c_int(which), ptr(lp), ptr(lt))
    rp, rt = th.lcl(p, t, td, mode="converged")
    assert np.abs(lp / rp - 1).max() < 2e-11
    assert np.abs(lt - rt).max() < 2e-9


def test_branch_free_log_exp_accuracy():
    """log64_fast / exp64_fast (xp_fast6.cuh) over the argument ranges of the path: the LCL solve needs ~1e-12."""
    import ctypes
    rng = np.random.default_rng(5)
    fn = hs.lib().hostsim_fast_math64
    fn.restype = None
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    x = np.concatenate([rng.uniform(0.02, 1.0, 200_000), rng.uniform(1e-3, 50.0, 100_000),
                        np.array([1.0, 0.5, 2.0, np.sqrt(2.0), np.nextafter(np.sqrt(2.0), 2), 0.7071067811865476])])
    y = np.empty_like(x)
    fn(ptr(x), ctypes.c_int64(x.size), ctypes.c_int(0), ptr(y))
    ref = np.log(x)
    assert np.abs(y - ref).max() <= 4e-16 + 4e-16 * np.abs(ref).max()
    assert np.abs((y - ref)[np.abs(ref) > 1e-3] / ref[np.abs(ref) > 1e-3]).max() < 2e-15
    x = np.concatenate([rng.uniform(-45.0, 8.0, 300_000), np.array([0.0, -0.3465, 0.3466, 1.0, -1.0])])
    y = np.empty_like(x)
    fn(ptr(x), ctypes.c_int64(x.size), ctypes.c_int(1), ptr(y))
    assert np.abs(y / np.exp(x) - 1).max() < 1e-15
    # pow_kappa64: x^(2/7) by a float32 estimate + one Newton step on y^7 = x^2 (the potential-temperature factor
    # (1000/p)^kappa of the mixed-layer pre-pass): 3 eps^2 ~ 3e-13 relative
    x = np.concatenate([rng.uniform(0.9, 400.0, 300_000), np.array([1.0, 1000.0 / 1100.0, 1000.0 / 2.5])])
    y = np.empty_like(x)
    fn(ptr(x), ctypes.c_int64(x.size), ctypes.c_int(2), ptr(y))
    assert np.abs(y / x ** (2.0 / 7.0) - 1).max() < 1e-12


def test_fast_suite_warm_stratopause(oracle_tables):
    """Top levels warm enough for es(T) to exceed the pressure (real stratopause temperatures at 3-5 hPa): the
    mixing ratio eps es(Td)/(p - es(T)) turns negative there (PF:684-710 as it is), the environment virtual
    temperature drops BELOW the temperature, and the early-termination bound of the v6 sweep (coldest
    temperature aloft) must not be trusted -- kept columns still reproduce the oracle."""
    p, t, td = synth.era5_columns(4000, seed=77)
    t = t.clone(); td = td.clone()
    rng = np.random.default_rng(8)
    warm = torch.from_numpy(rng.uniform(262.0, 285.0, (4, t.shape[1])).astype(np.float32))
    t[-6:-2] = warm                                     # 10, 7, 5, 3 hPa (2 and 1 hPa are outside the tables)
    td[-6:-2] = torch.minimum(td[-6:-2], t[-6:-2] - 60.0)
    # every other column: nearly saturated at 3 hPa with es(T) just above the pressure -> a large negative mixing
    # ratio, environment virtual temperature far below anything the parcel has: a crossing at the very top
    t[-3, ::2] = torch.from_numpy(rng.uniform(266.8, 268.5, t[-3, ::2].shape).astype(np.float32))
    td[-3, ::2] = t[-3, ::2] - 0.5
    P = np.broadcast_to(p.numpy().astype(np.float64)[:, None], t.shape)
    T, D = t.numpy().astype(np.float64), td.numpy().astype(np.float64)
    opts = op.Options(op.MoistLapseLUT(oracle_tables), lcl_mode="converged")
    ora = op.suite(P, T, D, opts)
    res, redo = hs.fast_suite(p.numpy(), t.numpy(), td.numpy(), oracle_tables)
    # columns with es(T) within 5 % of p go to the exact path; in the kept ones the negative mixing ratio aloft makes
    # |CIN| ~ 1e5 J/kg: float32 sums of such terms get 5 x the usual regression bound (still 1e-4 relative)
    check_fast(res, redo, ora, max_redo=0.25, loosen=5.0)
    assert (redo == 0).mean() > 0.7
