"""MetPy thermodynamic formulas used by the reference's parcel path, restated.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

The reference (``/root/reference/modules/parcel_functions.py``, "PF") calls
13 ``metpy.calc`` functions and two ``metpy.constants``; MetPy is a PyPI
dependency that is neither vendored under /root/reference nor installed here.
The reference has no lock file; the only version pins are notebook print-outs
(MetPy 1.4.1: parcel_functions_demo.ipynb:86-87; MetPy 1.6.2:
environment_changes_eval.ipynb:105-106).  Both versions use the Bolton (1980)
saturation vapour pressure; the formulas below are MetPy's published ones and
are pinned numerically by the reference's own known-answer tests
(``modules/unit_tests.py``; see ``tests/test_oracle_kat.py``).

Call sites restated (PF line -> function here):
  PF:644  metpy.calc.lcl                                  -> lcl()
  PF:480  metpy.calc.moist_lapse                          -> moist_lapse_ode()
  PF:258, 760  saturation_mixing_ratio                    -> saturation_mixing_ratio()
  PF:698  relative_humidity_from_dewpoint                 -> inside mixing_ratio_from_t_td()
  PF:701  mixing_ratio_from_relative_humidity             -> inside mixing_ratio_from_t_td()
  PF:123  equivalent_potential_temperature                -> equivalent_potential_temperature()
  PF:253  potential_temperature                           -> potential_temperature()
  PF:269  exner_function                                  -> exner_function()
  PF:275  vapor_pressure                                  -> vapor_pressure()
  PF:280  dewpoint                                        -> dewpoint_from_vapor_pressure()
  PF:313  mpconsts.kappa, PF:1361-1382 mpconsts.Rd        -> KAPPA, RD
"""

import numpy as np

# ---- metpy.constants (MetPy 1.x) -------------------------------------------
R_GAS = 8.314462618            # J / (mol K)
MD = 28.96546e-3               # kg / mol, dry air
MW = 18.015268e-3              # kg / mol, water
RD = R_GAS / MD                # 287.04749097718457 J / (kg K)
EPSILON = MW / MD              # 0.6219569100577033
_GAMMA = 1.4                   # dry_air_spec_heat_ratio
CP_D = _GAMMA * RD / (_GAMMA - 1.0)
KAPPA = RD / CP_D              # 2/7 up to rounding
LV = 2.50084e6                 # J / kg
SAT_PRESSURE_0C = 6.112        # hPa
ZERO_DEGC = 273.15
P0 = 1000.0                    # hPa, potential-temperature reference

METPY_COMPAT_DEFAULT = "1.4.1"


def saturation_vapor_pressure(temperature):
    """Bolton (1980) eq. 10; metpy.calc.saturation_vapor_pressure (<= 1.6). hPa, K."""
    t = np.asarray(temperature, dtype=np.float64)
    with np.errstate(all="ignore"):
        return SAT_PRESSURE_0C * np.exp(17.67 * (t - 273.15) / (t - 29.65))


def dewpoint_from_vapor_pressure(e):
    """metpy.calc.dewpoint (Bolton inverse), result converted degC -> K (PF:280-281)."""
    e = np.asarray(e, dtype=np.float64)
    with np.errstate(all="ignore"):
        val = np.log(e / SAT_PRESSURE_0C)
        return 243.5 * val / (17.67 - val) + ZERO_DEGC


def mixing_ratio_from_pressures(partial_press, total_press):
    """metpy.calc.mixing_ratio: eps * e / (p - e) (no clipping in MetPy 1.4.1)."""
    with np.errstate(all="ignore"):
        return EPSILON * partial_press / (total_press - partial_press)


def saturation_mixing_ratio(total_press, temperature):
    """metpy.calc.saturation_mixing_ratio (PF:258, PF:760)."""
    return mixing_ratio_from_pressures(saturation_vapor_pressure(temperature),
                                       np.asarray(total_press, dtype=np.float64))


def vapor_pressure(pressure, mixing_ratio):
    """metpy.calc.vapor_pressure: p * w / (eps + w) (PF:275)."""
    with np.errstate(all="ignore"):
        return pressure * mixing_ratio / (EPSILON + mixing_ratio)


def mixing_ratio_from_t_td(temperature, dewpoint, pressure,
                           metpy_compat=METPY_COMPAT_DEFAULT):
    """PF:684-710 ``mixing_ratio``: RH from dewpoint, then w from RH.

    relative_humidity_from_dewpoint = es(Td) / es(T).
    mixing_ratio_from_relative_humidity:
      MetPy 1.4.1:  w = rh * ws(p, T)
      MetPy >= 1.6: w = eps * ws * rh / (eps + ws * (1 - rh))   (recalled, unverified:
                    the reference's pins were produced with 1.4.1).
    """
    es_t = saturation_vapor_pressure(temperature)
    es_td = saturation_vapor_pressure(dewpoint)
    with np.errstate(all="ignore"):
        rh = es_td / es_t
        ws = mixing_ratio_from_pressures(es_t, np.asarray(pressure, dtype=np.float64))
        if metpy_compat == "1.4.1":
            return rh * ws
        elif metpy_compat == "1.6.2":
            return EPSILON * ws * rh / (EPSILON + ws * (1.0 - rh))
    raise ValueError(f"unknown metpy_compat {metpy_compat!r}")


def virtual_temperature(temperature, mixing_ratio, epsilon=0.608):
    """PF:782-804 (Doswell & Rasmussen 1994 form, NOT MetPy's)."""
    return temperature * (1 + epsilon * mixing_ratio)


def exner_function(pressure):
    """metpy.calc.exner_function with the default 1000 hPa reference (PF:269)."""
    with np.errstate(all="ignore"):
        return (np.asarray(pressure, dtype=np.float64) / P0) ** KAPPA


def potential_temperature(pressure, temperature):
    """metpy.calc.potential_temperature: T / exner(p) (PF:253)."""
    with np.errstate(all="ignore"):
        return temperature / exner_function(pressure)


def dry_lapse(pressure, parcel_temperature, parcel_pressure):
    """PF:291-316: T0 * (p / p0) ** kappa."""
    with np.errstate(all="ignore"):
        return parcel_temperature * (pressure / parcel_pressure) ** KAPPA


def equivalent_potential_temperature(pressure, temperature, dewpoint):
    """metpy.calc.equivalent_potential_temperature, Bolton (1980) eq. 39 (PF:123)."""
    t = np.asarray(temperature, dtype=np.float64)
    td = np.asarray(dewpoint, dtype=np.float64)
    p = np.asarray(pressure, dtype=np.float64)
    with np.errstate(all="ignore"):
        r = saturation_mixing_ratio(p, td)
        e = saturation_vapor_pressure(td)
        t_l = 56 + 1. / (1. / (td - 56) + np.log(t / td) / 800.)
        th_l = potential_temperature(p - e, t) * (t / t_l) ** (0.28 * r)
        return th_l * np.exp(r * (1 + 0.448 * r) * (3036. / t_l - 1.78))


# ---- LCL ---------------------------------------------------------------------
def _lcl_iter(p, p0, w, t):
    """One fixed-point step of metpy.calc.lcl: p <- p0 * (Td(e(p, w)) / T) ** (1/kappa)."""
    with np.errstate(all="ignore"):
        td = dewpoint_from_vapor_pressure(vapor_pressure(p, w))
        return p0 * (td / t) ** (1. / KAPPA)


def lcl(pressure, temperature, dewpoint, mode="converged", max_iters=50, eps=1e-5):
    """metpy.calc.lcl (MetPy <= 1.6: SciPy fixed-point iteration), PF:644.

    mode='scipy'     : literally ``scipy.optimize.fixed_point(..., xtol=1e-5, maxiter=50)``
                       on the whole array, i.e. Steffensen steps with an *array-wide*
                       convergence test, as MetPy does.  The result of one column then
                       depends (at the ~4e-4 hPa level, parcel_functions_demo.ipynb cell 26)
                       on which other columns share its array/chunk.
    mode='converged' : every column iterated to its own fixed point (to rounding).  This
                       is the limit the 'scipy' mode approaches and is chunking-independent;
                       it is what the CUDA path is compared against.

    Returns (lcl_pressure, lcl_temperature) as float64 arrays.  Inputs must not contain
    NaN (the reference substitutes valid values first, PF:627-634).
    """
    p0 = np.asarray(pressure, dtype=np.float64)
    t = np.asarray(temperature, dtype=np.float64)
    td0 = np.asarray(dewpoint, dtype=np.float64)
    p0, t, td0 = np.broadcast_arrays(p0, t, td0)
    w = mixing_ratio_from_pressures(saturation_vapor_pressure(td0), p0)

    nan_mask = np.zeros(p0.shape, dtype=bool)
    if mode == "scipy":
        import scipy.optimize as so

        def it(p, p0_, w_, t_):
            nonlocal nan_mask
            p_new = _lcl_iter(p, p0_, w_, t_)
            nan_mask = nan_mask | np.isnan(p_new)
            return np.where(np.isnan(p_new), p, p_new)

        lcl_p = so.fixed_point(it, p0.copy(), args=(p0, w, t), xtol=eps, maxiter=max_iters)
    elif mode == "converged":
        p = p0.copy()
        for _ in range(80):           # contraction factor ~0.19 per step
            p_new = _lcl_iter(p, p0, w, t)
            bad = np.isnan(p_new)
            nan_mask |= bad
            p_new = np.where(bad, p, p_new)
            if np.array_equal(p_new, p):
                break
            p = p_new
        lcl_p = p
    else:
        raise ValueError(mode)
    lcl_p = np.where(nan_mask, np.nan, lcl_p)
    # np.isclose(lcl_p, pressure): |a - b| <= 1e-8 + 1e-5 * |b|
    with np.errstate(all="ignore"):
        lcl_p = np.where(np.isclose(lcl_p, p0), p0, lcl_p)
    lcl_t = dewpoint_from_vapor_pressure(vapor_pressure(lcl_p, w))
    return lcl_p, lcl_t


# ---- moist adiabat (exact ODE) ------------------------------------------------
def moist_lapse_rhs(p, t):
    """dT/dp of the pseudo-adiabat as in metpy.calc.moist_lapse (also written out in the
    reference's dead file modules/moist_lapse_analytic.py:26-32)."""
    rs = saturation_mixing_ratio(p, t)
    with np.errstate(all="ignore"):
        frac = (RD * t + LV * rs) / (CP_D + (LV * LV * rs * EPSILON / (RD * t * t)))
        return frac / p


def moist_lapse_ode(pressure, temperature, reference_pressure=None, solver="lsoda-tight"):
    """metpy.calc.moist_lapse for ONE parcel: temperatures at ``pressure`` (1-D, any order,
    NaN allowed -> NaN) of the pseudo-adiabat through (reference_pressure, temperature).

    solver='odeint-1.4.1' : ``scipy.integrate.odeint`` at its default tolerances over
        [reference, pressures below it descending] and [reference, pressures above it
        ascending], which is how MetPy 1.4.1 (the version that produced the reference's pins)
        integrates.  Its global error is a few 1e-5 K, which is visible in the most
        sensitive pin (EL of unit_tests.py:588-607, 31 hPa per K).
    solver='lsoda-tight'  : solve_ivp(LSODA, rtol=atol=1e-10): the exact curve to ~1e-8 K.
    """
    from scipy.integrate import odeint, solve_ivp
    p = np.atleast_1d(np.asarray(pressure, dtype=np.float64))
    if reference_pressure is None:
        reference_pressure = p[0]
    pref = float(reference_pressure)
    t0 = float(temperature)
    out = np.full(p.shape, np.nan)
    if not (np.isfinite(pref) and np.isfinite(t0)):
        return out
    valid = np.isfinite(p) & (p > 0)

    def fun(pp, tt):
        return moist_lapse_rhs(pp, tt)

    for sel, direction in ((valid & (p <= pref), -1), (valid & (p > pref), +1)):
        if not sel.any():
            continue
        targets = np.unique(p[sel])
        if direction < 0:
            targets = targets[::-1]
        if solver == "odeint-1.4.1":
            trace = odeint(lambda tt, pp: moist_lapse_rhs(pp, tt), t0,
                           np.append(pref, targets))[1:, 0]
            lut = dict(zip(targets.tolist(), trace.tolist()))
        elif solver == "lsoda-tight":
            if targets[-1] == pref:
                lut = {pref: t0}
            else:
                sol = solve_ivp(fun, (pref, targets[-1]), [t0], t_eval=targets, method="LSODA",
                                rtol=1e-10, atol=1e-10)
                lut = dict(zip(sol.t.tolist(), sol.y[0].tolist()))
        else:
            raise ValueError(solver)
        out[sel] = [lut[v] for v in p[sel].tolist()]
    return out


# ---- specific humidity front end (PF:1889, 1969, 2048-2049) -----------------------------------------
def dewpoint_from_specific_humidity(pressure, temperature, specific_humidity, metpy_compat=METPY_COMPAT_DEFAULT):
    """metpy.calc.dewpoint_from_specific_humidity.

    MetPy 1.4.1 goes through relative humidity (w / ws(p, T), then dewpoint(rh * es(T)));
    MetPy >= 1.6 through the vapour pressure e = p w / (eps + w) (the change noted in
    environment_changes_eval.ipynb:278).  The >= 1.6 form is recalled, not verified.
    """
    with np.errstate(all="ignore"):
        w = specific_humidity / (1 - specific_humidity)          # mixing_ratio_from_specific_humidity
        if str(metpy_compat) in ("1.6.2", "162"):
            return dewpoint_from_vapor_pressure(vapor_pressure(pressure, w))
        rh = w / saturation_mixing_ratio(pressure, temperature)
        return dewpoint_from_vapor_pressure(rh * saturation_vapor_pressure(temperature))
