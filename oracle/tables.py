"""The reference's two moist-adiabat lookup tables (PF:447-523), regenerated.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

The reference builds the tables once with 14 300 calls to
``metpy.calc.moist_lapse`` and caches them to ``./adiabat_lookups/*.nc``
(PF:318-356); the cache is git-ignored upstream and absent from
/root/reference, so the tables must be regenerated.

* ``curves``: adiabat ``i`` (1-based, i = 1..2*len(temperatures)) is the
  pseudo-adiabat through (pressure_levels[0] = 1100 hPa, 173 + 0.01*(i-1) K)
  evaluated on the 2 196 pressures 1100, 1099.5, ..., 2.5 hPa (PF:478-482;
  MetPy's default reference pressure is the first pressure).  Here they are
  integrated with a fixed-step RK4 in ln(p) (global error < 1e-7 K, checked
  against a tight-tolerance SciPy solve in tests/test_oracle_tables.py) instead
  of MetPy's adaptive SciPy solver (rtol 1.5e-8, i.e. a few 1e-6 K): any two
  accurate solvers agree to far better than the table's 0.02 K resolution.
* ``index_grid``: for every (pressure, temperature) cell of the 0.5 hPa x
  0.02 K grid, the number of the *last* adiabat that touches the cell in the
  reference's two marking passes (PF:484-504), 0 where the reference leaves
  NaN.

Shared-artefact convention: the curves are stored as float32
and the index grid as uint16.  Both the oracle and the CUDA path consume the
*same* arrays, so parity between them does not depend on how the table was
generated.  float32 storage perturbs a curve by <= 1.6e-5 K, three orders of
magnitude below the table's own discretisation error (+-0.02 K).
"""

import hashlib
import os

import numpy as np

from . import thermo as th


def round_to(x, to, dp=2):
    """PF:358-362."""
    with np.errstate(invalid="ignore"):
        return np.round(np.round(x / to) * to, dp)


def _nearest_sorted(asc, x):
    """pandas ``Index.get_indexer(method='nearest')`` on a monotonic index, expressed on the
    ascending copy of its values: the closer neighbour, ties -> the larger value, clamped at
    both ends (``.sel(method='nearest')``, PF:554-556)."""
    x = np.asarray(x, dtype=np.float64)
    n = asc.size
    j = np.searchsorted(asc, x, side="left")
    hi = np.clip(j, 0, n - 1)
    lo = np.clip(j - 1, 0, n - 1)
    take_lo = (x - asc[lo]) < (asc[hi] - x)
    return np.where(take_lo, lo, hi)


class AdiabatTables:
    """Index grid + adiabat curves + their coordinate grids."""

    def __init__(self, pressure_desc, temperature, index_grid, curves_asc):
        self.pressure_desc = np.asarray(pressure_desc, dtype=np.float64)   # 1100 ... 2.5
        self.pressure_asc = self.pressure_desc[::-1].copy()                # PF:54 sortby('pressure')
        self.temperature = np.asarray(temperature, dtype=np.float64)       # 173 ... 315.98
        self.index_grid = index_grid          # uint16 [nP (descending pressure), nT]
        self.curves_asc = curves_asc          # float32 [n_adiabats, nP] ascending pressure

    @property
    def n_adiabats(self):
        return self.curves_asc.shape[0]

    def nearest_pressure_index(self, p):
        """Row of ``index_grid`` (descending-pressure order) nearest to p."""
        ia = _nearest_sorted(self.pressure_asc, p)
        return self.pressure_desc.size - 1 - ia

    def nearest_temperature_index(self, t):
        return _nearest_sorted(self.temperature, t)


def default_grids():
    """Default arguments of PF:447-451."""
    pressure_levels = np.round(np.arange(1100, 2, step=-0.5), 1)
    temperatures = np.round(np.arange(173, 316, step=0.02), 2)
    return pressure_levels, temperatures


def integrate_adiabats(pressure_levels, start_temperatures, max_dlnp=0.005):
    """All pseudo-adiabats through (pressure_levels[0], start_temperatures[i]) on
    ``pressure_levels`` (descending), float64 [n, nP].  Classical RK4 on dT/dln(p)."""
    p = np.asarray(pressure_levels, dtype=np.float64)
    t = np.asarray(start_temperatures, dtype=np.float64).copy()
    out = np.empty((t.size, p.size))
    out[:, 0] = t

    def f(x, tt):           # dT/dlnp = p * dT/dp
        pp = np.exp(x)
        return th.moist_lapse_rhs(pp, tt) * pp

    x_nodes = np.log(p)
    for k in range(1, p.size):
        h_tot = x_nodes[k] - x_nodes[k - 1]
        n_sub = max(1, int(np.ceil(abs(h_tot) / max_dlnp)))
        h = h_tot / n_sub
        x = x_nodes[k - 1]
        for _ in range(n_sub):
            k1 = f(x, t)
            k2 = f(x + 0.5 * h, t + 0.5 * h * k1)
            k3 = f(x + 0.5 * h, t + 0.5 * h * k2)
            k4 = f(x + h, t + h * k3)
            t = t + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
            x = x + h
        out[:, k] = t
    return out


def mark_index_grid(pressure_levels, temperatures, curves_desc, pres_step=0.5, temp_step=0.02):
    """PF:484-504 for every adiabat, in the reference's order (later adiabats overwrite).

    ``curves_desc``: float64 [n_adiabats, nP] on ``pressure_levels`` (descending).
    """
    nP, nT = pressure_levels.size, temperatures.size
    grid = np.zeros((nP, nT), dtype=np.uint16)
    t_first = int(np.round(temperatures[0] / temp_step))
    p_rows = np.arange(nP)
    t_cols = np.arange(nT)
    # position of a rounded pressure value in the (descending, regular) pressure grid
    p0 = pressure_levels[0]
    for n in range(curves_desc.shape[0]):
        i = n + 1
        profile_temps = curves_desc[n]
        if np.isnan(profile_temps[0]):
            continue
        # pass 1 (PF:484-489): nearest table temperature at every pressure level
        nearest_temps = round_to(profile_temps, temp_step)
        k = np.round(profile_temps / temp_step).astype(np.int64) - t_first
        ok = (k >= 0) & (k < nT)
        kk = np.where(ok, k, 0)
        ok &= temperatures[kk] == nearest_temps          # == np.isin(nearest_temps, temperatures)
        grid[p_rows[ok], kk[ok]] = i
        # pass 2 (PF:495-504): pressure at which this adiabat has each table temperature
        pres_per_temp = np.interp(x=temperatures, xp=profile_temps[::-1],
                                  fp=pressure_levels[::-1], right=np.nan, left=np.nan)
        pres_per_temp = round_to(pres_per_temp, pres_step)
        with np.errstate(invalid="ignore"):
            r = np.round((p0 - pres_per_temp) / pres_step)
        okp = np.isfinite(r) & (r >= 0) & (r < nP)
        rr = np.where(okp, r, 0).astype(np.int64)
        okp &= pressure_levels[rr] == pres_per_temp      # == np.isin(pres_per_temp, pressure_levels)
        grid[rr[okp], t_cols[okp]] = i
    return grid


def moist_adiabat_lookup(pressure_levels=None, temperatures=None, pres_step=0.5, temp_step=0.02,
                         adiabat_slice=None):
    """PF:447-523.  ``adiabat_slice`` restricts to a range of starting-temperature indices
    (tests only; adiabat numbering stays global)."""
    dp, dt = default_grids()
    pressure_levels = dp if pressure_levels is None else np.asarray(pressure_levels, float)
    temperatures = dt if temperatures is None else np.asarray(temperatures, float)
    # PF:478-482: for T in temperatures, for offset in (0, temp_step/2)
    starts = np.stack([temperatures, temperatures + temp_step / 2], axis=1).reshape(-1)
    curves = integrate_adiabats(pressure_levels, starts)
    if adiabat_slice is not None:
        sel = np.zeros(starts.size, dtype=bool)
        sel[adiabat_slice] = True
        marked = np.where(sel[:, None], curves, np.nan)
    else:
        marked = curves
    grid = mark_index_grid(pressure_levels, temperatures, marked, pres_step, temp_step)
    curves_asc = np.ascontiguousarray(curves[:, ::-1]).astype(np.float32)
    return AdiabatTables(pressure_levels, temperatures, grid, curves_asc)


_CACHE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_cache")
_GENERATOR_VERSION = "rk4-lnp-0.005-v1"


def load_tables(cache=True):
    """``load_moist_adiabat_lookups`` (PF:39-54) for the default grids, cached on disk under
    oracle/_cache (git-ignored) like the reference caches to ./adiabat_lookups."""
    pressure_levels, temperatures = default_grids()
    key = hashlib.sha1((_GENERATOR_VERSION + repr((th.RD, th.EPSILON, th.CP_D, th.LV,
                                                   pressure_levels.size, temperatures.size))
                        ).encode()).hexdigest()[:12]
    fn = os.path.join(_CACHE_DIR, f"adiabat_tables_{key}.npz")
    if cache and os.path.exists(fn):
        z = np.load(fn)
        return AdiabatTables(pressure_levels, temperatures, z["index_grid"], z["curves_asc"])
    tb = moist_adiabat_lookup(pressure_levels, temperatures)
    if cache:
        os.makedirs(_CACHE_DIR, exist_ok=True)
        tmp = fn + f".tmp{os.getpid()}.npz"
        np.savez(tmp, index_grid=tb.index_grid, curves_asc=tb.curves_asc)
        os.replace(tmp, fn)
    return tb
