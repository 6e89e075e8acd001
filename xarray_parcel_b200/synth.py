"""Seeded synthetic atmospheres for parity tests and the benchmark.

Level-major float32 arrays ``[L, N]`` (level 0 = surface, pressure decreasing with level),
generated with torch so the same code fills host or device memory.  Shapes follow the
configurations in BASELINE.json: hybrid-sigma model levels (L = 70 / 90, per-column
pressure) and the 37 ERA5 pressure levels (one shared 1-D pressure axis).
"""

import math

import torch

ERA5_LEVELS_HPA = [1000, 975, 950, 925, 900, 875, 850, 825, 800, 775, 750, 700, 650, 600, 550,
                   500, 450, 400, 350, 300, 250, 225, 200, 175, 150, 125, 100, 70, 50, 30, 20,
                   10, 7, 5, 3, 2, 1]


def _gen(seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def _uniform(g, n, lo, hi, device):
    return lo + (hi - lo) * torch.rand(n, generator=g, device=device, dtype=torch.float64)


def _thermo_profiles(p, p0, g, device, nan_columns, nan_levels, saturated, inversions,
                     allnan_columns):
    """Temperature / dewpoint on pressures ``p`` [L, N] for surface pressures ``p0`` [N]."""
    L, N = p.shape
    z = 7.5 * torch.log(p0[None, :] / p)                          # km above the surface
    t0 = _uniform(g, N, 270.0, 308.0, device)
    lapse = _uniform(g, N, 5.5, 9.5, device)                      # K / km
    t_strat = _uniform(g, N, 195.0, 225.0, device)
    t = torch.maximum(t0[None, :] - lapse[None, :] * z, t_strat[None, :])
    # low-level capping inversion in a fraction of the columns
    inv = torch.rand(N, generator=g, device=device) < inversions
    amp = _uniform(g, N, 1.0, 6.0, device) * inv
    pc = _uniform(g, N, 780.0, 920.0, device)
    wid = _uniform(g, N, 15.0, 40.0, device)
    t = t + amp[None, :] * torch.exp(-((p - pc[None, :]) / wid[None, :]) ** 2)
    # dewpoint depression: dd0 at the surface growing aloft; a few columns saturated at the surface
    dd0 = _uniform(g, N, 0.5, 25.0, device)
    sat = torch.rand(N, generator=g, device=device) < saturated
    dd0 = torch.where(sat, torch.zeros_like(dd0), dd0)
    dd_top = _uniform(g, N, 10.0, 40.0, device)
    frac = torch.clamp(z / 12.0, 0.0, 1.0)
    wiggle = 3.0 * torch.sin(z * _uniform(g, N, 0.5, 2.0, device)[None, :])
    dd = dd0[None, :] + (dd_top - dd0)[None, :] * frac + wiggle * frac
    # elevated moist layer in a third of the columns (most-unstable parcel above the surface)
    moist = torch.rand(N, generator=g, device=device) < 0.35
    mc = _uniform(g, N, 700.0, 930.0, device)
    mw = _uniform(g, N, 20.0, 60.0, device)
    dry_sfc = _uniform(g, N, 0.0, 12.0, device) * moist
    bump = torch.exp(-((p - mc[None, :]) / mw[None, :]) ** 2) * moist[None, :]
    dd = (dd + dry_sfc[None, :]) * (1.0 - 0.95 * bump)
    dd = torch.clamp(dd, min=0.0)
    td = t - dd
    t = t.to(torch.float32)
    td = torch.minimum(td.to(torch.float32), t)
    # missing data: scattered NaN levels in some columns, and a few all-NaN columns
    if nan_columns > 0:
        col = torch.rand(N, generator=g, device=device) < nan_columns
        lev = torch.rand((L, N), generator=g, device=device) < nan_levels
        lev[0, :] = False
        hole = lev & col[None, :]
        t = torch.where(hole, torch.full_like(t, float("nan")), t)
        td = torch.where(hole, torch.full_like(td, float("nan")), td)
    dead = None
    if allnan_columns > 0:
        dead = torch.rand(N, generator=g, device=device) < allnan_columns
        t = torch.where(dead[None, :], torch.full_like(t, float("nan")), t)
        td = torch.where(dead[None, :], torch.full_like(td, float("nan")), td)
    return t.contiguous(), td.contiguous(), dead


def model_level_columns(n_columns, n_levels=70, seed=1234, device="cpu", nan_columns=0.01,
                        nan_levels=0.1, saturated=0.03, inversions=0.2, allnan_columns=0.002,
                        p_top=20.0):
    """Hybrid-sigma model-level columns: per-column pressure.  Returns (p, T, Td) float32 [L, N]."""
    g = _gen(seed, device)
    N, L = int(n_columns), int(n_levels)
    p0 = _uniform(g, N, 950.0, 1030.0, device)
    s = torch.arange(L, device=device, dtype=torch.float64) / max(L - 1, 1)
    sigma = 1.0 - s ** 1.6                                         # finer spacing near the surface
    p = p_top + (p0[None, :] - p_top) * sigma[:, None]
    p32 = p.to(torch.float32)
    t, td, dead = _thermo_profiles(p32.to(torch.float64), p0, g, device, nan_columns, nan_levels,
                                   saturated, inversions, allnan_columns)
    if dead is not None:
        p32 = torch.where(dead[None, :], torch.full_like(p32, float("nan")), p32)
    return p32.contiguous(), t, td


def era5_columns(n_columns, seed=1234, device="cpu", nan_columns=0.0, nan_levels=0.1,
                 saturated=0.03, inversions=0.2, allnan_columns=0.0):
    """ERA5-shaped columns on the 37 fixed pressure levels.  Returns (p [37], T, Td [37, N]);
    the surface pressure only shapes the profile (data below ground are extrapolated, as in
    ERA5 itself)."""
    g = _gen(seed, device)
    N = int(n_columns)
    p1d = torch.tensor(ERA5_LEVELS_HPA, device=device, dtype=torch.float64)
    p0 = _uniform(g, N, 990.0, 1030.0, device)
    p = p1d[:, None].expand(-1, N)
    t, td, _ = _thermo_profiles(p, p0, g, device, nan_columns, nan_levels, saturated, inversions,
                                allnan_columns)
    return p1d.to(torch.float32).contiguous(), t, td


def algorithmic_bytes_per_column(n_levels, pressure_is_1d, kinds=("sb", "ml", "mu"), elt=4):
    """p/T/Td each read once (2 arrays when pressure is a shared axis), 9 scalar
    outputs per parcel kind, + parcel p/T/Td for ML and MU."""
    b = (2 if pressure_is_1d else 3) * n_levels * elt
    for k in kinds:
        b += 9 * elt + (3 * elt if k in ("ml", "mu") else 0)
    return b


def describe():
    return {"p0_hPa": "U(950,1030) model levels / U(990,1030) ERA5", "t0_K": "U(270,308)",
            "lapse_K_per_km": "U(5.5,9.5)", "stratosphere_K": "U(195,225)",
            "dewpoint_depression_K": "U(0.5,25) at surface growing aloft",
            "scale_height_km": 7.5, "pi": math.pi}
