"""Build and call tests/hostsim (TEST-ONLY host compile of the per-column kernel code)."""

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostsim", "hostsim.cpp")
# XP_HOSTSIM_SANITIZE=1: AddressSanitizer + UndefinedBehaviorSanitizer build of the same per-column code (run the
# CPU tests under LD_PRELOAD=$(gcc -print-file-name=libasan.so); tools/run_sanitizers.sh does it)
SANITIZE = os.environ.get("XP_HOSTSIM_SANITIZE", "") not in ("", "0")
BUILD = os.path.join(HERE, "hostsim", "_build_asan" if SANITIZE else "_build")
LIB = os.path.join(BUILD, "libhostsim.so")
CSRC = os.path.join(os.path.dirname(HERE), "xarray_parcel_b200", "csrc")

SCALARS = ["cape", "cin", "lcl_pressure", "lcl_temperature", "lcl_virtual_temperature",
           "lfc_pressure", "lfc_temperature", "el_pressure", "el_temperature",
           "parcel_pressure", "parcel_temperature", "parcel_dewpoint"]
PROFILE = ["pressure", "temperature", "virtual_temperature", "environment_temperature",
           "environment_virtual_temperature", "environment_dewpoint"]
KINDS = {"sb": 0, "ml": 1, "mu": 2, "explicit": 3}


def build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("xp_math.cuh", "xp_column.cuh", "xp_parcels.cuh", "xp_fast.cuh", "xp_fast_pcol.cuh", "xp_fast6.cuh", "xp_fast7.cuh", "xp_fast_pcol6.cuh", "xp_fast_pcol7.cuh", "xp_layers.cuh", "xp_levels.cuh")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    # -ffp-contract=off: keep host arithmetic un-fused, like NumPy's
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", SRC,
           "-o", LIB]
    if SANITIZE:
        cmd[1:2] = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-fno-sanitize-recover=undefined"]
    subprocess.run(cmd, check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.hostsim_cape_cin.restype = None
    return _lib


def cape_cin(p, t, td, tables, kind="sb", explicit=None, vtc=True, lcl_interp="log",
             pos_cape_neg_cin=True, post_zero_cin=False, metpy_compat=141, ml_depth=100.0,
             mu_depth=300.0, profile=False):
    """Run the host-compiled kernel code on [L, N] float64 arrays.  Returns dict of [N] arrays
    (+ 'profile' dict of [L+1, N] arrays, 'level_shift', 'flags')."""
    t = np.ascontiguousarray(t, dtype=np.float64)
    td = np.ascontiguousarray(td, dtype=np.float64)
    L, N = t.shape
    p = np.ascontiguousarray(p, dtype=np.float64)
    p1d = int(p.ndim == 1)
    out = np.empty((12, N))
    shift = np.empty(N, dtype=np.int32)
    prof = np.empty((6, L + 1, N)) if profile else None
    flags = ctypes.c_uint32(0)
    iopts = (ctypes.c_int * 5)(int(vtc), int(lcl_interp == "log"), int(pos_cape_neg_cin),
                               int(post_zero_cin), int(metpy_compat))
    ex = None
    if explicit is not None:
        ex = np.ascontiguousarray(np.stack([np.broadcast_to(np.asarray(e, dtype=np.float64), (N,))
                                            for e in explicit]))
    idx = np.ascontiguousarray(tables.index_grid, dtype=np.uint16)
    cur = np.ascontiguousarray(tables.curves_asc, dtype=np.float32)

    def ptr(a):
        return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None

    lib().hostsim_cape_cin(ptr(p), ptr(t), ptr(td), ctypes.c_int64(N), ctypes.c_int(L),
                           ctypes.c_int(p1d), ctypes.c_int(KINDS[kind]), ptr(ex), iopts,
                           ctypes.c_double(ml_depth), ctypes.c_double(mu_depth), ptr(idx), ptr(cur),
                           ptr(out), ptr(shift), ptr(prof), ctypes.byref(flags))
    res = {k: out[i] for i, k in enumerate(SCALARS)}
    res["level_shift"] = shift
    res["flags"] = flags.value
    if profile:
        res["profile"] = {k: prof[i] for i, k in enumerate(PROFILE)}
    return res


def set_fast_sweep(version=7, ka_floor=0):
    """Select the sweep (6 / 7) that columns 2, 3 mod 4 run under the default options, and for v7 the stand-in for
    the highest LCL row among the other lanes of the warp."""
    lib().hostsim_set_fast_sweep(int(version), int(ka_floor))


def set_pcol_kind(kind=0, ring_levels=0):
    """hostsim_fast_suite_pcol: 0 = all three kinds at once; 2 / 4 = ONE lifted kind (mixed layer / most unstable) with
    profile rows (the re-based sweep), read through a per-thread ring of ``ring_levels`` levels (0: direct loads)."""
    lib().hostsim_set_pcol_kind(int(kind), int(ring_levels))


def set_pcol_table(on=True):
    """Per-column pressure, default options, no profile rows: read the adiabat family from the two-segment
    shared-memory table (fast::PTabView; what suite_fast_ptab_kernel runs) instead of gathering the curve table."""
    lib().hostsim_set_pcol_table(int(bool(on)))


def ptab_error(tables, n_samples=60000):
    """max |table - reference evaluation| of the virtual temperature of the saturated parcel [K] over random
    (adiabat, pressure) pairs in 100..1100 hPa, 20..100 hPa and 2.5..20 hPa."""
    cur = np.ascontiguousarray(tables.curves_asc, dtype=np.float32)
    out = np.zeros(9)
    lib().hostsim_ptab_error(ctypes.c_void_p(cur.ctypes.data), int(n_samples), ctypes.c_void_p(out.ctypes.data))
    return out


def set_qmode(qmode=0):
    """0: the dewpoint arguments are dewpoints; 141 / 162: they hold specific humidity, converted on load in that
    MetPy form (xp_columns.dewpoint_is_specific_humidity)."""
    lib().hostsim_set_qmode(int(qmode))


def fast_suite(p, t, td, tables, vtc=True, lcl_interp="log", pos_cape_neg_cin=True, post_zero_cin=False,
               metpy_compat=141, ml_depth=100.0, mu_depth=300.0, profile=False):
    """Run the host-compiled float32 fast path on a shared pressure axis p [L] (xp_fast.cuh) or
    per-column pressure p [L, N] (xp_fast_pcol.cuh) and float32 [L, N] T/Td.  Returns ({kind: {field: float32 [N], 'level_shift'}}, redo mask [N]) or None
    if the axis does not qualify."""
    t = np.ascontiguousarray(t, dtype=np.float32)
    td = np.ascontiguousarray(td, dtype=np.float32)
    p = np.ascontiguousarray(p, dtype=np.float32)
    L, N = t.shape
    out = np.empty((3, 12, N), dtype=np.float32)
    shift = np.empty((3, N), dtype=np.int32)
    redo = np.empty(N, dtype=np.uint32)
    iopts = (ctypes.c_int * 5)(int(vtc), int(lcl_interp == "log"), int(pos_cape_neg_cin),
                               int(post_zero_cin), int(metpy_compat))
    idx = np.ascontiguousarray(tables.index_grid, dtype=np.uint16)
    cur = np.ascontiguousarray(tables.curves_asc, dtype=np.float32)

    def ptr(a):
        return a.ctypes.data_as(ctypes.c_void_p)

    l = lib()
    fn = l.hostsim_fast_suite if p.ndim == 1 else l.hostsim_fast_suite_pcol
    fn.restype = ctypes.c_int
    args = [ptr(p), ptr(t), ptr(td), ctypes.c_int64(N), ctypes.c_int(L), iopts, ctypes.c_double(ml_depth),
            ctypes.c_double(mu_depth), ptr(idx), ptr(cur), ptr(out), ptr(shift), ptr(redo)]
    prof = None
    if p.ndim == 2:
        prof = np.full((3, 6, L + 1, N), -12345.0, dtype=np.float32) if profile else None
        args.append(ptr(prof) if prof is not None else None)
    ok = fn(*args)
    if not ok:
        return None
    res = {}
    for q, kind in enumerate(("sb", "ml", "mu")):
        res[kind] = {k: out[q, i] for i, k in enumerate(SCALARS)}
        res[kind]["level_shift"] = shift[q]
        if prof is not None:
            res[kind]["profile"] = {k: prof[q, i] for i, k in enumerate(PROFILE)}
    return res, redo


def mixed_layer(p, fields, depth=100.0, pressure_field=-1):
    """xp_layers.cuh mixed_layer_means on float64 [L, N] fields; returns a list of [N] arrays."""
    x = np.ascontiguousarray(np.stack(fields), dtype=np.float64)
    F, L, N = x.shape
    p = np.ascontiguousarray(p, dtype=np.float64)
    out = np.empty((F, N))
    vp = ctypes.c_void_p
    lib().hostsim_mixed_layer(vp(p.ctypes.data), int(p.ndim == 1), vp(x.ctypes.data), F, ctypes.c_int64(N), L,
                              ctypes.c_double(depth), int(pressure_field), vp(out.ctypes.data))
    return list(out)


def mixed_parcel(p, t, td, depth=100.0):
    """xp_layers.cuh mixed_parcel_full; dict of the six [N] variables of PF:229-289."""
    t = np.ascontiguousarray(t, dtype=np.float64)
    td = np.ascontiguousarray(td, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    L, N = t.shape
    out = np.empty((6, N))
    vp = ctypes.c_void_p
    lib().hostsim_mixed_parcel(vp(p.ctypes.data), int(p.ndim == 1), vp(t.ctypes.data), vp(td.ctypes.data),
                               ctypes.c_int64(N), L, ctypes.c_double(depth), vp(out.ctypes.data))
    names = ["theta", "mixing_ratio", "temperature", "vapour_pressure", "dewpoint", "pressure"]
    return dict(zip(names, out))


def layer_bounds(p, n_columns, depth=100.0, interpolate=True):
    p = np.ascontiguousarray(p, dtype=np.float64)
    L = p.shape[0]
    bottom, top = np.empty(n_columns), np.empty(n_columns)
    vp = ctypes.c_void_p
    lib().hostsim_layer_bounds(vp(p.ctypes.data), int(p.ndim == 1), ctypes.c_int64(n_columns), L,
                               ctypes.c_double(depth), int(bool(interpolate)), vp(bottom.ctypes.data),
                               vp(top.ctypes.data))
    return bottom, top


def _vp(a):
    return ctypes.c_void_p(a.ctypes.data)


def insert_level(c, v, lev_c, lev_v):
    c, v, lev_c, lev_v = [np.ascontiguousarray(a, dtype=np.float64) for a in (c, v, lev_c, lev_v)]
    L, N = c.shape
    out = np.empty((L + 1, N))
    lib().hostsim_insert_level(_vp(c), _vp(v), _vp(lev_c), _vp(lev_v), ctypes.c_int64(N), L, _vp(out))
    return out


def shift_out_nans(ref, v):
    ref, v = [np.ascontiguousarray(a, dtype=np.float64) for a in (ref, v)]
    L, N = ref.shape
    out, shift = np.empty((L, N)), np.empty(N, dtype=np.int32)
    lib().hostsim_shift_out_nans(_vp(ref), _vp(v), ctypes.c_int64(N), L, _vp(out), _vp(shift))
    return out, shift


def trapz(x, v, mask=None, sign=0):
    x, v = [np.ascontiguousarray(a, dtype=np.float64) for a in (x, v)]
    L, N = v.shape
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    out = np.empty(N)
    lib().hostsim_trapz(_vp(x), _vp(v), _vp(m) if m is not None else None, ctypes.c_int64(N), L, int(sign), _vp(out))
    return out


def pressure_order(p):
    p = np.ascontiguousarray(p, dtype=np.float64)
    L, N = p.shape
    return int(lib().hostsim_pressure_order(_vp(p), ctypes.c_int64(N), L))


def find_intersections(x, a, b, log_x=False):
    """xp_levels.cuh interval_crossing over [L, N] arrays -> dict like oracle.parcel.find_intersections."""
    x, a, b = [np.ascontiguousarray(v, dtype=np.float64) for v in (x, a, b)]
    L, N = a.shape
    out = np.empty((3, L - 1, N))
    lib().hostsim_find_intersections(_vp(x), _vp(a), _vp(b), ctypes.c_int64(N), L, int(bool(log_x)), _vp(out))
    ix, iy, sc = out
    with np.errstate(invalid="ignore"):
        return {"all_intersect_x": ix, "all_intersect_y": iy,
                "increasing_x": np.where(sc > 0, ix, np.nan), "increasing_y": np.where(sc > 0, iy, np.nan),
                "decreasing_x": np.where(sc < 0, ix, np.nan), "decreasing_y": np.where(sc < 0, iy, np.nan)}


def interp1d(at, xp, fp):
    at, xp, fp = [np.ascontiguousarray(v, dtype=np.float64) for v in (at, xp, fp)]
    out = np.empty_like(at)
    lib().hostsim_interp1d(_vp(at), _vp(xp), _vp(fp), ctypes.c_int64(at.size), int(xp.size), _vp(out))
    return out


def trap_around_zeros(x, y, log_x=True):
    x, y = [np.ascontiguousarray(v, dtype=np.float64) for v in (x, y)]
    L, N = y.shape
    out = np.empty((3, 2 * L - 1, N))
    lib().hostsim_trap_around_zeros(_vp(x), _vp(y), ctypes.c_int64(N), L, int(bool(log_x)), _vp(out))
    areas = {"area": out[0], "x": out[1], "dx": out[2]}
    areas["x_from"] = areas["x"] - areas["dx"] / 2
    areas["x_to"] = areas["x"] + areas["dx"] / 2
    return areas, np.isnan(out[0][:L])
