// xp_levels.cuh -- per-column level primitives of the parcel path on their own: insert_level (PF:933-990),
// shift_out_nans (PF:1699-1720), trapz (PF:164-206) and the pressure-order check of valid_data (PF:2308-2321).
// Host-compilable for tests/hostsim (see xp_math.cuh).
#pragma once
#include "xp_column.cuh"

namespace xp {

constexpr uint32_t kFlagPressureNotDecreasing = 4u;         // XP_FLAG_PRESSURE_NOT_DECREASING (PF:2320)
constexpr uint32_t kFlagPressureOrderChecked = 8u;          // XP_FLAG_PRESSURE_ORDER_CHECKED
constexpr double kInsertFill = -999.0;                       // insert_level's fill_value default (PF:934)

// insert_level (PF:933-990), one output position j of 0..L for one variable.  c_j / c_jm: the coordinate at levels
// j and j-1 (ignored where the level does not exist), v_j / v_jm: the variable there; lev_c / lev_v: coordinate
// and value of the new level.  Value-based split: coordinates >= the new one stay in place ("below", PF:965),
// coordinates < it move up one (PF:966-970), what is left is the new level (PF:985); NaN coordinates travel as the
// fill value (with every variable of that level, PF:963) and come back as NaN (PF:988).
XP_HD double insert_level_value(int j, int L, double c_j, double c_jm, double v_j, double v_jm, double lev_c,
                                double lev_v) {
    const bool nan_j = j < L && isnan(c_j), nan_jm = j >= 1 && isnan(c_jm);
    const double cj = nan_j ? kInsertFill : c_j, cjm = nan_jm ? kInsertFill : c_jm;
    const bool below = j < L && cj >= lev_c;
    const bool above = j >= 1 && cjm < lev_c;
    double val;
    if (below) val = nan_j ? kInsertFill : v_j;
    else if (above) val = nan_jm ? kInsertFill : v_jm;
    else val = lev_v;
    return val == kInsertFill ? qnan() : val;
}

// shift_out_nans (PF:1714-1718): the number of leading NaN levels of the reference variable (L if all are NaN:
// the reference shifts L times and the column stays NaN).
template <class RefAt>
XP_HD int leading_nans(int L, RefAt ref_at) {
    int k0 = 0;
    while (k0 < L && isnan(ref_at(k0))) ++k0;
    return k0;
}

// trapz (PF:164-206) of one variable: sum over k of |x[k+1] - x[k]| * (v[k] + v[k+1]) / 2, intervals labelled by
// their lower level; `use(k)` is the mask (PF:193-196); sign > 0 keeps positive areas only, sign < 0 negative ones
// (PF:200-204); NaN areas are skipped (xarray .sum, PF:206).
template <class XAt, class VAt, class Use>
XP_HD double trapz_column(int L, XAt x_at, VAt v_at, Use use, int sign) {
    double sum = 0.0;
    for (int k = 0; k + 1 < L; ++k) {
        if (!use(k)) continue;
        const double dx = fabs(x_at(k + 1) - x_at(k));          // PF:186
        const double area = dx * ((v_at(k) + v_at(k + 1)) / 2);  // PF:188, 198
        if (isnan(area)) continue;
        if (sign > 0 && !(area > 0)) continue;
        if (sign < 0 && !(area < 0)) continue;
        sum += area;
    }
    return sum;
}

// find_intersections (PF:992-1064), one interval between two neighbouring levels.  x0 / x1 are the coordinates
// (already ln x when log_x), a / b the two curves.  A sign change of a - b across the interval (NaN signs count as a
// change, PF:1022) gives the crossing (ix, iy) by linear interpolation (PF:1046, 1050) and sign(a1 - b1) tells its
// direction (PF:1030); otherwise everything is NaN.
XP_HD void interval_crossing(double x0, double x1, double a0, double a1, double b0, double b1, bool log_x,
                             double &ix, double &iy, double &sign_change) {
    const double diff = sign_of(a1 - b1) - sign_of(a0 - b0);  // PF:1019
    if (diff == 0) { ix = iy = sign_change = qnan(); return; }
    sign_change = sign_of(a1 - b1);
    const double dy0 = a0 - b0, dy1 = a1 - b1;
    ix = (dy1 * x0 - dy0 * x1) / (dy1 - dy0);
    iy = ((ix - x0) / (x1 - x0)) * (a1 - a0) + a0;
    if (log_x) ix = exp(ix);                                  // PF:1053
}

// trap_around_zeros (PF:1200-1289, start = 0), one interval between two neighbouring levels: where y crosses zero
// (find_intersections against 0, PF:1225) the two triangles next to the zero -- "before" belongs to the lower
// level (shift_x = 1, PF:1272), "after" to the upper one (PF:1273); v[0..2] = area, x (centre), dx.  x0 / x1 are the
// raw coordinates; with log_x everything is in ln x, the zero taking the reference's exp / log round trip (PF:1053, 1232).
XP_HD void zero_half_areas(double x0, double x1, double y0, double y1, bool log_x, double (&before)[3],
                           double (&after)[3]) {
    const double c0 = log_x ? log(x0) : x0, c1 = log_x ? log(x1) : x1;
    double ix, iy, sc;
    interval_crossing(c0, c1, y0, y1, 0.0, 0.0, log_x, ix, iy, sc);
    if (isnan(iy)) {                                         // PF:1237, 1242-1244: no zero in this interval
        for (int q = 0; q < 3; ++q) before[q] = after[q] = qnan();
        return;
    }
    const double zx = log_x ? log(ix) : ix;
    const double dxb = c0 - zx, dxa = c1 - zx;               // PF:1258
    before[0] = (y0 / 2) * fabs(dxb); before[1] = c0 - dxb / 2; before[2] = fabs(dxb);   // PF:1261-1263
    after[0] = (y1 / 2) * fabs(dxa); after[1] = c1 - dxa / 2; after[2] = fabs(dxa);
}

// interp1d_numba (PF:23-37) = numpy.interp for one point: xp increasing, values outside take the end values, an
// exact hit returns the node value, NaN in gives NaN out; the NaN fall-backs of numpy's arr_interp are kept.
template <class XpAt, class FpAt>
XP_HD double interp1d_point(double x, int n, XpAt xp_at, FpAt fp_at) {
    if (isnan(x)) return x;
    if (x < xp_at(0)) return fp_at(0);
    if (x > xp_at(n - 1)) return fp_at(n - 1);
    int lo = 0, hi = n - 1;                                  // invariant: xp[lo] <= x, and x < xp[hi] or hi == n - 1
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= xp_at(mid)) lo = mid; else hi = mid;
    }
    int j = lo;
    if (x >= xp_at(hi)) j = hi;                              // only when x == xp[n - 1]
    const double xj = xp_at(j), fj = fp_at(j);
    if (j == n - 1 || xj == x) return fj;
    const double xj1 = xp_at(j + 1), fj1 = fp_at(j + 1);
    const double slope = (fj1 - fj) / (xj1 - xj);
    double r = slope * (x - xj) + fj;
    if (isnan(r)) {
        r = slope * (x - xj1) + fj1;
        if (isnan(r) && fj == fj1) r = fj;
    }
    return r;
}

// valid_data (PF:2320): pressure.diff(vert_dim).max() < 0.  Returns bit 0 = a difference >= 0 exists,
// bit 1 = a non-NaN difference exists (the reference's max skips NaN; with no valid difference it fails).
template <class PressureAt>
XP_HD int pressure_order(int L, PressureAt pressure_at) {
    int r = 0;
    double prev = L > 0 ? pressure_at(0) : qnan();
    for (int k = 1; k < L; ++k) {
        const double p = pressure_at(k);
        const double d = p - prev;
        if (!isnan(d)) { r |= 2; if (!(d < 0)) r |= 1; }
        prev = p;
    }
    return r;
}

}  // namespace xp
