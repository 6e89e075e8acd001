// xp_layers.cuh -- per-column layer primitives (mixed_layer PF:137-162 on get_layer PF:63-100 and trapz
// PF:164-206; bound_pressure PF:208-227).  Host-compilable for tests/hostsim (see xp_math.cuh).
#pragma once
#include "xp_column.cuh"

namespace xp {

// Column-serial form of get_layer(interpolate=True) + trapz + 1/depth for columns whose pressure decreases with
// the level index (valid_data PF:2308-2321); trailing NaN pressures end the column.  `load(k, x)` fills the NF
// values of level k.  Areas that are NaN are skipped (xarray .sum skips NaN, PF:206).
// `pressure_field` (or -1) names the variable that is the pressure itself: its value at the inserted level is the
// top pressure, not an interpolated one (PF:87).
template <int NF, class PressureAt, class Load>
XP_HD void mixed_layer_means(int L, PressureAt pressure_at, Load load, double depth, double &bottom, double &top,
                             double (&mean)[NF], int pressure_field = -1) {
    bottom = qnan();
    for (int k = 0; k < L; ++k) {                             // PF:80 pressure.max over the column
        const double p = pressure_at(k);
        if (!isnan(p) && !(p <= bottom)) bottom = p;
    }
    top = bottom - depth;                                    // PF:84
    double sum[NF], prev[NF], cur[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) { sum[f] = 0.0; prev[f] = qnan(); cur[f] = qnan(); }
    bool have_prev = false;
    double pp = qnan();
    for (int k = 0; k < L; ++k) {
        const double p = pressure_at(k);
        if (isnan(p)) break;
        load(k, cur);
        if (p >= top) {                                      // inside the layer (PF:97-98)
            if (have_prev) {
                const double dx = fabs(p - pp);              // PF:186
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    const double a = dx * ((prev[f] + cur[f]) / 2);      // PF:188, 198
                    if (!isnan(a)) sum[f] += a;
                }
            }
            have_prev = true; pp = p;
#pragma unroll
            for (int f = 0; f < NF; ++f) prev[f] = cur[f];
        } else {
            // first level above the layer: the inserted top level, interpolated in ln p (PF:85-90).  A level that
            // sits exactly on the top is its own bracket on both sides (PF:1774-1775, 1806).
            if (have_prev) {
                const bool on_top = pp == top;
                const double cb = log(pp), ca = on_top ? cb : log(p), at = log(top);
                const double dx = fabs(top - pp);
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    double x_top = interp_bracket(prev[f], on_top ? prev[f] : cur[f], cb, ca, at);
                    if (f == pressure_field) x_top = top;                // PF:87
                    const double a = dx * ((prev[f] + x_top) / 2);
                    if (!isnan(a)) sum[f] += a;
                }
            }
            break;
        }
    }
    const double pressure_depth = fabs(top - bottom);        // PF:158-159 (the inserted level is the layer minimum)
#pragma unroll
    for (int f = 0; f < NF; ++f) mean[f] = (1. / pressure_depth) * sum[f];   // PF:161
}

// get_layer's bounds: bottom = the column's largest pressure (PF:80); top = bottom - depth (PF:84) or, without
// interpolation, bound_pressure (PF:224-226): the level pressure closest to it, ties -> the larger pressure.
template <class PressureAt>
XP_HD void layer_bounds(int L, PressureAt pressure_at, double depth, bool interpolate, double &bottom, double &top) {
    bottom = qnan();
    for (int k = 0; k < L; ++k) {
        const double p = pressure_at(k);
        if (!isnan(p) && !(p <= bottom)) bottom = p;
    }
    const double bound = bottom - depth;
    top = bound;
    if (!interpolate) {
        double best = qnan();
        top = qnan();
        for (int k = 0; k < L; ++k) {
            const double p = pressure_at(k);
            const double d = fabs(p - bound);
            if (isnan(d)) continue;
            if (!(d >= best)) { best = d; top = p; }
            else if (d == best && p > top) top = p;
        }
    }
}

// mixed_parcel (PF:229-289): out6 = theta, mixing_ratio, temperature, vapour_pressure, dewpoint, pressure.
template <class PressureAt, class TempAt, class DewAt>
XP_HD void mixed_parcel_full(int L, PressureAt pressure_at, TempAt t_at, DewAt td_at, double depth, double (&out6)[6]) {
    auto load = [&](int k, double (&x)[2]) {
        const double p = pressure_at(k);
        x[0] = potential_temperature(p, t_at(k));            // PF:253
        x[1] = sat_mixing_ratio(p, td_at(k));                // PF:258
    };
    double bottom, top, mean[2];
    mixed_layer_means<2>(L, pressure_at, load, depth, bottom, top, mean);
    const double p0 = pressure_at(0);                        // PF:250, 287
    const double e = vapor_pressure(p0, mean[1]);            // PF:275
    out6[0] = mean[0];
    out6[1] = mean[1];
    out6[2] = mean[0] * exner(p0);                           // PF:268-269
    out6[3] = e;
    out6[4] = dewpoint_from_e(e);                            // PF:280-282
    out6[5] = p0;
}

}  // namespace xp
