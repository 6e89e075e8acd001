// xp_column.cuh -- one atmospheric column, one parcel: the fused lifting sweep.
//
// Column-serial form of the reference's whole-array pipeline
//   cape_cin (PF:1394) = parcel_profile_with_lcl (PF:806) -> lfc_el (PF:1066) -> cape_cin_base (PF:1291)
// A single upward sweep over the (L+1)-level profile (input levels + inserted LCL level) feeds
// a small state machine that finds the crossings (find_intersections PF:992-1064), selects
// LFC/EL (PF:1127-1185) and accumulates the log-pressure trapezoids and zero-crossing
// triangles (trapz PF:164-206, trap_around_zeros PF:1200-1289) as running sums that are
// snapshotted at every LFC/EL candidate, so CAPE/CIN follow from differences at the end.
//
// Contract on inputs (the reference's valid_data, PF:2308-2321): pressure strictly
// decreasing with level wherever it is finite; NaN pressures only as trailing levels or
// whole columns.  Temperature/dewpoint may be NaN anywhere.
#pragma once
#include "xp_math.cuh"

namespace xp {

constexpr int kNP = 2196;       // table pressures 2.5 .. 1100 hPa step 0.5 (ascending)
constexpr int kNT = 7150;       // table temperatures 173 .. 315.98 K step 0.02
constexpr int kNAdiabats = 14300;

struct Tables {
    const uint16_t *index_grid;   // [kNP descending pressure][kNT], 0 = none
    const float *curves;          // [kNAdiabats][kNP] ascending pressure
};

struct Opts {
    int vtc, log_interp, pos_neg, post_zero, compat;
    int exact_only;      // 1: never take the float32 fast path
    int vote_mask;       // fast path: early-termination vote every (vote_mask + 1) iterations (tuning)
    double ml_depth, mu_depth;
    int qmode = 0;       // fast paths: != 0 (141 / 162 = compat): the dewpoint array holds specific humidity (ColsArg::qmode)
};

// ---- moist-adiabat lookup (PF:525-607) -------------------------------------------------
XP_HD double table_pressure(int j) { return 2.5 + 0.5 * j; }                 // exact
XP_HD double table_temperature(int k) { return (double)(17300 + 2 * k) / 100.0; }
// np.round(np.arange(173, 316, .02), 2)[k]: the double nearest to the 2-decimal value.

// pandas nearest on a monotonic index (xarray .sel(method='nearest'), PF:554-556): the closer
// neighbour, ties to the larger value, clamped at both ends.  Returns the adiabat number
// (1-based) or 0.
XP_HD int adiabat_lookup(const Tables &tb, double p, double t) {
    if (isnan(p) || isnan(t)) return 0;
    // pressure: j = #nodes < p  (nodes are exact multiples of 0.5)
    double xs = (p - 2.5) * 2.0;
    int jp = (xs <= 0.0) ? 0 : ((xs >= (double)kNP) ? kNP : (int)ceil(xs));
    int hi = min(max(jp, 0), kNP - 1), lo = min(max(jp - 1, 0), kNP - 1);
    int ip = ((p - table_pressure(lo)) < (table_pressure(hi) - p)) ? lo : hi;
    // temperature: j = #nodes < t
    double ts = (t - 173.0) * 50.0;
    int jt = (ts <= 0.0) ? 0 : ((ts >= (double)kNT) ? kNT : (int)ceil(ts));
    while (jt > 0 && table_temperature(jt - 1) >= t) --jt;
    while (jt < kNT && table_temperature(jt) < t) ++jt;
    hi = min(max(jt, 0), kNT - 1);
    lo = min(max(jt - 1, 0), kNT - 1);
    int it = ((t - table_temperature(lo)) < (table_temperature(hi) - t)) ? lo : hi;
    int row = kNP - 1 - ip;     // index grid rows are in descending pressure order
    return (int)XP_LDG(tb.index_grid + (size_t)row * kNT + it);
}

// np.interp(p, P_ascending, curve) with no extrapolation (PF:585-600).
XP_HD double adiabat_temperature(const float *__restrict__ curve, double p) {
    if (!(p >= 2.5) || !(p <= 1100.0)) return qnan();     // also NaN pressure
    int j = (int)floor((p - 2.5) * 2.0);
    j = min(j, kNP - 1);
    double f0 = (double)XP_LDG(curve + j);
    double xj = table_pressure(j);
    if (j == kNP - 1 || xj == p) return f0;
    double f1 = (double)XP_LDG(curve + j + 1);
    double slope = (f1 - f0) / (table_pressure(j + 1) - xj);
    return slope * (p - xj) + f0;
}

// ---- per-parcel results -------------------------------------------------------------------
struct ParcelResult {
    double cape, cin;
    double lcl_p, lcl_t, lcl_tv;
    double lfc_p, lfc_t, el_p, el_t;
    uint32_t flags;
};

// A profile row of parcel_profile_with_lcl (PF:806-931).
struct ProfileRow {
    double p, t, tv, env_t, env_tv, env_td;
};

// Sweep state.  `A` = parcel curve, `B` = environment curve (virtual temperatures with the
// virtual temperature correction, real temperatures without; PF:1436-1470).
struct Sweep {
    // LCL and options
    double lcl_p, lcl_targ;      // lcl_targ: the lcl_temperature argument of lfc_el (PF:1442/1461)
    int pos_neg;
    // previous profile level
    int j;                       // index of the next level to be emitted
    double xp_, pp_, ap_, bp_;   // ln p, p, A, B of the previous level
    bool skip_first;             // A_0 == B_0: ignore the first interval for LFC (PF:1117-1120)
    // running area sums (positive / negative parts) in ln-p units
    double s_pos, s_neg;
    // snapshots
    bool lcl_seen;
    double lcl_pos, lcl_neg;
    double lfc_p, lfc_t, lfc_pos, lfc_neg;    // max-pressure increasing crossing above the LCL
    double el_p, el_t, el_pos, el_neg;        // min-pressure decreasing crossing (first interval excluded)
    bool any_increasing, pos_parcel;
    // top of the profile where both curves are finite (PF:1143-1151)
    double top_p, top_a, top_b;
    bool any_avail, any_b;
    double min_p;

    XP_HD void init(double lcl_p_, double lcl_targ_, int pos_neg_) {
        lcl_p = lcl_p_; lcl_targ = lcl_targ_; pos_neg = pos_neg_;
        j = 0; xp_ = pp_ = ap_ = bp_ = qnan(); skip_first = false;
        s_pos = s_neg = 0.0;
        lcl_seen = false; lcl_pos = lcl_neg = 0.0;
        lfc_p = lfc_t = qnan(); lfc_pos = lfc_neg = 0.0;
        el_p = el_t = qnan(); el_pos = el_neg = 0.0;
        any_increasing = pos_parcel = false;
        top_p = top_a = top_b = qnan(); any_avail = any_b = false;
        min_p = qnan();
    }

    XP_HD void add_area(double area) {
        // NaN areas are skipped by the NaN-skipping sums (PF:206, 1365, 1382)
        if (area > 0.0) s_pos += area;
        else if (area < 0.0) s_neg += area;
    }

    // Feed the next profile level (pressure p, parcel curve a, environment curve b).
    XP_HD void emit(double p, double a, double b, bool is_lcl_level) { emit_x(p, xp_log(p), a, b, is_lcl_level); }
    // The same with x = ln p supplied by the caller (a reader that tabulates the logarithms of a shared pressure axis).
    XP_HD void emit_x(double p, double x, double a, double b, bool is_lcl_level) {
        // bookkeeping that is per level, not per interval
        if (!isnan(p) && !(p >= min_p)) min_p = p;                      // PF:1329 pressure.min()
        if (!isnan(b)) any_b = true;
        if (!isnan(a) && !isnan(b)) {                                   // PF:1143-1147
            if (!any_avail || p < top_p) { top_p = p; top_a = a; top_b = b; }
            any_avail = true;
        }
        if ((p < lcl_p) && (a > b)) pos_parcel = true;                  // PF:1166-1169
        if (j == 0) {
            skip_first = (a == b);                                      // PF:1117-1120
        } else {
            const double d0 = ap_ - bp_, d1 = a - b;
            const double s0 = sign_of(d0), s1 = sign_of(d1);
            const bool crossing = (s0 == s0) && (s1 == s1) && (s0 != s1);   // PF:1019-1022
            bool masked = false;
            if (crossing) {
                // find_intersections on (ln p, A, B): PF:1044-1053
                const double ix = (d1 * xp_ - d0 * x) / (d1 - d0);
                const double frac = (ix - xp_) / (x - xp_);
                const double iy = frac * (a - ap_) + ap_;
                const double px = xp_exp(ix);
                // the same crossing seen by trap_around_zeros on y = A - B vs 0: PF:1225-1237
                const double zy = frac * (d1 - d0) + d0;
                const double zx = xp_log(px);
                double area_lo = qnan(), area_hi = qnan();
                if (!isnan(zy)) {                                       // PF:1241-1244 masks
                    area_lo = (d0 / 2) * fabs(xp_ - zx);                // PF:1256-1261 (before zero)
                    area_hi = (d1 / 2) * fabs(x - zx);                  //              (after zero)
                    masked = !isnan(area_lo);                           // PF:1285-1287
                }
                add_area(area_lo);
                const bool in_above = (j >= 2);                         // PF:1108-1112
                const bool in_use = in_above || !skip_first;
                if (s1 > 0.0 && in_use && !isnan(px)) {                 // increasing (PF:1058)
                    any_increasing = true;                              // PF:1161
                    if ((px < lcl_p) && !(px <= lfc_p)) {               // PF:1127-1132 (max pressure)
                        lfc_p = px; lfc_t = iy; lfc_pos = s_pos; lfc_neg = s_neg;
                    }
                }
                if (s1 < 0.0 && in_above && !isnan(px)) {               // decreasing (PF:1060)
                    if (!(px >= el_p)) {                                // PF:1136 (min pressure)
                        el_p = px; el_t = iy; el_pos = s_pos; el_neg = s_neg;
                    }
                }
                add_area(area_hi);
            }
            if (!masked) {
                // plain trapezoid in ln p: PF:186-198 (|dx| * rolling mean)
                add_area(fabs(x - xp_) * ((d0 + d1) / 2));
            }
        }
        if (is_lcl_level && !lcl_seen) { lcl_seen = true; lcl_pos = s_pos; lcl_neg = s_neg; }
        xp_ = x; pp_ = p; ap_ = a; bp_ = b;
        ++j;
    }

    XP_HD void finish(ParcelResult &r, int post_zero) {
        // EL exists only if the parcel is not warmer than the environment at the top and the
        // EL is above the LCL (PF:1140-1155).
        const bool top_colder = top_a <= top_b;
        const bool el_exists = top_colder && (el_p < lcl_p);
        if (any_b && !any_avail) r.flags |= 1u;                         // PF:1149 assert
        double el_pp = el_exists ? el_p : qnan();
        double el_tt = el_exists ? el_t : qnan();
        const bool lfc_missing = !any_increasing;                       // PF:1161
        const bool replace = (pos_parcel && lfc_missing) ||             // PF:1170
                             (!lfc_missing && isnan(lfc_p) && (el_pp < lcl_p));   // PF:1174-1177
        double l_pos = lfc_pos, l_neg = lfc_neg;
        if (replace) { lfc_p = lcl_p; lfc_t = lcl_targ; l_pos = lcl_pos; l_neg = lcl_neg; }
        r.lfc_p = lfc_p; r.lfc_t = lfc_t; r.el_p = el_pp; r.el_t = el_tt;
        // cape_cin_base PF:1291-1392
        double cape = 0.0, cin = 0.0;
        if (!isnan(lfc_p)) {
            double e_pos = el_exists ? el_pos : s_pos;
            double e_neg = el_exists ? el_neg : s_neg;
            double el_eff = el_exists ? el_p : min_p;                   // PF:1329
            if (pos_neg) {
                cin = l_neg;
                cape = (el_eff > lfc_p) ? 0.0 : (e_pos - l_pos);
            } else {
                cin = l_pos + l_neg;
                cape = (el_eff > lfc_p) ? 0.0 : ((e_pos + e_neg) - (l_pos + l_neg));
            }
        }
        cape *= kRd; cin *= kRd;
        if (post_zero && !(cin <= 0.0)) cin = 0.0;                      // PF:1387-1388
        r.cape = cape; r.cin = cin;
    }
};

// linear_interp / log_interp (PF:1758-1828) between the bracketing levels `b` (coords >= at,
// "before") and `a` (coords <= at, "after"); cb/ca are their coordinates.
XP_HD double interp_bracket(double xb, double xa, double cb, double ca, double at) {
    double res = xb + (xa - xb) * ((at - cb) / (ca - cb));
    return (xb == xa) ? xb : res;                                       // PF:1806
}

// The environment level inserted at the LCL (PF:893-920).
XP_HD void env_at_lcl(bool have_before, double pb, double tb_, double tdb,
                                           bool have_after, double pa, double ta, double tda,
                                           double lcl_p, const Opts &o, double &t, double &td,
                                           double &tv) {
    if (!have_before || !have_after) { t = td = tv = qnan(); return; }
    double cb, ca, at;
    if (o.log_interp) { cb = xp_log(pb); ca = xp_log(pa); at = xp_log(lcl_p); }
    else { cb = pb; ca = pa; at = lcl_p; }
    t = interp_bracket(tb_, ta, cb, ca, at);
    td = interp_bracket(tdb, tda, cb, ca, at);
    tv = virtual_temperature(t, mixing_ratio_t_td(t, td, lcl_p, o.compat));   // PF:916-920
}

// One lifted column.  `Levels` provides n() and get(v, p, t, td) for the levels of the LIFTED
// column (for ML: mixed parcel + levels above the mixed layer; for MU: levels from the MU
// level up), already masked to NaN where the reference masks them.  `Prof` receives profile
// rows (may be a no-op).
template <class Levels, class Prof>
XP_HD void lift_parcel(const Levels &lv, double p0, double t0, double td0, const Tables &tb,
                            const Opts &o, ParcelResult &r, Prof &prof) {
    r.flags = 0;
    const int n = lv.n();
    // ---- LCL (PF:609-682) -----------------------------------------------------------------
    const bool valid = !(isnan(p0) || isnan(t0) || isnan(td0));
    double lcl_p = qnan(), lcl_t = qnan(), lcl_tv = qnan();
    if (valid) {
        lcl_solve(p0, t0, td0, lcl_p, lcl_t);
        lcl_tv = virtual_temperature(lcl_t, mixing_ratio_t_td(lcl_t, lcl_t, lcl_p, o.compat));
    }
    r.lcl_p = lcl_p; r.lcl_t = lcl_t; r.lcl_tv = lcl_tv;
    if (isnan(lcl_p)) {
        // insert_level with a NaN coordinate turns the whole profile into NaN (PF:965-985);
        // the NaN-skipping sums then give CAPE = CIN = 0.
        r.cape = r.cin = 0.0;
        r.lfc_p = r.lfc_t = r.el_p = r.el_t = qnan();
        ProfileRow row = {qnan(), qnan(), qnan(), qnan(), qnan(), qnan()};
        for (int v = 0; v <= n; ++v) prof.put(v, row);
        return;
    }
    // ---- parcel constants (PF:742-757) ------------------------------------------------------
    const double w_parcel = mixing_ratio_t_td(t0, td0, p0, o.compat);           // PF:748
    const int adiabat = adiabat_lookup(tb, lcl_p, lcl_t);                       // PF:554-557
    const float *curve = tb.curves + (size_t)(adiabat > 0 ? adiabat - 1 : 0) * kNP;

    Sweep sw;
    sw.init(lcl_p, o.vtc ? lcl_tv : lcl_t, o.pos_neg);
    bool inserted = false;
    bool have_prev = false;
    double pb = qnan(), tb_ = qnan(), tdb = qnan();
    int row_idx = 0;

    auto emit_lcl = [&](bool have_after, double pa, double ta, double tda) {
        // "before" = closest level with p >= lcl_p, "after" = closest with p <= lcl_p; when the
        // LCL coincides with a level both are that level (PF:1774-1775, 1798-1806).
        bool hb = have_prev, ha = have_after;
        double apa = pa, ata = ta, atda = tda;
        if (have_prev && pb == lcl_p) { ha = true; apa = pb; ata = tb_; atda = tdb; }
        double et, etd, etv;
        env_at_lcl(hb, pb, tb_, tdb, ha, apa, ata, atda, lcl_p, o, et, etd, etv);
        ProfileRow row = {lcl_p, lcl_t, lcl_tv, et, etv, etd};
        prof.put(row_idx++, row);
        sw.emit(lcl_p, o.vtc ? lcl_tv : lcl_t, o.vtc ? etv : et, true);
        inserted = true;
    };

    // Row evaluation (independent of the sweep state) is issued one level ahead of its use so that its
    // long dependent chains (exp / pow / table gathers) overlap the sweep of the previous row.
    struct RowEval { double p, x, t, td, env_tv, tp, tvp; };
    auto eval_row = [&](int v) {
        RowEval e;
        lv.get(v, e.p, e.t, e.td);
        e.x = lv.lnp(v, e.p);                                                   // ln p (PF:1019: crossings in ln p)
        // environment (PF:839-843)
        e.env_tv = virtual_temperature(e.t, mixing_ratio_t_td(e.t, e.td, e.p, o.compat));
        // parcel (PF:742-777)
        double wp;
        const double above = (adiabat > 0) ? adiabat_temperature(curve, e.p) : qnan();
        if (e.p >= lcl_p) e.tp = dry_lapse(e.p, t0, p0); else e.tp = above;     // PF:767
        if (e.p <= lcl_p) wp = sat_mixing_ratio(e.p, above); else wp = w_parcel;    // PF:760, 773
        e.tvp = virtual_temperature(e.tp, wp);
        return e;
    };
    RowEval cur = {qnan(), qnan(), qnan(), qnan(), qnan(), qnan(), qnan()};
    if (n > 0) cur = eval_row(0);
    for (int v = 0; v < n; ++v) {
        RowEval nxt = cur;
        if (v + 1 < n) nxt = eval_row(v + 1);
        const double p = cur.p, t = cur.t, td = cur.td;
        if (!inserted && !(p >= lcl_p)) emit_lcl(!isnan(p), p, t, td);          // PF:965-966
        // insert_level maps every variable at a NaN-pressure level to NaN (PF:963, 988)
        ProfileRow row = {p, cur.tp, cur.tvp, t, cur.env_tv, td};
        if (isnan(p)) row = {qnan(), qnan(), qnan(), qnan(), qnan(), qnan()};
        prof.put(row_idx++, row);
        sw.emit_x(row.p, cur.x, o.vtc ? row.tv : row.t, o.vtc ? row.env_tv : row.env_t, false);
        if (!isnan(p)) { have_prev = true; pb = p; tb_ = t; tdb = td; }
        cur = nxt;
    }
    if (!inserted) emit_lcl(false, qnan(), qnan(), qnan());
    sw.finish(r, o.post_zero);
}

}  // namespace xp
