// xp_parcels.cuh -- parcel selection (surface / mixed layer / most unstable) and the levels
// of the lifted column, per column.  Host-compilable for tests/hostsim (see xp_math.cuh).
#pragma once
#include "xp_column.cuh"

namespace xp {

// ln P(k): readers may tabulate it (`double LnP(int k) const`: xp_list.cu on a shared pressure axis); the others
// evaluate the logarithm.  Same function, same argument: the bits are the same either way.
template <class R, class = void>
struct ReaderLnP {
    static XP_HD double get(const R &, int, double p) { return xp_log(p); }
};
template <class R>
struct ReaderLnP<R, decltype((void)&R::LnP)> {
    static XP_HD double get(const R &r, int k, double) { return r.LnP(k); }
};

// Levels of the lifted column (see lift_parcel).  `Reader` gives P(k), Tk(k), Td(k), L.
template <class Reader>
struct LiftedLevels {
    Reader rd;
    int k0;             // first input level of the lifted column
    int pre;            // 1: a parcel level is prepended (mixed layer, PF:1641-1644)
    int mode;           // 0: keep all; 1: keep p <= thresh (PF:1551); 2: keep p < thresh (PF:1636)
    double thresh;
    double pre_p, pre_t, pre_td;
    XP_HD int n() const { return rd.L - k0 + pre; }
    XP_HD void get(int v, double &p, double &t, double &td) const {
        if (pre && v == 0) { p = pre_p; t = pre_t; td = pre_td; return; }
        const int k = k0 + v - pre;
        p = rd.P(k); t = rd.Tk(k); td = rd.Td(k);
        bool keep = (mode == 0) || (mode == 1 ? (p <= thresh) : (p < thresh));
        if (!keep) { p = t = td = qnan(); }
    }
    // ln of the pressure `p` that get(v, ...) returned
    XP_HD double lnp(int v, double p) const {
        if ((pre && v == 0) || isnan(p)) return xp_log(p);          // the prepended parcel level / a masked level (NaN)
        return ReaderLnP<Reader>::get(rd, k0 + v - pre, p);
    }
};

template <class Reader>
XP_HD double column_max_pressure(const Reader &rd) {
    double m = qnan();
    for (int k = 0; k < rd.L; ++k) {
        double p = rd.P(k);
        if (!isnan(p) && !(p <= m)) m = p;
    }
    return m;
}

// mixed_parcel (PF:229-289) = mixed_layer (PF:137-162) of theta and saturation mixing ratio of
// the dewpoint over get_layer(interpolate=True) (PF:63-100), then back to T and Td at the
// level-0 pressure.  Also returns the first level above the mixed layer (PF:1636).
template <class Reader>
XP_HD void mixed_parcel(const Reader &rd, double depth, double &mp_p, double &mp_t,
                        double &mp_td, int &k0, double &thresh) {
    const double bottom = column_max_pressure(rd);           // PF:80
    const double top = bottom - depth;                       // PF:84
    thresh = top;                                            // PF:1636 (same expression)
    k0 = rd.L;
    mp_p = rd.P(0);                                          // PF:250, 287
    double sum_th = 0.0, sum_w = 0.0;
    bool have_prev = false;
    double pp = qnan(), thp = qnan(), wp = qnan();
    for (int k = 0; k < rd.L; ++k) {
        const double p = rd.P(k);
        if (isnan(p)) break;
        const double th = potential_temperature(p, rd.Tk(k));      // PF:253
        const double w = sat_mixing_ratio(p, rd.Td(k));            // PF:258
        if (p >= top) {
            if (have_prev) {                                       // trapz(x='pressure') PF:186-198
                const double dx = fabs(p - pp);
                const double a_th = dx * ((thp + th) / 2), a_w = dx * ((wp + w) / 2);
                if (!isnan(a_th)) sum_th += a_th;
                if (!isnan(a_w)) sum_w += a_w;
            }
            have_prev = true; pp = p; thp = th; wp = w;
        } else {
            // first level above the layer: interpolate the layer top in ln p (PF:85-90)
            if (have_prev) {
                double tha = th, wa = w, pa = p;
                if (pp == top) { tha = thp; wa = wp; pa = pp; }
                const double cb = xp_log(pp), ca = xp_log(pa), at = xp_log(top);
                const double th_top = interp_bracket(thp, tha, cb, ca, at);
                const double w_top = interp_bracket(wp, wa, cb, ca, at);
                const double dx = fabs(top - pp);
                const double a_th = dx * ((thp + th_top) / 2), a_w = dx * ((wp + w_top) / 2);
                if (!isnan(a_th)) sum_th += a_th;
                if (!isnan(a_w)) sum_w += a_w;
            }
            k0 = k;
            break;
        }
    }
    const double pressure_depth = fabs(top - bottom);        // PF:158-159
    const double theta = (1. / pressure_depth) * sum_th;     // PF:161
    const double mr = (1. / pressure_depth) * sum_w;
    mp_t = theta * exner(mp_p);                              // PF:268-269
    mp_td = dewpoint_from_e(vapor_pressure(mp_p, mr));       // PF:275-282
    if (isnan(bottom)) { mp_t = mp_td = qnan(); }
}

// most_unstable_parcel (PF:102-135) with get_layer(interpolate=False) / bound_pressure (PF:208-227).
template <class Reader>
XP_HD void most_unstable_parcel(const Reader &rd, double depth, double &mu_p, double &mu_t,
                                double &mu_td, int &k0, uint32_t &flags) {
    const double bottom = column_max_pressure(rd);
    const double bound = bottom - depth;
    double best = qnan(), top = qnan();
    for (int k = 0; k < rd.L; ++k) {                          // bound_pressure PF:224-226
        const double p = rd.P(k);
        if (isnan(p)) continue;
        const double d = fabs(p - bound);
        if (isnan(d)) continue;
        if (!(d >= best)) { best = d; top = p; }
        else if (d == best && p > top) top = p;
    }
    double max_eq = qnan();
    mu_p = qnan(); k0 = rd.L;
    for (int k = 0; k < rd.L; ++k) {
        const double p = rd.P(k);
        if (!(p <= bottom) || !(p >= top)) continue;          // PF:97-98
        const double eq = theta_e(p, rd.Tk(k), rd.Td(k));     // PF:123
        if (isnan(eq)) continue;
        if (!(eq <= max_eq)) { max_eq = eq; mu_p = p; k0 = k; }        // PF:127
        else if (eq == max_eq && p > mu_p) { mu_p = p; k0 = k; }       // PF:128 (ties -> max pressure)
    }
    if (isnan(mu_p)) { mu_t = mu_td = qnan(); k0 = rd.L; return; }
    int count = 0;
    for (int k = 0; k < rd.L; ++k) {
        const double p = rd.P(k);
        if ((p <= bottom) && (p >= top) && p == mu_p) ++count;
    }
    if (count != 1) flags |= 2u;                              // PF:130-131
    mu_t = rd.Tk(k0);                                         // PF:133
    mu_td = rd.Td(k0);
}

// One column, one parcel kind (0 SB, 1 ML, 2 MU, 3 explicit): select the parcel, build the
// lifted column, lift.  `ex_*` is the explicit parcel (kind 3 only).  Writes profile rows
// 0..L through `prof` (NaN-padded above the lifted column).
template <class Reader, class Prof>
XP_HD void run_column(const Reader &rd, int kind, const Tables &tb, const Opts &o, double ex_p,
                      double ex_t, double ex_td, ParcelResult &r, double &p0, double &t0,
                      double &td0, int &shift, Prof &prof) {
    LiftedLevels<Reader> lv;
    lv.rd = rd; lv.k0 = 0; lv.pre = 0; lv.mode = 0; lv.thresh = qnan();
    lv.pre_p = lv.pre_t = lv.pre_td = qnan();
    uint32_t flags = 0;
    shift = 0;
    if (kind == 0) {                                        // surface-based, PF:1502-1504
        p0 = rd.P(0); t0 = rd.Tk(0); td0 = rd.Td(0);
    } else if (kind == 1) {                                 // mixed layer, PF:1604-1649
        int k0; double thresh;
        mixed_parcel(rd, o.ml_depth, p0, t0, td0, k0, thresh);
        lv.k0 = k0; lv.pre = 1; lv.mode = 2; lv.thresh = thresh;
        lv.pre_p = p0; lv.pre_t = t0; lv.pre_td = td0;
        shift = k0;
    } else if (kind == 2) {                                 // most unstable, PF:1517-1555
        int k0;
        most_unstable_parcel(rd, o.mu_depth, p0, t0, td0, k0, flags);
        lv.k0 = k0; lv.mode = 1; lv.thresh = p0;
        shift = k0;
    } else {                                                // explicit parcel, PF:1394
        p0 = ex_p; t0 = ex_t; td0 = ex_td;
    }
    lift_parcel(lv, p0, t0, td0, tb, o, r, prof);
    r.flags |= flags;
    const ProfileRow nanrow = {qnan(), qnan(), qnan(), qnan(), qnan(), qnan()};
    for (int v = lv.n() + 1; v <= rd.L; ++v) prof.put(v, nanrow);
}

}  // namespace xp
