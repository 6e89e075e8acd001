// xp_fast_pcol6.cuh -- float32 fast path for columns with PER-COLUMN pressure (model levels: BASELINE.json
// configs[1], [2]; the layout of the reference's own test_data.nc), sweep version 6, for the reference's
// default options and scalar outputs (profile rows and other option sets stay on xp_fast_pcol.cuh).
//
// Same decisions and hand-over rules as xp_fast_pcol.cuh; from xp_fast6.cuh it takes the branch-free step
// (P/N area sums, crossing fraction snapshotted and its position rebuilt after the sweep, three parcels as
// independent instruction streams), the staged parcel set-up and the cheaper float64 LCL polish.  Specific
// to per-column pressure:
//  * p/T/Td of the lowest levels are stashed per thread (shared memory in the kernel) by the pre-pass, whose
//    reads are announced as L2 prefetches -- the LCL search, the environment at the LCL and the first
//    iterations of the sweep read the stash instead of going back to global memory;
//  * the moist adiabat of a parcel is read from the float32 curve table (L2-resident) exactly as the
//    reference evaluates it: two neighbouring 0.5 hPa nodes of adiabat i, linear in p (PF:585-592), gathered
//    one iteration ahead;
//  * the mixed-layer top and the most-unstable level differ per column, so rows below a parcel's start are
//    neutralised (d := 0, x := start row) instead of splitting the loop.
#pragma once
#include "xp_fast6.cuh"
#include "xp_fast_pcol.cuh"

namespace xp {
namespace fast {

struct NoStash3 {
    XP_HD int capacity() const { return 0; }
    XP_HD void put(int, float, float, float) const {}
    XP_HD void get(int, float &p, float &t, float &td) const { p = t = td = 0.0f; }
};

struct Setup6P {
    double lp, lt, p0, t0, td0;
    int adiabat;
    float pka, pkb;             // pressure of the levels after / before the LCL
    float ta, tda, tb_, tdb;
    bool before_is_start;
};

// `lev(k, p, t, td)` reads level k (stash or global memory).
template <class Lev>
XP_HD void setup6p_b(const Lev &lev, int L, const Tables &tb, int knext, PColParcel &pc, Setup6P &u) {
    float edge;
    u.adiabat = adiabat_cell(tb, u.lp, u.lt, edge);                   // PF:554-557 (2-byte gather)
    pc.curve = tb.curves + (size_t)(u.adiabat > 0 ? u.adiabat - 1 : 0) * kNP;
    pc.kfirst = knext;
    // LCL position among the levels of the lifted column (insert_level PF:965-966): float32 pressures against
    // the float32-rounded LCL pressure decide every case except equality (-> exact path)
    const float lpf = (float)u.lp;
    int ka = knext;
    float pka = 0.0f, pkb = (float)u.p0, ta = 0.0f, tda = 0.0f, tb_ = 0.0f, tdb = 0.0f;
    bool tie = (pkb == lpf);
    while (ka < L) {
        float t, td;
        lev(ka, pka, t, td);
        tb_ = ta; tdb = tda; ta = t; tda = td;
        if (!(pka >= lpf)) break;
        tie = tie || (pka == lpf);
        pkb = pka; ++ka;
    }
    u.before_is_start = (ka == knext);
    if (ka >= L || tie) { pc.bad = true; ka = L; }                   // LCL above the top / (to float32) on a level
    pc.ka = ka;
    u.pka = pka; u.pkb = pkb; u.ta = ta; u.tda = tda; u.tb_ = tb_; u.tdb = tdb;
}

XP_HD void setup6p_c(const Opts &o, float x_start, int L, PColParcel &pc, const Setup6P &u) {
    if (u.adiabat <= 0) pc.bad = true;
    pc.m = 0; pc.f = 0.0f; pc.f0 = pc.f1 = 0.0f;
    const float p0f = (float)u.p0, t0f = (float)u.t0, td0f = (float)u.td0;
    const float lpf = (float)u.lp, ltf = (float)u.lt;
    pc.lcl_p = lpf; pc.lcl_t = ltf;
    const float es_l = f_es(ltf);
    pc.lcl_tv = f_tv(ltf, f_mixing_ratio(es_l, es_l, lpf, 141));                   // PF:653-657
    const float w_parcel = f_mixing_ratio(f_es(t0f), f_es(td0f), p0f, 141);        // PF:748
    pc.c_dryv = t0f * f_ex2(-(float)kKappa * f_lg2(p0f)) * f_fma(0.608f, w_parcel, 1.0f);   // PF:291-316, 767-775
    sweep_init6(pc, x_start);
    pc.x_lcl = x_start; pc.a_lcl = pc.b_lcl = 0.0f;
    if (pc.ka >= L) return;                                                        // bound for the exact path
    float tb_ = u.tb_, tdb = u.tdb;
    if (u.before_is_start) { tb_ = t0f; tdb = td0f; }
    const float x_l = kLn2 * f_lg2(lpf);
    float xb, xa, at;
    if (o.log_interp) { xb = kLn2 * f_lg2(u.pkb); xa = kLn2 * f_lg2(u.pka); at = x_l; }
    else { xb = u.pkb; xa = u.pka; at = lpf; }
    const float g = (at - xb) * f_rcp(xa - xb);
    const float te = f_fma(u.ta - tb_, g, tb_), tde = f_fma(u.tda - tdb, g, tdb);   // PF:1802
    const float etv = f_tv(te, f_mixing_ratio(f_es(te), f_es(tde), lpf, 141));      // PF:916-920
    pc.x_lcl = x_l;
    pc.a_lcl = pc.lcl_tv;
    pc.b_lcl = pc.lcl_tv - etv;                                                    // the LCL row as a difference
    if (!(te == te) || !(tde == tde)) pc.bad = true;
}

// Virtual temperature of the parcel on its moist adiabat at pressure p (PF:585-592, 760, 775).
XP_HD float moist_tv_at(const PColParcel &c, float p) {
    const float s = (p - 2.5f) * 2.0f;
    const int j = min(max((int)s, 0), kNP - 2);
    const float tm = adiabat_temperature_f32(c.curve, j, s - (float)j);
    const float es = f_es(tm);
    return f_tv(tm, kEpsF * es * f_rcp(p - es));
}

// lfc_el PF:1140-1185 + cape_cin_base PF:1329-1388 on the v6 state; `pres(k)` gives the pressure of level k.
template <class Pres>
XP_HD void sweep_finish6p(const PColParcel &s, const Pres &pres, const Opts &o, FResult &r) {
    const bool top_colder = s.dprev <= 0.0f;                            // PF:1151
    const bool el_exists = top_colder && s.el_it > s.ka;                // PF:1152-1153
    const bool lfc_missing = s.n_inc == 0;                              // PF:1161
    const bool lfc_found = s.lfc_it != 0;
    const bool pos_parcel = s.max_d_above > 0.0f;
    const bool replace = (pos_parcel && lfc_missing) || (!lfc_missing && !lfc_found && el_exists);
    const bool have_lfc = lfc_found || replace;
    auto crossing = [&](int itc, float fr, float &px, float &y) {
        const int kc = itc - 1;                                          // level of the upper row
        const float p1 = pres(kc);
        const float x1 = kLn2 * f_lg2(p1), a1 = moist_tv_at(s, p1);
        float x0 = s.x_lcl, a0 = s.a_lcl;                                // lower row: the LCL row ...
        if (itc != s.ka + 1) { const float p0 = pres(kc - 1); x0 = kLn2 * f_lg2(p0); a0 = moist_tv_at(s, p0); }
        px = f_ex2(f_fma(-fr, x0 - x1, x0) * kLog2e);
        y = f_fma(fr, a1 - a0, a0);
    };
    r.lfc_p = r.lfc_t = r.el_p = r.el_t = f_qnan();
    if (lfc_found) crossing(s.lfc_it, s.lfc_x, r.lfc_p, r.lfc_t);
    if (replace) { r.lfc_p = s.lcl_p; r.lfc_t = s.lcl_tv; }
    if (el_exists) crossing(s.el_it, s.el_x, r.el_p, r.el_t);
    float cape = 0.0f, cin = 0.0f;
    if (have_lfc) {
        const float l_P = replace ? s.lcl_pos : s.lfc_pos;
        const float l_N = replace ? s.lcl_tot : s.lfc_tot;
        const float e_P = el_exists ? s.el_pos : s.pos;
        const bool el_below_lfc = el_exists && lfc_found && !replace && s.el_it < s.lfc_it;   // PF:1352-1353
        cin = l_N;
        cape = el_below_lfc ? 0.0f : (e_P - l_P);
    }
    cape *= (float)kRd; cin *= (float)kRd;
    if (o.post_zero && !(cin <= 0.0f)) cin = 0.0f;
    r.cape = cape; r.cin = cin;
    r.lcl_p = s.lcl_p; r.lcl_t = s.lcl_t; r.lcl_tv = s.lcl_tv;
}

// One parcel, one iteration: gathers for the next iteration, moist adiabat of level it-1, the step.
template <int GUARD>
XP_HD void parcel_iteration_pcol6(PColParcel &c, int it, int j_cur, float w_prv, float pk_cur, float p_prv,
                                  float x_cur, float x_prv, float b_cur, float b_prv) {
    const float f0 = c.f0, f1 = c.f1;
    c.f0 = XP_LDG(c.curve + j_cur); c.f1 = XP_LDG(c.curve + j_cur + 1);         // for the next iteration
    const float tm = f_fma(f1 - f0, w_prv, f0);                                // np.interp, PF:585-592
    const float es = f_es(tm);
    const float a_m = f_tv(tm, kEpsF * es * f_rcp(p_prv - es));                 // PF:760, 775
    step6_core<GUARD>(c, it, a_m - b_prv, f_fma(c.c_dryv, pk_cur, -b_cur), x_cur, x_prv);
}

// The suite for one column with its own pressure profile, default options, scalar outputs.
// Rd: ldP/ldT/ldTd(off), prefetch(off_p, off), ls(), pls(), off0() (32-bit element offsets, see Sweep6).
// Stash: capacity(), put(k, p, t, td), get(k, p, t, td).  Returns the redo mask (see suite_column).
template <unsigned KINDS, class Rd, class Stash>
XP_HD unsigned suite_column_pcol6(const Rd &rd, int L, const Tables &tb, const Opts &o, Stash &stash, FResult res[3]) {
    unsigned redo = 0;
    float nanacc = 0.0f;
    bool bad_axis = false;                 // pressure not finite / not strictly decreasing / outside the table
    const uint32_t ls = rd.ls(), pls = rd.pls();
    const int cap = min(stash.capacity(), L);
    // ask for the lowest levels now (the pre-pass rarely needs more than the stash holds)
    {
        const int n_pf = min(L, max(cap, 8));
        for (int k = 1; k < n_pf; ++k) rd.prefetch(rd.off0() + (uint32_t)k * pls, rd.off0() + (uint32_t)k * ls);
    }
    const float p_sfc = rd.ldP(rd.off0()), t_sfc = rd.ldT(rd.off0()), td_sfc = rd.ldTd(rd.off0());
    const double bottom = (double)p_sfc;                                         // PF:80 (pressure decreases upward)
    if (!(p_sfc <= 1100.0f)) bad_axis = true;
    // ---- pre-pass: mixed-layer means (float64) and most-unstable argmax over the lowest levels ------------
    const double top_ml = bottom - o.ml_depth;                                   // PF:84, PF:1636
    const double bound_mu = bottom - o.mu_depth;                                 // PF:92
    double sum_th = 0.0, sum_w = 0.0, pp = bottom, thp = 0.0, wp = 0.0;
    int K_ml = L;
    bool ml_done = !(KINDS & 2u), mu_done = !(KINDS & 4u);
    float best = -1e30f, second = -1e30f, mu_t = 0.0f, mu_td = 0.0f, mu_p = p_sfc;
    int k_mu = 0, n_pre = 0;
    {
        float p_nx = p_sfc, t_nx = t_sfc, td_nx = td_sfc;
        uint32_t offp = rd.off0(), off = rd.off0();
#pragma unroll 1
        for (int k = 0; k < L && (k < cap || !(ml_done && mu_done)); ++k) {
            const float pf = p_nx, t = t_nx, td = td_nx;
            offp += pls; off += ls;
            if (k + 1 < L) { p_nx = rd.ldP(offp); t_nx = rd.ldT(off); td_nx = rd.ldTd(off); }
            if (k < cap) stash.put(k, pf, t, td);
            n_pre = k + 1;
            if (ml_done && mu_done) continue;                                    // only filling the stash
            nanacc = f_fma(pf, 0.0f, f_fma(t, 0.0f, f_fma(td, 0.0f, nanacc)));
            const double p = (double)pf;
            if (k > 0 && !(p < pp)) bad_axis = true;
            if ((KINDS & 2u) && !ml_done) {
                // mixed_parcel PF:229-289 on get_layer(interpolate=True) PF:63-100, trapz in p PF:186-198
                const double tdd = (double)td;
                const double e = kSat0 * exp(17.67 * (tdd - 273.15) * rcp64(tdd - 29.65));
                const double th = (double)t * exp64_fast(-kKappa * log64_fast(p * 1e-3));   // PF:253
                const double w = kEps * e * rcp64(p - e);                            // PF:258
                if (p >= top_ml) {
                    if (k > 0) {
                        const double dx = fabs(p - pp);
                        sum_th += dx * ((thp + th) / 2); sum_w += dx * ((wp + w) / 2);
                    }
                    thp = th; wp = w;
                } else {
                    if (k > 0 && pp != top_ml) {                                     // layer top in ln p (PF:85-90)
                        const double cb = log64_fast(pp), ca = log64_fast(p), at = log64_fast(top_ml);
                        const double g = (at - cb) * rcp64(ca - cb);
                        const double dx = fabs(top_ml - pp);
                        sum_th += dx * ((thp + (thp + (th - thp) * g)) / 2);
                        sum_w += dx * ((wp + (wp + (w - wp) * g)) / 2);
                    }
                    K_ml = k; ml_done = true;
                }
            }
            if ((KINDS & 4u) && !mu_done) {
                // layer of most_unstable_parcel: levels down to the one closest to bottom - depth (PF:208-227)
                bool in_layer = true;
                if (p < bound_mu) {
                    in_layer = (k > 0) && ((bound_mu - p) < (pp - bound_mu));
                    mu_done = true;
                }
                if (in_layer) {
                    const float e = f_es(td);
                    const float ipe = f_rcp(pf - e);
                    const float r = kEpsF * e * ipe;
                    const float l2t = f_lg2(t), l2td = f_lg2(td);
                    const float t_l = 56.0f + f_rcp(f_rcp(td - 56.0f) + (l2t - l2td) * (kLn2 / 800.0f));
                    const float it_l = f_rcp(t_l);
                    float v = l2t * kLn2;
                    v = f_fma((float)kKappa * kLn2, f_lg2(1000.0f * ipe), v);
                    v = f_fma(0.28f * r * kLn2, l2t - f_lg2(t_l), v);
                    v = f_fma(r * f_fma(0.448f, r, 1.0f), f_fma(3036.0f, it_l, -1.78f), v);
                    nanacc = f_fma(v, 0.0f, nanacc);
                    if (v > best) { second = best; best = v; k_mu = k; mu_t = t; mu_td = td; mu_p = pf; }
                    else if (v > second) second = v;
                }
            }
            pp = p;
        }
    }
    const int n_stash = min(cap, n_pre);
    auto lev = [&](int k, float &p, float &t, float &td) {
        if (k < n_stash) stash.get(k, p, t, td);
        else { p = rd.ldP(rd.off0() + (uint32_t)k * pls); const uint32_t o_ = rd.off0() + (uint32_t)k * ls; t = rd.ldT(o_); td = rd.ldTd(o_); }
    };
    auto pres = [&](int k) { float p, t, td; lev(k, p, t, td); return p; };
    // start the global pipeline of the sweep at the first level the stash does not hold
    const int k_g = max(n_stash, 1);
    uint32_t g_offp = rd.off0() + (uint32_t)k_g * pls, g_off = rd.off0() + (uint32_t)k_g * ls;
    float p_n1 = 0.0f, t_n1 = 0.0f, td_n1 = 0.0f;
    if (k_g < L) { p_n1 = rd.ldP(g_offp); t_n1 = rd.ldT(g_off); td_n1 = rd.ldTd(g_off); }
#pragma unroll
    for (int j = 1; j <= kL2Ahead; ++j)
        if (k_g + j < L) rd.prefetch(g_offp + (uint32_t)j * pls, g_off + (uint32_t)j * ls);
    g_offp += pls; g_off += ls;
    int k_pf = k_g + 1;
    // ---- parcels: staged (LCL solves, gathers, consumers) -----------------------------------------------------
    PColParcel sb, ml, mu;
    Setup6P u_sb, u_ml, u_mu;
    const float x_sfc = kLn2 * f_lg2(p_sfc);
    nanacc = f_fma(p_sfc, 0.0f, f_fma(t_sfc, 0.0f, f_fma(td_sfc, 0.0f, nanacc)));
    double mp_t = 0.0, mp_td = 0.0;
    auto stage_a = [&](double p0, double t0, double td0, PColParcel &pc, Setup6P &u) {
        pc.bad = false;
        if (!(t0 - td0 >= kSaturationMargin) || !(p0 > 0.0)) { pc.bad = true; t0 = 280.0; td0 = 270.0; p0 = 1000.0; }
        u.p0 = p0; u.t0 = t0; u.td0 = td0;
        lcl_fast6(p0, t0, td0, u.lp, u.lt);
    };
    if (KINDS & 1u) stage_a(bottom, (double)t_sfc, (double)td_sfc, sb, u_sb);
    if (KINDS & 2u) {
        const double depth = fabs(top_ml - bottom);                              // PF:158-159
        mixed_parcel_t_td(bottom, (1. / depth) * sum_th, (1. / depth) * sum_w, mp_t, mp_td);   // PF:161, 268-282
        if (!ml_done || K_ml < 1) { redo |= 2u; K_ml = max(K_ml, 1); }           // no level above / NaN layer: exact path
        stage_a(bottom, mp_t, mp_td, ml, u_ml);
    }
    if (KINDS & 4u) {
        if (!(best - second >= kThetaEMargin)) redo |= 4u;
        stage_a((double)mu_p, (double)mu_t, (double)mu_td, mu, u_mu);
    }
    if (KINDS & 1u) setup6p_b(lev, L, tb, 1, sb, u_sb);
    if (KINDS & 2u) setup6p_b(lev, L, tb, K_ml, ml, u_ml);
    if (KINDS & 4u) setup6p_b(lev, L, tb, k_mu + 1, mu, u_mu);
    if (KINDS & 1u) {
        setup6p_c(o, x_sfc, L, sb, u_sb);
        res[0].par_p = p_sfc; res[0].par_t = t_sfc; res[0].par_td = td_sfc; res[0].shift = 0;
    }
    if (KINDS & 2u) {
        setup6p_c(o, x_sfc, L, ml, u_ml);
        res[1].par_p = p_sfc; res[1].par_t = (float)mp_t; res[1].par_td = (float)mp_td; res[1].shift = K_ml;
    }
    if (KINDS & 4u) {
        setup6p_c(o, kLn2 * f_lg2(mu_p), L, mu, u_mu);
        res[2].par_p = mu_p; res[2].par_t = mu_t; res[2].par_td = mu_td; res[2].shift = k_mu;
    }
    // ---- the sweep ----------------------------------------------------------------------------------------
    float b_prv = 0.0f, x_prv = x_sfc, p_prv = p_sfc, w_prv;
    {   // node/weight of the surface pressure and the first gathers
        const float s0 = (p_sfc - 2.5f) * 2.0f;
        const int j0 = min(max((int)s0, 0), kNP - 2);
        w_prv = s0 - (float)j0;
        if (KINDS & 1u) { sb.f0 = XP_LDG(sb.curve + j0); sb.f1 = XP_LDG(sb.curve + j0 + 1); }
        if (KINDS & 2u) { ml.f0 = XP_LDG(ml.curve + j0); ml.f1 = XP_LDG(ml.curve + j0 + 1); }
        if (KINDS & 4u) { mu.f0 = XP_LDG(mu.curve + j0); mu.f1 = XP_LDG(mu.curve + j0 + 1); }
    }
#pragma unroll 1
    for (int it = 1; it < L; ++it) {
        float p_cur, t, td;
        if (it < n_stash) {
            stash.get(it, p_cur, t, td);
        } else {
            p_cur = p_n1; t = t_n1; td = td_n1;
            if (k_pf < L) { p_n1 = rd.ldP(g_offp); t_n1 = rd.ldT(g_off); td_n1 = rd.ldTd(g_off); }
            if (k_pf + kL2Ahead < L) rd.prefetch(g_offp + kL2Ahead * pls, g_off + kL2Ahead * ls);
            g_offp += pls; g_off += ls; ++k_pf;
        }
        if (!(p_cur < p_prv) || !(p_cur >= 2.5f)) bad_axis = true;
        // table node and weight of this level's pressure (used by the lagging rows of the next iteration)
        const float s_cur = (p_cur - 2.5f) * 2.0f;
        const int j_cur = min(max((int)s_cur, 0), kNP - 2);
        const float w_cur = s_cur - (float)j_cur;
        const float l2p = f_lg2(p_cur);
        const float x_cur = kLn2 * l2p, pk_cur = f_ex2((float)kKappa * l2p);
        const float b_cur = f_tv(t, f_mixing_ratio(f_es(t), f_es(td), p_cur, 141));   // PF:839-843
        if (KINDS & 1u) parcel_iteration_pcol6<0>(sb, it, j_cur, w_prv, pk_cur, p_prv, x_cur, x_prv, b_cur, b_prv);
        if (KINDS & 2u) parcel_iteration_pcol6<2>(ml, it, j_cur, w_prv, pk_cur, p_prv, x_cur, x_prv, b_cur, b_prv);
        if (KINDS & 4u) parcel_iteration_pcol6<2>(mu, it, j_cur, w_prv, pk_cur, p_prv, x_cur, x_prv, b_cur, b_prv);
        b_prv = b_cur; x_prv = x_cur; p_prv = p_cur; w_prv = w_cur;
    }
    // last iteration: no level L; every parcel that is not bound for the exact path is above its LCL
    if (KINDS & 1u) parcel_iteration_pcol6<0>(sb, L, 0, w_prv, 0.0f, p_prv, x_prv, x_prv, 1e30f, b_prv);
    if (KINDS & 2u) parcel_iteration_pcol6<0>(ml, L, 0, w_prv, 0.0f, p_prv, x_prv, x_prv, 1e30f, b_prv);
    if (KINDS & 4u) parcel_iteration_pcol6<0>(mu, L, 0, w_prv, 0.0f, p_prv, x_prv, x_prv, 1e30f, b_prv);
    // ---- results ----------------------------------------------------------------------------------------------
    bool nan_seen = !(nanacc == 0.0f) || bad_axis;
    if (KINDS & 1u) nan_seen = nan_seen || !(sb.pos - sb.tot < 3e38f);
    if (KINDS & 2u) nan_seen = nan_seen || !(ml.pos - ml.tot < 3e38f);
    if (KINDS & 4u) nan_seen = nan_seen || !(mu.pos - mu.tot < 3e38f);
    auto wrap = [&](const PColParcel &c, FResult &r, unsigned bit) {
        sweep_finish6p(c, pres, o, r);
        const bool unc = !(c.min_abs_d >= kDecisionEps) || !(c.min_slope >= 0.0f);
        if (c.bad || unc || nan_seen) redo |= bit;
    };
    if (KINDS & 1u) wrap(sb, res[0], 1u);
    if (KINDS & 2u) wrap(ml, res[1], 2u);
    if (KINDS & 4u) wrap(mu, res[2], 4u);
    if ((KINDS & 5u) == 5u && (redo & 4u) && k_mu == 0 && !nan_seen && (best - second >= kThetaEMargin))
        redo = (redo & ~4u) | 1u | kRedoMuIsSb;
    return redo;
}

}  // namespace fast
}  // namespace xp
