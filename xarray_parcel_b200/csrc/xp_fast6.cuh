// xp_fast6.cuh -- the float32 fast path of the suite on a shared pressure axis, sweep version 6,
// for the reference's DEFAULT options (virtual temperature correction PF:1394-1475, MetPy 1.4.1
// formulas, pos_cape_neg_cin PF:1329-1388).  Same decisions and hand-over rules as xp_fast.cuh; what
// changes is the instruction count of the per-(level, parcel) step:
//   * the shared-memory cubics give the parcel's VIRTUAL temperature above the LCL directly
//     (saturation mixing ratio of the adiabat temperature folded into the table, PF:760/775), so a
//     row above the LCL costs one LDS.128 + 3 FFMA instead of cubic + Bolton es + mixing ratio;
//   * areas: with S = h (d0 + d1) the whole trapezoid and, at a sign change, a_lo = h d0^2/(d0 - d1),
//     a_hi = S - a_lo, the positive / negative parts of ANY interval are max(a_lo', a_hi') and
//     min(a_lo', a_hi') with a_lo' = cross ? a_lo : S  (PF:1200-1289 + PF:164-206 in one form);
//   * running sums P (positive parts) and N (negative parts) replace pos/tot: CAPE = P(EL) - P(LFC),
//     CIN = N(LFC)  (PF:1329-1388 with pos_cape_neg_cin);
//   * the crossing position is NOT evaluated in the loop: only the fraction and the iteration are
//     snapshotted, and LFC/EL pressure and temperature are rebuilt after the sweep from the two rows;
//   * no branches in the step: the three parcels of a column are independent instruction streams
//     that the scheduler interleaves; rows below a parcel's start are neutralised by d := 0, and the
//     loop is split where the set of live parcels changes (mixed layer / most-unstable search top)
//     so that guards are only paid where they can matter.
#pragma once
#include "xp_fast.cuh"

namespace xp {
namespace fast {

// coef[k][m] of the VIRTUAL temperature of the saturated parcel on adiabat (m-1..m+2)*64 at level k
XP_HD Coef compute_coef_tv(const Prep &pr, const float *curves, int k, int m) {
    double y[4];
    for (int j = 0; j < 4; ++j) {
        int a = (m - 1 + j) * kNodeStride;
        a = min(max(a, 0), kNAdiabats - 1);
        const double t = adiabat_temperature(curves + (size_t)a * kNP, pr.p64[k]);
        y[j] = virtual_temperature(t, sat_mixing_ratio(pr.p64[k], t));          // PF:760, 775
    }
    Coef c;
    c.c0 = (float)y[1];
    c.c1 = (float)(-y[0] / 3 - y[1] / 2 + y[2] - y[3] / 6);
    c.c2 = (float)(y[0] / 2 - y[1] + y[2] / 2);
    c.c3 = (float)(-y[0] / 6 + y[1] / 2 - y[2] / 2 + y[3] / 6);
    return c;
}

XP_HD float cubic_at(const Coef &cc, float f) {
    return f_fma(f_fma(f_fma(cc.c3, f, cc.c2), f, cc.c1), f, cc.c0);
}

// Parcel state of the v6 sweep.  FParcel's fields are reused with these meanings:
//   pos = P, tot = N, lcl_pos/lcl_tot = P/N after the LCL row, lfc_pos = P below the LFC crossing,
//   lfc_tot = N including the triangle below the LFC, el_pos = P including the triangle below the EL,
//   lfc_x / el_x = crossing FRACTION of the LFC / EL interval, b_lcl = (parcel - environment) at the LCL row.
XP_HD void sweep_init6(FParcel &c, float x0) {
    c.xprev = x0; c.dprev = 0.0f;
    c.pos = c.tot = c.lcl_pos = c.lcl_tot = 0.0f;
    c.lfc_pos = c.lfc_tot = c.lfc_x = 0.0f;
    c.el_pos = c.el_x = 0.0f;
    c.lfc_it = c.el_it = 0; c.n_inc = 0;
    c.min_abs_d = 1e30f; c.min_slope = 1e30f; c.max_d_above = -1e30f;
}

// One row of one parcel at iteration `it` (row schedule: see FParcel in xp_fast.cuh).
//   cc      cubic of level it-1 for this parcel's adiabat interval
//   *_cur   level it, *_prv level it-1:  pk = p^kappa, x = ln p, b = environment virtual temperature
// GUARD: the parcel may start above this row (most-unstable parcel): rows it < kfirst are neutral.
template <bool GUARD>
XP_HD void step6(FParcel &c, int it, const Coef &cc, float pk_cur, float x_cur, float x_prv, float b_cur,
                 float b_prv) {
    const bool above = it > c.ka, is_lcl = it == c.ka;
    const float d_m = cubic_at(cc, c.f) - b_prv;                  // moist adiabat (PF:585-592), level it-1
    const float d_d = f_fma(c.c_dryv, pk_cur, -b_cur);            // dry adiabat (PF:742), level it
    float d = above ? d_m : d_d;
    d = is_lcl ? c.b_lcl : d;
    float x = above ? x_prv : x_cur;
    x = is_lcl ? c.x_lcl : x;
    bool active = true;
    if (GUARD) { active = it >= c.kfirst; d = active ? d : 0.0f; }
    const float dx = c.xprev - x;
    const float h = 0.5f * dx;
    const float den = c.dprev - d;
    const bool cross = c.dprev * d < 0.0f;                       // PF:1026-1031
    const float S = (c.dprev + d) * h;
    const float fr = c.dprev * f_rcp(den);                       // zero at xprev - fr dx
    const float alo_c = (c.dprev * h) * fr;
    const float alo = cross ? alo_c : S;
    const float ahi = S - alo;
    const float pinc = fmaxf(alo, ahi), ninc = fminf(alo, ahi);
    const bool inc = cross && d > 0.0f;                          // PF:1058
    const bool dec = cross && !(d > 0.0f);                       // PF:1060
    // LFC: max-pressure increasing crossing above the LCL (PF:1127-1132) = the first one met
    const bool take = inc && above && c.lfc_it == 0;
    c.lfc_pos = take ? c.pos : c.lfc_pos;
    c.pos += pinc; c.tot += ninc;
    c.lfc_tot = take ? c.tot : c.lfc_tot;
    c.lfc_x = take ? fr : c.lfc_x;
    c.lfc_it = take ? it : c.lfc_it;
    // EL: min-pressure decreasing crossing (PF:1136) = the last one met
    c.el_it = dec ? it : c.el_it;
    c.el_x = dec ? fr : c.el_x;
    c.el_pos = dec ? c.pos : c.el_pos;
    c.lcl_pos = is_lcl ? c.pos : c.lcl_pos; c.lcl_tot = is_lcl ? c.tot : c.lcl_tot;
    c.n_inc = inc ? 1 : c.n_inc;                                 // PF:1161 only asks "any"
    c.max_d_above = fmaxf(c.max_d_above, above ? d : -1e30f);    // PF:1166-1169
    c.min_abs_d = fminf(c.min_abs_d, (!GUARD || active) ? fabsf(d) : 1e30f);
    c.min_slope = fminf(c.min_slope, cross ? f_fma(-kCrossSlope, dx, fabsf(den)) : 1e30f);
    c.xprev = x; c.dprev = d;
}

// ln p and parcel virtual temperature of a crossing found at iteration `itc` (> ka) with fraction fr.
template <class Cf>
XP_HD void crossing6(const FParcel &c, const Cf &cf, const Prep &pr, int itc, float fr, float &x, float &y) {
    const int kc = itc - 1;                                      // level of the upper row
    const float x1 = pr.lnp[kc];
    const float a1 = cubic_at(cf.row(kc).at(c.m), c.f);
    float x0 = c.x_lcl, a0 = c.a_lcl;                            // lower row: the LCL row ...
    if (itc != c.ka + 1) { x0 = pr.lnp[kc - 1]; a0 = cubic_at(cf.row(kc - 1).at(c.m), c.f); }   // ... or level kc-1
    x = f_fma(-fr, x0 - x1, x0);
    y = f_fma(fr, a1 - a0, a0);
}

// lfc_el PF:1140-1185 + cape_cin_base PF:1329-1388 on the v6 state.
template <class Cf>
XP_HD void sweep_finish6(const FParcel &s, const Cf &cf, const Prep &pr, const Opts &o, FResult &r) {
    const bool top_colder = s.dprev <= 0.0f;                            // PF:1151
    const bool el_exists = top_colder && s.el_it > s.ka;                // PF:1152-1153
    const bool lfc_missing = s.n_inc == 0;                              // PF:1161
    const bool lfc_found = s.lfc_it != 0;
    const bool pos_parcel = s.max_d_above > 0.0f;
    const bool replace = (pos_parcel && lfc_missing) || (!lfc_missing && !lfc_found && el_exists);
    const bool have_lfc = lfc_found || replace;
    r.lfc_p = r.lfc_t = r.el_p = r.el_t = f_qnan();
    if (lfc_found) {
        float x, y;
        crossing6(s, cf, pr, s.lfc_it, s.lfc_x, x, y);
        r.lfc_p = f_ex2(x * kLog2e); r.lfc_t = y;
    }
    if (replace) { r.lfc_p = s.lcl_p; r.lfc_t = s.lcl_tv; }
    if (el_exists) {
        float x, y;
        crossing6(s, cf, pr, s.el_it, s.el_x, x, y);
        r.el_p = f_ex2(x * kLog2e); r.el_t = y;
    }
    float cape = 0.0f, cin = 0.0f;
    if (have_lfc) {
        const float l_P = replace ? s.lcl_pos : s.lfc_pos;
        const float l_N = replace ? s.lcl_tot : s.lfc_tot;
        const float e_P = el_exists ? s.el_pos : s.pos;
        // EL below the LFC (PF:1352-1353 leaves no level between them)
        const bool el_below_lfc = el_exists && lfc_found && !replace && s.el_it < s.lfc_it;
        cin = l_N;
        cape = el_below_lfc ? 0.0f : (e_P - l_P);
    }
    cape *= (float)kRd; cin *= (float)kRd;
    if (o.post_zero && !(cin <= 0.0f)) cin = 0.0f;
    r.cape = cape; r.cin = cin;
    r.lcl_p = s.lcl_p; r.lcl_t = s.lcl_t; r.lcl_tv = s.lcl_tv;
}

// Shared per-level state of the sweep.
template <class Rd>
struct Sweep6 {
    const float *lp_x, *lp_k, *lp_p;      // axis constants of level `it`
    const float *tp, *tdp;                // T/Td of level it + 1 (prefetch)
    int64_t ls;
    float t_nxt, td_nxt;
    float b_prv, x_prv;
};

// Iterations [it0, it1) of the sweep for the parcels in KACT (subset of KINDS); it1 <= nt - 1 + 1.
template <unsigned KACT, bool GUARD_MU, class Rd, class CoefRow>
XP_HD void sweep_segment6(Sweep6<Rd> &s, CoefRow &crow, int it0, int it1, int nt, FParcel &sb, FParcel &ml, FParcel &mu) {
    for (int it = it0; it < it1; ++it) {
        const float t = s.t_nxt, td = s.td_nxt;
        s.tp += s.ls; s.tdp += s.ls;
        if (it + 1 < nt) { s.t_nxt = Rd::ld(s.tp); s.td_nxt = Rd::ld(s.tdp); }      // prefetch the next level
        const float p_cur = *s.lp_p++, x_cur = *s.lp_x++, pk_cur = *s.lp_k++;
        const float b_cur = f_tv(t, f_mixing_ratio(f_es(t), f_es(td), p_cur, 141));   // PF:839-843
        if (KACT & 1u) step6<false>(sb, it, crow.at(sb.m), pk_cur, x_cur, s.x_prv, b_cur, s.b_prv);
        if (KACT & 2u) step6<false>(ml, it, crow.at(ml.m), pk_cur, x_cur, s.x_prv, b_cur, s.b_prv);
        if (KACT & 4u) step6<GUARD_MU>(mu, it, crow.at(mu.m), pk_cur, x_cur, s.x_prv, b_cur, s.b_prv);
        s.b_prv = b_cur; s.x_prv = x_cur;
        crow.advance();
    }
}

// The whole suite for one column, default options.  Interfaces as suite_column (xp_fast.cuh); `cf` must be
// the VIRTUAL-temperature table (compute_coef_tv).  Returns the mask of kinds for the exact path.
template <unsigned KINDS, class Rd, class Cf>
XP_HD unsigned suite_column6(const Rd &rd, const Cf &cf, const Prep &pr, const Tables &tb, const Opts &o,
                             FResult res[3]) {
    unsigned redo = 0;
    float nanacc = 0.0f;                   // becomes NaN if a T/Td read of the pre-pass is NaN or infinite
    const int nt = pr.n_table;
    // ---- pre-pass over the lowest levels: mixed-layer means (float64) and most-unstable argmax ----
    double sum_th = 0.0, sum_w = 0.0;
    float best = -1e30f, second = -1e30f, mu_t = 0.0f, mu_td = 0.0f;
    int k_mu = 0;
    const int n_low = max((KINDS & 4u) ? pr.K_mu : 0, (KINDS & 2u) ? pr.n_ml_w : 0);
    const float *tp0 = rd.tptr(0), *tdp0 = rd.tdptr(0);
    const int64_t ls = rd.stride();
    constexpr int kPre = 4;
    float tq[kPre], tdq[kPre], tn[kPre], tdn[kPre];
#pragma unroll
    for (int j = 0; j < kPre; ++j) {
        tn[j] = tdn[j] = 0.0f;
        if (j < n_low) { tn[j] = Rd::ld(tp0 + (int64_t)j * ls); tdn[j] = Rd::ld(tdp0 + (int64_t)j * ls); }
    }
    for (int k0 = 0; k0 < n_low; k0 += kPre) {
#pragma unroll
        for (int j = 0; j < kPre; ++j) { tq[j] = tn[j]; tdq[j] = tdn[j]; }
        tp0 += (int64_t)kPre * ls; tdp0 += (int64_t)kPre * ls;
#pragma unroll
        for (int j = 0; j < kPre; ++j)
            if (k0 + kPre + j < n_low) { tn[j] = Rd::ld(tp0 + (int64_t)j * ls); tdn[j] = Rd::ld(tdp0 + (int64_t)j * ls); }
#pragma unroll
        for (int j = 0; j < kPre; ++j) {
            const int k = k0 + j;
            if (k >= n_low) break;
            const float t = tq[j], td = tdq[j];
            nanacc = f_fma(t, 0.0f, f_fma(td, 0.0f, nanacc));
            const float p = pr.p[k];
            const float e = f_es(td);
            const float ipe = f_rcp(p - e);
            const float r = kEpsF * e * ipe;                 // saturation mixing ratio of the dewpoint (PF:258)
            if ((KINDS & 2u) && k < pr.n_ml_w) {
                // mixed_parcel PF:253-258 in float64 (see suite_column)
                const double e64 = sat_vapor_pressure((double)td);
                sum_th += pr.mlw[k] * ((double)t * pr.thfac[k]);
                sum_w += pr.mlw[k] * (kEps * e64 / (pr.p64[k] - e64));
            }
            if ((KINDS & 4u) && k < pr.K_mu) {
                // ln(theta_e), Bolton (1980) eq. 39 as in metpy.calc.equivalent_potential_temperature (PF:123)
                const float l2t = f_lg2(t), l2td = f_lg2(td);
                const float t_l = 56.0f + f_rcp(f_rcp(td - 56.0f) + (l2t - l2td) * (kLn2 / 800.0f));
                const float it_l = f_rcp(t_l);
                float v = l2t * kLn2;                                                   // ln T
                v = f_fma((float)kKappa * kLn2, f_lg2(1000.0f * ipe), v);               // + kappa ln(1000/(p-e))
                v = f_fma(0.28f * r * kLn2, l2t - f_lg2(t_l), v);                       // + 0.28 r ln(T/t_l)
                v = f_fma(r * f_fma(0.448f, r, 1.0f), f_fma(3036.0f, it_l, -1.78f), v);
                nanacc = f_fma(v, 0.0f, nanacc);
                if (v > best) { second = best; best = v; k_mu = k; mu_t = t; mu_td = td; }   // ties: larger p (PF:128)
                else if (v > second) second = v;
            }
        }
    }
    // ---- parcels --------------------------------------------------------------------------------
    FParcel sb, ml, mu;
    const float t_sfc = rd.T(0), td_sfc = rd.Td(0);
    nanacc = f_fma(t_sfc, 0.0f, f_fma(td_sfc, 0.0f, nanacc));
    Opts od = o; od.vtc = 1; od.compat = 141; od.pos_neg = 1;
    if (KINDS & 1u) {
        setup_parcel(rd, pr, tb, od, pr.p0, (double)t_sfc, (double)td_sfc, 0, 1, false, sb);
        res[0].par_p = pr.p[0]; res[0].par_t = t_sfc; res[0].par_td = td_sfc; res[0].shift = 0;
    }
    if (KINDS & 2u) {
        const double mp_t = sum_th * pr.exner0;                                  // PF:268-269
        const double mp_td = dewpoint_from_e(vapor_pressure(pr.p0, sum_w));      // PF:275-282
        setup_parcel(rd, pr, tb, od, pr.p0, mp_t, mp_td, 0, pr.K_ml, false, ml);
        res[1].par_p = pr.p[0]; res[1].par_t = (float)mp_t; res[1].par_td = (float)mp_td; res[1].shift = pr.K_ml;
    }
    if (KINDS & 4u) {
        if (!(best - second >= kThetaEMargin)) redo |= 4u;                       // argmax within float32 error
        setup_parcel(rd, pr, tb, od, pr.p64[k_mu], (double)mu_t, (double)mu_td, k_mu, k_mu + 1, false, mu);
        res[2].par_p = pr.p[k_mu]; res[2].par_t = mu_t; res[2].par_td = mu_td; res[2].shift = k_mu;
    }
    // v6 state: the LCL row as a difference, zeroed sums, start row of each parcel
    if (KINDS & 1u) { sb.b_lcl = sb.a_lcl - sb.b_lcl; sweep_init6(sb, pr.lnp[0]); }
    if (KINDS & 2u) { ml.b_lcl = ml.a_lcl - ml.b_lcl; sweep_init6(ml, pr.lnp[0]); }
    if (KINDS & 4u) { mu.b_lcl = mu.a_lcl - mu.b_lcl; sweep_init6(mu, pr.lnp[0]); }
    // ---- the sweep --------------------------------------------------------------------------------------
    Sweep6<Rd> s;
    s.lp_p = pr.p + 1; s.lp_x = pr.lnp + 1; s.lp_k = pr.pk + 1;
    s.tp = rd.tptr(1); s.tdp = rd.tdptr(1); s.ls = ls;
    s.t_nxt = Rd::ld(s.tp); s.td_nxt = Rd::ld(s.tdp);
    s.b_prv = 0.0f; s.x_prv = pr.lnp[0];
    auto crow = cf.row(0);
    // segment bounds: the mixed-layer parcel joins at K_ml; most-unstable parcels have all started by K_mu
    const int it_a = (KINDS & 2u) ? min(pr.K_ml, nt) : 1;
    const int it_b = (KINDS & 4u) ? max(it_a, min(pr.K_mu, nt)) : it_a;
    sweep_segment6<KINDS & 5u, true>(s, crow, 1, it_a, nt, sb, ml, mu);
    sweep_segment6<KINDS, true>(s, crow, it_a, it_b, nt, sb, ml, mu);
    sweep_segment6<KINDS, false>(s, crow, it_b, nt, nt, sb, ml, mu);
    // last iteration: no level `nt`; every parcel that is not bound for the exact path is above its LCL
    {
        const float big = 1e30f;
        if (KINDS & 1u) step6<false>(sb, nt, crow.at(sb.m), 0.0f, s.x_prv, s.x_prv, big, s.b_prv);
        if (KINDS & 2u) step6<false>(ml, nt, crow.at(ml.m), 0.0f, s.x_prv, s.x_prv, big, s.b_prv);
        if (KINDS & 4u) step6<false>(mu, nt, crow.at(mu.m), 0.0f, s.x_prv, s.x_prv, big, s.b_prv);
    }
    // ---- results ---------------------------------------------------------------------------------------
    // A NaN/Inf T or Td in the pre-pass levels poisons nanacc; one in the swept levels makes the area
    // sums of every parcel that sweeps it non-finite (each parcel sweeps every level above its start).
    bool nan_seen = !(nanacc == 0.0f);
    if (KINDS & 1u) nan_seen = nan_seen || !(sb.pos - sb.tot < 3e38f);
    if (KINDS & 2u) nan_seen = nan_seen || !(ml.pos - ml.tot < 3e38f);
    if (KINDS & 4u) nan_seen = nan_seen || !(mu.pos - mu.tot < 3e38f);
    auto wrap = [&](const FParcel &c, FResult &r, unsigned bit) {
        sweep_finish6(c, cf, pr, o, r);
        const bool unc = !(c.min_abs_d >= kDecisionEps) || !(c.min_slope >= 0.0f);
        if (c.bad || unc || nan_seen) redo |= bit;
    };
    if (KINDS & 1u) wrap(sb, res[0], 1u);
    if (KINDS & 2u) wrap(ml, res[1], 2u);
    if (KINDS & 4u) wrap(mu, res[2], 4u);
    // The most-unstable parcel is (certainly) the surface parcel: its exact recomputation is the
    // surface-based one -- tell the fix-up to do it once and write both (bit 3 replaces bit 2).
    if ((KINDS & 5u) == 5u && (redo & 4u) && k_mu == 0 && !nan_seen && (best - second >= kThetaEMargin))
        redo = (redo & ~4u) | 1u | kRedoMuIsSb;
    return redo;
}

}  // namespace fast
}  // namespace xp
