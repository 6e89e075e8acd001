// xp_fast_pcol7.cuh -- the v7 sweep (xp_fast7.cuh) for columns with PER-COLUMN pressure, on the shared-memory table of
// the adiabat family (fast::PTabView, xp_fast_pcol.cuh).  Default options, scalar outputs: BASELINE configs[1], [2]
// and the model-level suite.
//
// The generic per-column sweep (parcel_iteration_pcol + sweep_step) costs ~150 instructions per (level, parcel): the
// crossing position and temperature are evaluated at every row, every option is a run-time select, the adiabat is two
// dependent gathers from L2.  Here a row is a v7 step -- mixed phase until every lane of the warp is past its LCL and its
// guard rows, then the above-phase step -- the adiabat comes from four LDS.128 + 16 FFMA, and LFC / EL positions are
// rebuilt once after the sweep.  ln p and p^kappa are per thread (2 MUFU per level), the interval half-width of the
// above phase is per lane.  Pre-pass and parcel set-up are suite_column_pcol's.
#pragma once
#include <type_traits>

#include "xp_fast7.cuh"

namespace xp {
namespace fast {

// ln p and parcel virtual temperature of the crossing remembered as v = iteration + fraction (iteration > ka).
template <class Rd, class PTab>
XP_HD void crossing_ptab7(const PColParcel &c, const Rd &rd, const PTab &ptab, int L, float v, float &x, float &y) {
    const int itc = min(max((int)v, 1), L);
    const float fr = v - (float)itc;
    const int kc = itc - 1;                                      // level of the upper row
    const float l1 = f_lg2(rd.P(kc));
    const float x1 = kLn2 * l1;
    const float a1 = ptab.eval(ptab.level(f_ex2((float)kKappa * l1)), c.m, c.f);
    float x0 = c.x_lcl, a0 = c.lcl_tv;                           // lower row: the LCL row ...
    if (itc != c.ka + 1) {                                       // ... or level kc-1
        const float l0 = f_lg2(rd.P(max(kc - 1, 0)));
        x0 = kLn2 * l0;
        a0 = ptab.eval(ptab.level(f_ex2((float)kKappa * l0)), c.m, c.f);
    }
    x = f_fma(-fr, x0 - x1, x0);
    y = f_fma(fr, a1 - a0, a0);
}

// lfc_el PF:1140-1185 + cape_cin_base PF:1329-1388 on the v7 state (as sweep_finish7).
template <class Rd, class PTab>
XP_HD void finish_ptab7(const PColParcel &s, const Rd &rd, const PTab &ptab, int L, const Opts &o, FResult &r) {
    const bool lfc_found = s.lfc_it != 0;
    const int el_it = (int)s.el_x;
    const bool top_colder = s.dprev <= 0.0f;                            // PF:1151
    const bool el_exists = top_colder && el_it > s.ka;                  // PF:1152-1153
    const bool lfc_missing = s.n_inc == 0 && !lfc_found;                // PF:1161
    const bool pos_parcel = s.max_d_above > 0.0f;
    const bool replace = (pos_parcel && lfc_missing) || (!lfc_missing && !lfc_found && el_exists);
    const bool have_lfc = lfc_found || replace;
    r.lfc_p = r.lfc_t = r.el_p = r.el_t = f_qnan();
    if (lfc_found) {
        float x, y;
        crossing_ptab7(s, rd, ptab, L, s.lfc_x, x, y);
        r.lfc_p = f_ex2(x * kLog2e); r.lfc_t = y;
    }
    if (replace) { r.lfc_p = s.lcl_p; r.lfc_t = s.lcl_tv; }
    if (el_exists) {
        float x, y;
        crossing_ptab7(s, rd, ptab, L, s.el_x, x, y);
        r.el_p = f_ex2(x * kLog2e); r.el_t = y;
    }
    float cape = 0.0f, cin = 0.0f;
    if (have_lfc) {
        const float l_P = replace ? s.lcl_pos : 0.0f;                   // P restarted at the LFC crossing
        const float l_N = replace ? s.lcl_tot : s.tot;
        const float e_P = el_exists ? s.el_pos : s.pos;
        const bool el_below_lfc = el_exists && lfc_found && !replace && s.el_x < s.lfc_x;
        cin = l_N;
        cape = el_below_lfc ? 0.0f : (e_P - l_P);
    }
    cape *= (float)kRd; cin *= (float)kRd;
    if (o.post_zero && !(cin <= 0.0f)) cin = 0.0f;
    r.cape = cape; r.cin = cin;
    r.lcl_p = s.lcl_p; r.lcl_t = s.lcl_t; r.lcl_tv = s.lcl_tv;
}

template <unsigned KINDS, class Rd, class PTab>
XP_HD unsigned ptab7_sweep(const Rd &rd, int L, const Opts &o, const PTab &ptab, PColParcel &sb, PColParcel &ml,
                           PColParcel &mu, FResult res[3], unsigned redo, float nanacc, bool bad_axis, float best,
                           float second, int k_mu, int qm, float p_sfc, float t_sfc, float td_sfc, float x_sfc) {
    // v7 state: the LCL row as a difference, P / N / crossing bookkeeping of sweep_init7
    auto init = [&](PColParcel &c, float x0) {
        const float d_lcl = c.a_lcl - c.b_lcl;
        sweep_init7(c, x0);
        c.b_lcl = d_lcl;
    };
    if (KINDS & 1u) init(sb, x_sfc);
    if (KINDS & 2u) init(ml, x_sfc);
    if (KINDS & 4u) init(mu, kLn2 * f_lg2(res[2].par_p));
    // last iteration at which this lane still needs the mixed step: its LCL rows and its guard rows
    int lim = 1;
    if (KINDS & 1u) lim = max(lim, sb.bad ? 0 : sb.ka + 1);
    if (KINDS & 2u) lim = max(lim, max(ml.kfirst, ml.bad ? 0 : ml.ka + 1));
    if (KINDS & 4u) lim = max(lim, max(mu.kfirst, mu.bad ? 0 : mu.ka + 1));
    const int it_abv = min(XP_WARP_MAX_INT(lim) + 1, L + 1);     // mixed phase [1, it_abv), above phase [it_abv, L]
    float b_prv = 0.0f, x_prv = x_sfc, p_prv = p_sfc, h_prv = 0.0f, itf = 1.0f;
    PTabLevel lv_prv = ptab.level(f_ex2((float)kKappa * f_lg2(p_sfc)));
    const float *ppp = rd.pptr(min(1, L - 1)), *tp = rd.tptr(min(1, L - 1)), *tdp = rd.tdptr(min(1, L - 1));
    const int64_t ls = rd.stride(), pls = rd.pstride();
    // levels arrive two iterations ahead (r2p profile of this kernel: 7 % of its time waiting on a one-level-ahead load)
    float p_nxt = Rd::ld(ppp), t_nxt = Rd::ld(tp), td_nxt = Rd::ld(tdp);
    float p_n2 = 0.0f, t_n2 = 0.0f, td_n2 = 0.0f;
    ppp += pls; tp += ls; tdp += ls;
    if (2 < L) { p_n2 = Rd::ld(ppp); t_n2 = Rd::ld(tp); td_n2 = Rd::ld(tdp); }
    for (int it = 1; it <= L; ++it) {
        const bool last = (it == L);
        const float p_cur0 = p_nxt, t = t_nxt;
        float td = td_nxt;
        if (qm && !last) td = f_td_from_q(p_cur0, t, td, qm);
        p_nxt = p_n2; t_nxt = t_n2; td_nxt = td_n2;
        ppp += pls; tp += ls; tdp += ls;
        if (it + 2 < L) { p_n2 = Rd::ld(ppp); t_n2 = Rd::ld(tp); td_n2 = Rd::ld(tdp); }
        float b_cur = 1e30f, x_cur = x_prv, pk_cur = 0.0f, p_cur = p_prv;
        PTabLevel lv_cur = lv_prv;
        if (!last) {
            nanacc = f_fma(p_cur0, 0.0f, f_fma(t, 0.0f, f_fma(td, 0.0f, nanacc)));
            p_cur = p_cur0;
            if (!(p_cur < p_prv) || !(p_cur >= 2.5f)) bad_axis = true;
            const float l2p = f_lg2(p_cur);
            x_cur = kLn2 * l2p; pk_cur = f_ex2((float)kKappa * l2p);
            b_cur = f_env_tv7(t, td, p_cur);                                        // PF:839-843
            lv_cur = ptab.level(pk_cur);
        }
        if (it < it_abv) {
            // mixed phase (row schedule of FParcel): dry adiabat at level it, moist adiabat at level it-1
            auto row = [&](PColParcel &c, auto guard) {
                const float d_m = ptab.eval(lv_prv, c.m, c.f) - b_prv;
                const float d_d = last ? -1e30f : f_fma(c.c_dryv, pk_cur, -b_cur);
                step7_mixed<decltype(guard)::value>(c, it, itf, d_m, d_d, x_cur, x_prv);
                // rows read from the coarser parts of the table are decided with a wider margin
                if (it > c.ka) c.min_abs_d = fminf(c.min_abs_d, fabsf(d_m) * lv_prv.margin_scale);
            };
            if (KINDS & 1u) row(sb, std::integral_constant<int, 0>());
            if (KINDS & 2u) row(ml, std::integral_constant<int, 2>());
            if (KINDS & 4u) row(mu, std::integral_constant<int, 1>());
        } else {
            auto row = [&](PColParcel &c) {
                const float d = ptab.eval(lv_prv, c.m, c.f) - b_prv;
                step7_above(c, itf, d, h_prv);
                c.min_abs_d = fminf(c.min_abs_d, fabsf(d) * lv_prv.margin_scale);
            };
            if (KINDS & 1u) row(sb);
            if (KINDS & 2u) row(ml);
            if (KINDS & 4u) row(mu);
        }
        h_prv = 0.5f * (x_prv - x_cur);
        b_prv = b_cur; x_prv = x_cur; p_prv = p_cur; lv_prv = lv_cur;
        itf += 1.0f;
    }
    // ---- results (as suite_column7) -------------------------------------------------------------------------
    bool nan_seen = !(nanacc == 0.0f) || bad_axis;
    if (KINDS & 1u) nan_seen = nan_seen || !(sb.pos - sb.tot < 3e38f);
    if (KINDS & 2u) nan_seen = nan_seen || !(ml.pos - ml.tot < 3e38f);
    if (KINDS & 4u) nan_seen = nan_seen || !(mu.pos - mu.tot < 3e38f);
    auto wrap = [&](const PColParcel &c, FResult &r, unsigned bit) {
        finish_ptab7(c, rd, ptab, L, o, r);
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
