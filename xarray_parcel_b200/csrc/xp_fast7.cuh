// xp_fast7.cuh -- sweep version 7 of the float32 fast path on a shared pressure axis (default options, as
// xp_fast6.cuh: virtual temperature correction PF:1394-1475, MetPy 1.4.1 formulas, pos_cape_neg_cin PF:1329-1388).
//
// Same decisions, same hand-over rules and same parcel set-up as v6; what changes is the instruction count of the
// sweep (the r1d profile: 8.9 k warp instructions per 32 columns, 38 % of them select / compare / integer work on
// the half-rate ALU pipe):
//   * PHASES.  The row schedule of FParcel (level `it` below the LCL, the LCL row at it == ka, level it-1 above)
//     costs selects on every row.  Once EVERY lane of the warp is past the first row above its LCL
//     (it >= max ka + 2, one REDUX per column), all lanes process level it-1 on the moist adiabat over the shared
//     interval (it-2, it-1): the "above" step has no row selects, no per-parcel ln p state and takes the
//     half-width of the interval from the packed axis constants.
//   * LFC bookkeeping without snapshots.  CIN = N(LFC) and CAPE = P(EL) - P(LFC) (PF:1329-1388): instead of
//     snapshotting P and N at the first increasing crossing above the LCL, N simply stops accumulating there (a
//     predicated add) and P restarts from zero (a predicated move) -- two selects and two registers less per
//     parcel; "found" replaces the lfc_it == 0 compare.
//   * a crossing is remembered as ONE float, iteration + fraction (the fraction keeps 18 bits: 4e-6 of the
//     interval), instead of an integer and a float: one select per LFC / EL candidate instead of two.
//   * the shallow-crossing guard (|d0 - d1| < kCrossSlope dx) is a sticky predicate instead of a running minimum;
//     "any increasing crossing" is only tracked below the first row above the LCL (above it, an increasing crossing
//     either is the LFC or comes after it); the negative part of an interval is S - positive part.
//   * Bolton's es with the constant 6.112 folded into the exponent (one multiply less, twice per level).
#pragma once
#include "xp_fast6.cuh"
#include "xp_fast_pcol.cuh"

namespace xp {
namespace fast {

// XP_WARP_MAX_INT (warp-wide integer maximum; on the host a settable floor that stands in for the other lanes of a
// warp -- tests/hostsim) is defined in xp_fast_pcol.cuh.

// es(T) = 6.112 * 2^(kEsC (1 - 243.5/(T - 29.65))) = 2^(kEsA - kEsB/(T - 29.65))
constexpr float kEsA = 17.67f * 1.4426950408889634f + 2.611644f;   // + log2(6.112) = 2.6116443...
constexpr float kEsB = 243.5f * 17.67f * 1.4426950408889634f;
XP_HD float f_es7(float t) { return f_ex2(f_fma(-kEsB, f_rcp(t - 29.65f), kEsA)); }
// environment virtual temperature, PF:839-843 with the MetPy 1.4.1 mixing ratio eps es(Td) / (p - es(T))
XP_HD float f_env_tv7(float t, float td, float p) {
    const float w608 = (0.608f * kEpsF) * f_es7(td) * f_rcp(p - f_es7(t));
    return f_fma(t, w608, t);
}

// v7 meanings of the FParcel fields (the rest as v6):
//   pos      P: positive area since the start row, restarted from zero at the LFC crossing
//   tot      N: negative area, frozen once the LFC is found (= CIN / Rd)
//   lfc_x    iteration + fraction of the LFC crossing (0: none)        lfc_it   != 0: LFC found
//   el_x     iteration + fraction of the last decreasing crossing (0: none)
//   el_pos   P including the triangle below that crossing
//   n_inc    != 0: an increasing crossing was seen in a row that is not above the LCL
//   min_slope < 0: a shallow crossing was seen
XP_HD void sweep_init7(FParcel &c, float x0) {
    c.xprev = x0; c.dprev = 0.0f;
    c.pos = c.tot = c.lcl_pos = c.lcl_tot = 0.0f;
    c.lfc_x = 0.0f; c.el_pos = c.el_x = 0.0f;
    c.lfc_it = 0; c.n_inc = 0;
    c.min_abs_d = 1e30f; c.min_slope = 1.0f; c.max_d_above = -1e30f;
}

// The part of a row shared by both phases: the interval (dprev -> d) of half-width h (dx = 2 h), `itf` = float(it).
// ABOVE: every lane is above its LCL (compile time); otherwise `above` says so per lane.
template <bool ABOVE>
XP_HD void step7_core(FParcel &c, float itf, float d, float h, bool above) {
    const float t = c.dprev * h;
    const float S = f_fma(d, h, t);                               // whole trapezoid (PF:164-206)
    const float den = c.dprev - d;
    const bool cross = c.dprev * d < 0.0f;                        // PF:1026-1031
    const float fr = c.dprev * f_rcp(den);                        // zero at xprev - fr dx
    const float u = cross ? t * fr : S;                           // lower triangle (PF:1200-1289) or everything
    const float ahi = S - u;
    const float pinc = fmaxf(u, ahi), ninc = S - pinc;
    const bool dpos = d > 0.0f;
    const bool inc = cross && dpos;                               // PF:1058
    const bool dec = cross && !dpos;                              // PF:1060
    const bool found = c.lfc_it != 0;
    // LFC: max-pressure increasing crossing above the LCL (PF:1127-1132) = the first one met
    const bool take = ABOVE ? (inc && !found) : (inc && above && !found);
    if (!found) c.tot += ninc;                                    // includes the triangle below the LFC
    c.pos = take ? 0.0f : c.pos;
    c.pos += pinc;                                                // the triangle above the LFC starts P
    const float v = itf + fr;
    c.lfc_x = take ? v : c.lfc_x;
    c.lfc_it = take ? 1 : c.lfc_it;
    // EL: min-pressure decreasing crossing (PF:1136) = the last one met
    c.el_x = dec ? v : c.el_x;
    c.el_pos = dec ? c.pos : c.el_pos;
    if (!ABOVE) c.n_inc = (inc && !above) ? 1 : c.n_inc;          // PF:1161 only asks "any"
    c.max_d_above = ABOVE ? fmaxf(c.max_d_above, d) : fmaxf(c.max_d_above, above ? d : -1e30f);   // PF:1166-1169
    c.min_slope = (cross && fabsf(den) < (2.0f * kCrossSlope) * h) ? -1.0f : c.min_slope;
    c.dprev = d;
}

// Mixed phase: lanes may be below, at or above their LCL (row schedule of FParcel).
//   d_m   parcel - environment on the moist adiabat at level it-1 (PF:585-592)
//   d_d   parcel - environment on the dry adiabat at level it     (PF:742)
// GUARD 0: the parcel is live at every row.  1: rows it < kfirst are neutral and leave xprev at the row's ln p
// (most-unstable parcel: its start row is the level below kfirst).  2: rows it < kfirst are neutral and leave xprev
// untouched (mixed-layer parcel on per-column pressure: its start row is the surface, the levels inside the layer are
// dropped, PF:1636).
template <int GUARD>
XP_HD void step7_mixed(FParcel &c, int it, float itf, float d_m, float d_d, float x_cur, float x_prv) {
    const bool above = it > c.ka, is_lcl = it == c.ka;
    float d = above ? d_m : d_d;
    d = is_lcl ? c.b_lcl : d;
    float x = above ? x_prv : x_cur;
    x = is_lcl ? c.x_lcl : x;
    float absd = fabsf(d);
    if (GUARD) {
        const bool active = it >= c.kfirst;
        d = active ? d : 0.0f; absd = active ? absd : 1e30f;
        if (GUARD == 2) x = active ? x : c.xprev;
    }
    const float h = 0.5f * (c.xprev - x);
    c.min_abs_d = fminf(c.min_abs_d, absd);
    step7_core<false>(c, itf, d, h, above);
    c.lcl_pos = is_lcl ? c.pos : c.lcl_pos; c.lcl_tot = is_lcl ? c.tot : c.lcl_tot;
    c.xprev = x;
}

// Above phase: every lane processes level it-1 on its moist adiabat over the interval (it-2, it-1) of half-width h.
XP_HD void step7_above(FParcel &c, float itf, float d, float h) {
    c.min_abs_d = fminf(c.min_abs_d, fabsf(d));
    step7_core<true>(c, itf, d, h, true);
}

// ln p and parcel virtual temperature of the crossing remembered as v = iteration + fraction (iteration > ka).
template <class Cf>
XP_HD void crossing7(const FParcel &c, const Cf &cf, const Prep &pr, float v, int &itc, float &x, float &y) {
    // (a fraction that rounds up to 1 would name the next iteration; such a crossing sits on a level to 2e-6 of the
    //  interval, |d| there is far below kDecisionEps and the column is redone -- only keep the table index in range)
    itc = min((int)v, pr.n_table);
    const float fr = v - (float)itc;
    const int kc = itc - 1;                                      // level of the upper row
    const float x1 = pr.lnp[kc];
    const float a1 = cubic_at(cf.row(kc).at(c.m), c.f);
    float x0 = c.x_lcl, a0 = c.lcl_tv;                           // lower row: the LCL row ...
    if (itc != c.ka + 1) { x0 = pr.lnp[kc - 1]; a0 = cubic_at(cf.row(kc - 1).at(c.m), c.f); }   // ... or level kc-1
    x = f_fma(-fr, x0 - x1, x0);
    y = f_fma(fr, a1 - a0, a0);
}

// lfc_el PF:1140-1185 + cape_cin_base PF:1329-1388 on the v7 state.
template <class Cf>
XP_HD void sweep_finish7(const FParcel &s, const Cf &cf, const Prep &pr, const Opts &o, FResult &r) {
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
        float x, y; int itc;
        crossing7(s, cf, pr, s.lfc_x, itc, x, y);
        r.lfc_p = f_ex2(x * kLog2e); r.lfc_t = y;
    }
    if (replace) { r.lfc_p = s.lcl_p; r.lfc_t = s.lcl_tv; }
    if (el_exists) {
        float x, y; int itc;
        crossing7(s, cf, pr, s.el_x, itc, x, y);
        r.el_p = f_ex2(x * kLog2e); r.el_t = y;
    }
    float cape = 0.0f, cin = 0.0f;
    if (have_lfc) {
        // P restarted at the LFC crossing; with the LFC at the LCL (replace) it did not restart
        const float l_P = replace ? s.lcl_pos : 0.0f;
        const float l_N = replace ? s.lcl_tot : s.tot;
        const float e_P = el_exists ? s.el_pos : s.pos;
        // EL below the LFC (PF:1352-1353 leaves no level between them)
        const bool el_below_lfc = el_exists && lfc_found && !replace && s.el_x < s.lfc_x;
        cin = l_N;
        cape = el_below_lfc ? 0.0f : (e_P - l_P);
    }
    cape *= (float)kRd; cin *= (float)kRd;
    if (o.post_zero && !(cin <= 0.0f)) cin = 0.0f;
    r.cape = cape; r.cin = cin;
    r.lcl_p = s.lcl_p; r.lcl_t = s.lcl_t; r.lcl_tv = s.lcl_tv;
}

// Shared per-level state of the v7 sweep (the T/Td pipeline of Sweep6 plus the half-width of the last interval).
struct Sweep7 {
    Sweep6 s;                             // t_n1 / td_n1: the next global level; off / k_pf: the level after t_n2
    float t_n2, td_n2;                    // the level after that (two levels in flight: the r2a profile had 30 % of
                                          // the above-phase loop waiting on a one-level-ahead load)
    float h_prv;                          // half-width of the interval (it-2, it-1)
    float itf;                            // float(it)
};

// One level of T/Td: from the stash or from the global prefetch pipeline (as sweep_segment6).
template <class Rd, class Stash>
XP_HD void next_level7(const Rd &rd, Sweep7 &w, const Stash &stash, bool from_stash, int it, int nt, float &t, float &td) {
    Sweep6 &s = w.s;
    if (from_stash) {
        stash.get(it, t, td);
    } else {
        t = s.t_n1; td = s.td_n1;
        s.t_n1 = w.t_n2; s.td_n1 = w.td_n2;
        if (s.k_pf < nt) { w.t_n2 = rd.ldT(s.off); w.td_n2 = rd.ldTd(s.off); }    // two levels ahead
        if (s.k_pf + kL2Ahead < nt) rd.prefetch(s.off + kL2Ahead * s.ls);
        s.off += s.ls; ++s.k_pf;
    }
}

// Mixed-phase iterations [it0, it1) for the parcels in KACT.
template <unsigned KACT, bool GUARD_MU, class Rd, class CoefRow, class Stash>
XP_HD void sweep_mixed7(const Rd &rd, Sweep7 &w, CoefRow &crow, const Stash &stash, bool from_stash, int it0, int it1, int nt,
                        int qmode, FParcel &sb, FParcel &ml, FParcel &mu) {
    Sweep6 &s = w.s;
    for (int it = it0; it < it1; ++it) {
        float t, td;
        next_level7(rd, w, stash, from_stash, it, nt, t, td);
        const float p_cur = s.lp[0], x_cur = s.lp[1], pk_cur = s.lp[2];
        w.h_prv = s.lp[3];
        s.lp += 4;
        if (qmode && !from_stash) td = f_td_from_q(p_cur, t, td, qmode);              // the stash holds dewpoints already
        const float b_cur = f_env_tv7(t, td, p_cur);                                  // PF:839-843
        if (KACT & 1u) step7_mixed<0>(sb, it, w.itf, cubic_at(crow.at(sb.m), sb.f) - s.b_prv, f_fma(sb.c_dryv, pk_cur, -b_cur), x_cur, s.x_prv);
        if (KACT & 2u) step7_mixed<0>(ml, it, w.itf, cubic_at(crow.at(ml.m), ml.f) - s.b_prv, f_fma(ml.c_dryv, pk_cur, -b_cur), x_cur, s.x_prv);
        if (KACT & 4u) step7_mixed<GUARD_MU ? 1 : 0>(mu, it, w.itf, cubic_at(crow.at(mu.m), mu.f) - s.b_prv, f_fma(mu.c_dryv, pk_cur, -b_cur), x_cur, s.x_prv);
        s.b_prv = b_cur; s.x_prv = x_cur;
        w.itf += 1.0f;
        crow.advance();
    }
}

// Above-phase iterations [it0, it1): every parcel of every lane is past the first row above its LCL.  With
// TOP the warp checks after every iteration whether it can stop (see sweep_top6); returns true if it did.
template <unsigned KINDS, bool TOP, class Rd, class CoefRow>
XP_HD bool sweep_above7(const Rd &rd, Sweep7 &w, CoefRow &crow, int it0, int it1, int nt, float stop_below,
                        int qmode, FParcel &sb, FParcel &ml, FParcel &mu) {
    Sweep6 &s = w.s;
    NoStash ns;
    // (unrolling this loop by two -- so that the load pipeline rotates through registers without moves -- was
    //  measured slower: 1.352 vs 1.277 ms per step)
    for (int it = it0; it < it1; ++it) {
        float t, td;
        next_level7(rd, w, ns, false, it, nt, t, td);
        const float p_cur = s.lp[0], h_cur = s.lp[3];
        s.lp += 4;
        if (qmode) td = f_td_from_q(p_cur, t, td, qmode);
        const float b_cur = f_env_tv7(t, td, p_cur);                                  // PF:839-843
        const float h = w.h_prv, b = s.b_prv;
        if (KINDS & 1u) step7_above(sb, w.itf, cubic_at(crow.at(sb.m), sb.f) - b, h);
        if (KINDS & 2u) step7_above(ml, w.itf, cubic_at(crow.at(ml.m), ml.f) - b, h);
        if (KINDS & 4u) step7_above(mu, w.itf, cubic_at(crow.at(mu.m), mu.f) - b, h);
        s.b_prv = b_cur; w.h_prv = h_cur;
        w.itf += 1.0f;
        crow.advance();
        if (TOP) {
            // the row just processed is level it-1: the parcel's curve there is dprev + b
            bool done = true;
            if (KINDS & 1u) done = done && (sb.bad || (sb.dprev < 0.0f && sb.dprev + b < stop_below));
            if (KINDS & 2u) done = done && (ml.bad || (ml.dprev < 0.0f && ml.dprev + b < stop_below));
            if (KINDS & 4u) done = done && (mu.bad || (mu.dprev < 0.0f && mu.dprev + b < stop_below));
            if (XP_WARP_ALL(done)) return true;
        }
    }
    return false;
}

// The whole suite for one column, default options.  Interfaces as suite_column6; `pr.plk[k][3]` must hold the
// half-width 0.5 (ln p[k-1] - ln p[k]) of the interval below level k.
// QIN: the dewpoint array may hold specific humidity (o.qmode says in which MetPy form); false compiles the conversion away.
template <unsigned KINDS, bool QIN, class Rd, class Cf, class Stash>
XP_HD unsigned suite_column7(const Rd &rd, const Cf &cf, const Prep &pr, const Tables &tb, const Opts &o,
                             Stash &stash, FResult res[3]) {
    unsigned redo = 0;
    float nanacc = 0.0f;                   // becomes NaN if a T/Td read of the pre-pass is NaN or infinite
    const int nt = pr.n_table;
    const int n_low = max(1, max((KINDS & 4u) ? pr.K_mu : 0, (KINDS & 2u) ? pr.n_ml_w : 0));
    const uint32_t ls = rd.ls();
    const int n_stash = (stash.capacity() >= n_low) ? n_low : 0;
    // != 0: the dewpoint array holds specific humidity (xp_columns.dewpoint_is_specific_humidity): every level is
    // converted as it is loaded (float32), the parcels' dewpoints and the mixed-layer vapour pressures in float64
    const int qm = QIN ? o.qmode : 0;
    Sweep7 w;
    Sweep6 &s = w.s;
    {
        const int k0 = (n_stash > 0) ? n_stash : 1;          // first level the sweep reads from global memory
        s.off = rd.off0() + (uint32_t)k0 * ls; s.ls = ls;
        s.t_n1 = s.td_n1 = w.t_n2 = w.td_n2 = 0.0f;
        if (k0 < nt) { s.t_n1 = rd.ldT(s.off); s.td_n1 = rd.ldTd(s.off); }
        if (k0 + 1 < nt) { w.t_n2 = rd.ldT(s.off + ls); w.td_n2 = rd.ldTd(s.off + ls); }
#pragma unroll
        for (int j = 2; j <= kL2Ahead + 1; ++j)
            if (k0 + j < nt) rd.prefetch(s.off + (uint32_t)j * ls);
        s.off += 2 * ls; s.k_pf = k0 + 2;
    }
    for (int k = pr.k_top; k < nt; ++k) rd.prefetch(rd.off0() + (uint32_t)k * ls);
    // ---- pre-pass over the lowest levels: mixed-layer means (float64) and most-unstable argmax (as v6) ----
    double sum_th = 0.0, sum_w = 0.0;
    float best = -1e30f, second = -1e30f, mu_t = 0.0f, mu_td = 0.0f;
    int k_mu = 0;
    uint32_t off0 = rd.off0();
    for (int k = 1; k < n_low; ++k) rd.prefetch(off0 + (uint32_t)k * ls);
    const float t_sfc = rd.ldT(off0), raw_sfc = rd.ldTd(off0);
    // float64 columns (Rd::kDouble): the SWEEP runs on the values rounded to float32 (1.5e-5 K, inside the decision
    // margins), but everything the table cell of a parcel's LCL hangs on -- the parcel's own T / Td and the
    // mixed-layer sums -- is read in float64, exactly as for float32 columns (whose values ARE their float64 values)
    double t_sfc64 = (double)t_sfc, raw_sfc64 = (double)raw_sfc;
    if (Rd::kDouble) { t_sfc64 = rd.ldT64(off0); raw_sfc64 = rd.ldTd64(off0); }
    float t_nx = t_sfc, td_nx = raw_sfc, mu_raw = raw_sfc;
    float t_n2 = 0.0f, td_n2 = 0.0f;                         // levels are read two iterations ahead
    off0 += ls;
    if (1 < n_low) { t_n2 = rd.ldT(off0); td_n2 = rd.ldTd(off0); }
#pragma unroll(kPrepassUnroll)
    for (int k = 0; k < n_low; ++k) {
        const float t = t_nx, raw = td_nx;
        t_nx = t_n2; td_nx = td_n2;
        off0 += ls;
        if (k + 2 < n_low) { t_n2 = rd.ldT(off0); td_n2 = rd.ldTd(off0); }
        const float p = pr.p[k];
        const float td = qm ? f_td_from_q(p, t, raw, qm) : raw;
        if (k < n_stash) stash.put(k, t, td);
        nanacc = f_fma(t, 0.0f, f_fma(td, 0.0f, nanacc));
        const float e = f_es7(td);
        const float ipe = f_rcp(p - e);
        const float r = kEpsF * e * ipe;                 // saturation mixing ratio of the dewpoint (PF:258)
        if ((KINDS & 2u) && k < pr.n_ml_w) {
            // mixed_parcel PF:253-258 in float64 (see suite_column)
            double t64 = (double)t, raw64 = (double)raw;
            if (Rd::kDouble) { const uint32_t ok_ = rd.off0() + (uint32_t)k * ls; t64 = rd.ldT64(ok_); raw64 = rd.ldTd64(ok_); }
            const double tdd = qm ? (double)td : raw64;
            const double e64 = qm ? e64_from_q_fast(pr.p64[k], t64, raw64, qm)
                                  : kSat0 * exp64_fast(17.67 * (tdd - 273.15) * rcp64(tdd - 29.65));
            sum_th += pr.mlw[k] * (t64 * pr.thfac[k]);
            sum_w += pr.mlw[k] * (kEps * e64 * rcp64(pr.p64[k] - e64));
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
            if (v > best) { second = best; best = v; k_mu = k; mu_t = t; mu_td = td; mu_raw = raw; }   // ties: larger p (PF:128)
            else if (v > second) second = v;
        }
    }
    // the surface dewpoint (pre-pass level 0) and the parcels' dewpoints in float64
    double td_sfc64, mu_td64;
    float td_sfc;
    if (qm) {
        td_sfc64 = td64_from_q_fast(pr.p0, t_sfc64, raw_sfc64, qm);
        td_sfc = (float)td_sfc64;
    } else {
        td_sfc = raw_sfc; td_sfc64 = raw_sfc64;
    }
    // the top of the column: coldest environment temperature above kTopCheckHpa (see suite_column6)
    float tmin_top = 1e30f, tmax_top = -1e30f;
    for (int k = pr.k_top; k < nt; k += 4) {               // four levels (eight loads) in flight per round trip
        float tq[4], dq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t o_ = rd.off0() + (uint32_t)min(k + j, nt - 1) * ls;
            tq[j] = rd.ldT(o_); dq[j] = rd.ldTd(o_);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            nanacc = f_fma(tq[j], 0.0f, f_fma(dq[j], 0.0f, nanacc));
            tmin_top = fminf(tmin_top, tq[j]);
            tmax_top = fmaxf(tmax_top, tq[j]);
        }
    }
    bool singular = false;
    if (nt > pr.k_top && !(f_es(tmax_top) < 0.5f * pr.p[nt - 1])) {
        tmin_top = -1e30f;
        for (int k = pr.k_top; k < nt; ++k) {
            const float pk = pr.p[k];
            if (fabsf(pk - f_es(rd.ldT(rd.off0() + (uint32_t)k * ls))) < 0.05f * pk) singular = true;
        }
    }
    // ---- parcels: staged so that the LCL solves, the gathers and their consumers overlap (as v6) --------------
    FParcel sb, ml, mu;
    Setup6 u_sb, u_ml, u_mu;
    auto lev = [&](int k, float &t, float &td) {
        if (k < n_stash) stash.get(k, t, td);
        else {
            const uint32_t o_ = rd.off0() + (uint32_t)k * ls; t = rd.ldT(o_); td = rd.ldTd(o_);
            if (qm) td = f_td_from_q(pr.p[k], t, td, qm);
        }
    };
    double mp_t = 0.0, mp_td = 0.0;
    if (KINDS & 1u) setup6_a(pr.p0, t_sfc64, td_sfc64, sb, u_sb);
    if (KINDS & 2u) {
        mp_t = sum_th * pr.exner0;                                               // PF:268-269
        {
            const double val = log64_fast(pr.p0 * sum_w * rcp64(kEps + sum_w) * (1.0 / kSat0));
            mp_td = 243.5 * val * rcp64(17.67 - val) + kZeroC;
        }
        setup6_a(pr.p0, mp_t, mp_td, ml, u_ml);
    }
    if (KINDS & 4u) {
        if (!(best - second >= kThetaEMargin)) redo |= 4u;                       // argmax within float32 error
        double mu_t64 = (double)mu_t, mu_raw64 = (double)mu_raw;
        if (Rd::kDouble) { const uint32_t om_ = rd.off0() + (uint32_t)k_mu * ls; mu_t64 = rd.ldT64(om_); mu_raw64 = rd.ldTd64(om_); }
        mu_td64 = qm ? td64_from_q_fast(pr.p64[k_mu], mu_t64, mu_raw64, qm) : mu_raw64;
        if (qm) mu_td = (float)mu_td64;
        setup6_a(pr.p64[k_mu], mu_t64, mu_td64, mu, u_mu);
    }
    if (KINDS & 1u) setup6_b(lev, pr, tb, 1, sb, u_sb);
    if (KINDS & 2u) setup6_b(lev, pr, tb, pr.K_ml, ml, u_ml);
    if (KINDS & 4u) setup6_b(lev, pr, tb, k_mu + 1, mu, u_mu);
    if (KINDS & 1u) {
        setup6_c(pr, o, 0, sb, u_sb);
        res[0].par_p = pr.p[0]; res[0].par_t = t_sfc; res[0].par_td = td_sfc; res[0].shift = 0;
    }
    if (KINDS & 2u) {
        setup6_c(pr, o, 0, ml, u_ml);
        res[1].par_p = pr.p[0]; res[1].par_t = (float)mp_t; res[1].par_td = (float)mp_td; res[1].shift = pr.K_ml;
    }
    if (KINDS & 4u) {
        setup6_c(pr, o, k_mu, mu, u_mu);
        res[2].par_p = pr.p[k_mu]; res[2].par_t = mu_t; res[2].par_td = mu_td; res[2].shift = k_mu;
    }
    if (KINDS & 1u) sweep_init7(sb, pr.lnp[0]);
    if (KINDS & 2u) sweep_init7(ml, pr.lnp[0]);
    if (KINDS & 4u) sweep_init7(mu, pr.lnp[k_mu]);
    // ---- the sweep --------------------------------------------------------------------------------------
    s.lp = &pr.plk[1][0];
    s.b_prv = 0.0f; s.x_prv = pr.lnp[0];
    w.h_prv = 0.0f; w.itf = 1.0f;
    auto crow = cf.row(0);
    // First iteration of the above phase: every lane of the warp is past the first row above its LCL.  Parcels
    // bound for the exact path (ka = n_table) do not hold the warp back: their rows are garbage either way.
    int ka_max = 0;
    if (KINDS & 1u) ka_max = max(ka_max, sb.bad ? 0 : sb.ka);
    if (KINDS & 2u) ka_max = max(ka_max, ml.bad ? 0 : ml.ka);
    if (KINDS & 4u) ka_max = max(ka_max, mu.bad ? 0 : mu.ka);
    ka_max = XP_WARP_MAX_INT(ka_max);
    // (a parcel that turned bad keeps ka = n_table in the mixed phase -- it stays "below" -- and runs moist rows in
    //  the above phase: both harmless)
    const int it_a = (KINDS & 2u) ? min(pr.K_ml, nt) : 1;
    const int it_b = (KINDS & 4u) ? max(it_a, min(pr.K_mu, nt)) : it_a;
    const bool fs = n_stash > 0;
    const int it_c = fs ? max(it_b, min(n_stash, nt)) : it_b;     // guards and the stash end here
    const int it_abv = min(max(it_c, ka_max + 2), nt + 1);        // mixed phase: [1, it_abv), above phase: [it_abv, nt]
    sweep_mixed7<KINDS & 5u, true>(rd, w, crow, stash, fs, 1, it_a, nt, qm, sb, ml, mu);
    sweep_mixed7<KINDS, true>(rd, w, crow, stash, fs, it_a, it_c, nt, qm, sb, ml, mu);
    sweep_mixed7<KINDS, false>(rd, w, crow, stash, false, it_c, min(it_abv, nt), nt, qm, sb, ml, mu);
    const int it_d = max(it_abv, min(pr.k_top + 1, nt));
    bool stopped = false;
    if (it_abv <= nt) {
        sweep_above7<KINDS, false>(rd, w, crow, it_abv, min(it_d, nt), nt, 0.0f, qm, sb, ml, mu);
        stopped = sweep_above7<KINDS, true>(rd, w, crow, max(it_d, it_abv), nt, nt, tmin_top - kStopMargin, qm, sb, ml, mu);
    }
    // last iteration (it == nt): there is no level nt; every parcel not bound for the exact path is above its LCL
    if (!stopped) {
        if (it_abv <= nt) {
            const float b = s.b_prv, h = w.h_prv;
            if (KINDS & 1u) step7_above(sb, w.itf, cubic_at(crow.at(sb.m), sb.f) - b, h);
            if (KINDS & 2u) step7_above(ml, w.itf, cubic_at(crow.at(ml.m), ml.f) - b, h);
            if (KINDS & 4u) step7_above(mu, w.itf, cubic_at(crow.at(mu.m), mu.f) - b, h);
        } else {
            const float big = 1e30f;
            if (KINDS & 1u) step7_mixed<0>(sb, nt, w.itf, cubic_at(crow.at(sb.m), sb.f) - s.b_prv, -big, s.x_prv, s.x_prv);
            if (KINDS & 2u) step7_mixed<0>(ml, nt, w.itf, cubic_at(crow.at(ml.m), ml.f) - s.b_prv, -big, s.x_prv, s.x_prv);
            if (KINDS & 4u) step7_mixed<0>(mu, nt, w.itf, cubic_at(crow.at(mu.m), mu.f) - s.b_prv, -big, s.x_prv, s.x_prv);
        }
    }
    // ---- results ---------------------------------------------------------------------------------------
    // A NaN/Inf T or Td in the pre-pass levels poisons nanacc; one in the swept levels makes P (or, before the LFC,
    // N) of every parcel that sweeps it non-finite.
    bool nan_seen = !(nanacc == 0.0f) || singular;
    if (KINDS & 1u) nan_seen = nan_seen || !(sb.pos - sb.tot < 3e38f);
    if (KINDS & 2u) nan_seen = nan_seen || !(ml.pos - ml.tot < 3e38f);
    if (KINDS & 4u) nan_seen = nan_seen || !(mu.pos - mu.tot < 3e38f);
    auto wrap = [&](const FParcel &c, FResult &r, unsigned bit) {
        sweep_finish7(c, cf, pr, o, r);
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
