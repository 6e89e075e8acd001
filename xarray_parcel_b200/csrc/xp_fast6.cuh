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

#ifndef XP_PREPASS_UNROLL
#define XP_PREPASS_UNROLL 1
#endif

namespace xp {
namespace fast {

constexpr int kPrepassUnroll = XP_PREPASS_UNROLL;    // the pre-pass loop is rolled: its float64 body is long

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

// One row of one parcel at iteration `it` (row schedule: see FParcel in xp_fast.cuh), given
//   d_m   parcel - environment on the moist adiabat at level it-1 (PF:585-592)
//   d_d   parcel - environment on the dry adiabat at level it     (PF:742)
//   x_cur / x_prv   ln p of level it / it-1
// GUARD 0: the parcel is live at every row.  1: rows it < kfirst are neutral and leave xprev at the row's
// ln p (most-unstable parcel on a shared axis: its start row is the level below kfirst).  2: rows it < kfirst
// are neutral and leave xprev untouched (per-column pressure: xprev was set to the parcel's start row).
template <int GUARD>
XP_HD void step6_core(FParcel &c, int it, float d_m, float d_d, float x_cur, float x_prv) {
    const bool above = it > c.ka, is_lcl = it == c.ka;
    float d = above ? d_m : d_d;
    d = is_lcl ? c.b_lcl : d;
    float x = above ? x_prv : x_cur;
    x = is_lcl ? c.x_lcl : x;
    bool active = true;
    if (GUARD) { active = it >= c.kfirst; d = active ? d : 0.0f; }
    if (GUARD == 2) x = active ? x : c.xprev;
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

// Shared pressure axis: the moist adiabat comes from the shared-memory cubic of level it-1 (`cc`);
// *_cur = level it, *_prv = level it-1: pk = p^kappa, x = ln p, b = environment virtual temperature.
// GUARD: the parcel may start above this row (most-unstable parcel).
template <bool GUARD>
XP_HD void step6(FParcel &c, int it, const Coef &cc, float pk_cur, float x_cur, float x_prv, float b_cur,
                 float b_prv) {
    step6_core<GUARD ? 1 : 0>(c, it, cubic_at(cc, c.f) - b_prv, f_fma(c.c_dryv, pk_cur, -b_cur), x_cur, x_prv);
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

// ---- parcel set-up in three stages, so that the three parcels of a column overlap their latencies -------------------
// (setup_parcel of xp_fast.cuh does the same work for one parcel start to end.)
//   A  LCL solve (arithmetic only)
//   B  issue the gathers: table cell of the LCL, T/Td of the two levels bracketing the LCL
//   C  everything that consumes them
struct Setup6 {
    double lp, lt;              // LCL (float64-polished)
    double p0, t0, td0;         // parcel (replaced by a dummy when the parcel is bound for the exact path)
    int adiabat;                // table cell -> adiabat number (gathered)
    float ta, tda, tb_, tdb;    // T/Td of the levels after / before the LCL
    bool before_is_start;
};

XP_HD void setup6_a(double p0, double t0, double td0, FParcel &pc, Setup6 &u) {
    pc.bad = false;
    // saturated / supersaturated / NaN parcels: exact path (see setup_parcel)
    if (!(t0 - td0 >= kSaturationMargin) || !(p0 > 0.0)) { pc.bad = true; t0 = 280.0; td0 = 270.0; p0 = 1000.0; }
    u.p0 = p0; u.t0 = t0; u.td0 = td0;
    lcl_fast6(p0, t0, td0, u.lp, u.lt);
}

// `lev(k, t, td)` reads T/Td of level k.
template <class Lev>
XP_HD void setup6_b(const Lev &lev, const Prep &pr, const Tables &tb, int knext, FParcel &pc, Setup6 &u) {
    float edge;
    u.adiabat = adiabat_cell(tb, u.lp, u.lt, edge);                   // PF:554-557 (2-byte gather)
    // LCL position among the levels of the lifted column (insert_level PF:965-966).  The axis is float32 data,
    // so comparing it with the float32-rounded LCL pressure decides every case except equality.
    const float lpf = (float)u.lp;
    int ka = knext;
#pragma unroll
    for (int step = 32; step > 0; step >>= 1) {
        const int k2 = ka + step;
        const bool ok = k2 <= pr.L && pr.p[min(k2, kMaxLevels) - 1] >= lpf;
        ka = ok ? k2 : ka;
    }
    pc.kfirst = knext;
    u.before_is_start = (ka == knext);
    const int kb = ka - 1;
    // LCL above the table top, or (to float32) exactly on a level: exact path
    if (ka >= pr.n_table || pr.p[kb] == lpf || pr.p[min(ka, pr.L - 1)] == lpf) { pc.bad = true; ka = pr.n_table; }
    pc.ka = ka;
    const int kl = min(ka, pr.L - 1);
    lev(kl, u.ta, u.tda);
    lev(max(kl - 1, 0), u.tb_, u.tdb);
}

XP_HD void setup6_c(const Prep &pr, const Opts &o, int kstart, FParcel &pc, const Setup6 &u) {
    const int a0 = u.adiabat - 1;
    pc.m = a0 / kNodeStride;
    if (u.adiabat <= 0 || pc.m < kFirstInterval || pc.m > kLastInterval) { pc.bad = true; pc.m = kFirstInterval; }
    pc.f = (float)(a0 - pc.m * kNodeStride) * (1.0f / kNodeStride);
    const float p0f = (float)u.p0, t0f = (float)u.t0, td0f = (float)u.td0;
    const float lpf = (float)u.lp, ltf = (float)u.lt;
    pc.lcl_p = lpf; pc.lcl_t = ltf;
    const float es_l = f_es(ltf);
    pc.lcl_tv = f_tv(ltf, f_mixing_ratio(es_l, es_l, lpf, 141));                   // PF:653-657
    const float w_parcel = f_mixing_ratio(f_es(t0f), f_es(td0f), p0f, 141);        // PF:748
    pc.c_dryv = t0f * f_rcp(pr.pk[kstart]) * f_fma(0.608f, w_parcel, 1.0f);        // PF:291-316, 767-775
    sweep_init6(pc, pr.lnp[kstart]);
    pc.x_lcl = pr.lnp[kstart]; pc.a_lcl = pc.b_lcl = 0.0f;
    if (pc.ka >= pr.n_table) return;                                               // bound for the exact path
    // environment at the LCL (PF:1774-1806): "before" = last level with p >= lcl_p (the start row when the LCL
    // is below the first swept level), "after" = level ka
    const int ka = pc.ka, kb = ka - 1;
    float tb_ = u.tb_, tdb = u.tdb, xb = o.log_interp ? pr.lnp[kb] : pr.p[kb];
    if (u.before_is_start) { tb_ = t0f; tdb = td0f; xb = o.log_interp ? pr.lnp[kstart] : pr.p[kstart]; }
    const float xa = o.log_interp ? pr.lnp[ka] : pr.p[ka];
    const float x_l = kLn2 * f_lg2(lpf);
    const float at = o.log_interp ? x_l : lpf;
    const float g = (at - xb) * f_rcp(xa - xb);
    const float te = f_fma(u.ta - tb_, g, tb_), tde = f_fma(u.tda - tdb, g, tdb);   // PF:1802
    const float etv = f_tv(te, f_mixing_ratio(f_es(te), f_es(tde), lpf, 141));      // PF:916-920
    pc.x_lcl = x_l;
    pc.a_lcl = pc.lcl_tv;
    pc.b_lcl = pc.lcl_tv - etv;                                                    // v6: the LCL row as a difference
    if (!(te == te) || !(tde == tde)) pc.bad = true;
}

// Per-thread stash of the T/Td of the lowest levels (filled by the pre-pass, read by the parcel set-up and
// by the first iterations of the sweep).  The kernel keeps it in shared memory.
struct NoStash {
    XP_HD int capacity() const { return 0; }
    XP_HD void put(int, float, float) const {}
    XP_HD void get(int, float &t, float &td) const { t = td = 0.0f; }
};

// Shared per-level state of the sweep: axis constants of level `it`, T/Td one level ahead in registers and
// kL2Ahead levels ahead as L2 prefetches (the register load then costs an L2 hit, which one iteration hides).
constexpr int kL2Ahead = 4;
// T/Td are addressed as base[off] with a 32-bit element offset off = column + level * level_stride from the
// (warp-uniform) array bases -- one integer add per level instead of 64-bit pointer arithmetic per array; the
// launcher takes this path only when every offset fits 32 bits.  Rd: ldT(off), ldTd(off), prefetch(off), ls(), off0().
struct Sweep6 {
    const float *lp;                      // packed axis constants {p, ln p, p^kappa, -} of level `it`
    uint32_t off;                         // offset of the level after the prefetched one
    uint32_t ls;
    int k_pf;                             // that level
    float t_n1, td_n1;                    // prefetched: the next global level
    float b_prv, x_prv;
};

// Iterations [it0, it1) of the sweep for the parcels in KACT (subset of KINDS).  FROM_STASH: T/Td of these
// levels come from the stash (it1 <= stash levels) instead of the global prefetch pipeline.
template <unsigned KACT, bool GUARD_MU, class Rd, class CoefRow, class Stash>
XP_HD void sweep_segment6(const Rd &rd, Sweep6 &s, CoefRow &crow, const Stash &stash, bool from_stash, int it0, int it1, int nt,
                          FParcel &sb, FParcel &ml, FParcel &mu) {
    for (int it = it0; it < it1; ++it) {
        float t, td;
        if (from_stash) {
            stash.get(it, t, td);
        } else {
            t = s.t_n1; td = s.td_n1;
            if (s.k_pf < nt) { s.t_n1 = rd.ldT(s.off); s.td_n1 = rd.ldTd(s.off); }    // one level ahead
            if (s.k_pf + kL2Ahead < nt) rd.prefetch(s.off + kL2Ahead * s.ls);
            s.off += s.ls; ++s.k_pf;
        }
        const float p_cur = s.lp[0], x_cur = s.lp[1], pk_cur = s.lp[2];
        s.lp += 4;
        const float b_cur = f_tv(t, f_mixing_ratio(f_es(t), f_es(td), p_cur, 141));   // PF:839-843
        if (KACT & 1u) step6<false>(sb, it, crow.at(sb.m), pk_cur, x_cur, s.x_prv, b_cur, s.b_prv);
        if (KACT & 2u) step6<false>(ml, it, crow.at(ml.m), pk_cur, x_cur, s.x_prv, b_cur, s.b_prv);
        if (KACT & 4u) step6<GUARD_MU>(mu, it, crow.at(mu.m), pk_cur, x_cur, s.x_prv, b_cur, s.b_prv);
        s.b_prv = b_cur; s.x_prv = x_cur;
        crow.advance();
    }
}

// The top of the sweep, iterations [it0, nt) with it0 > k_top: as sweep_segment6 (no guards, global pipeline), but
// after every iteration the warp checks whether it can stop.  A parcel is finished when it is above its LCL,
// colder than the environment, and its virtual temperature (which only falls with height along the moist
// adiabat) is more than kStopMargin below `tmin_top`, the coldest environment TEMPERATURE (<= virtual
// temperature) of all levels from k_top up: no crossing and no positive area can follow, and with
// pos_cape_neg_cin nothing the outputs depend on changes any more (PF:1329-1388; the negative area above the
// last crossing is never used).  Returns true if the sweep stopped early.
template <unsigned KINDS, class Rd, class CoefRow>
XP_HD bool sweep_top6(const Rd &rd, Sweep6 &s, CoefRow &crow, int it0, int nt, float tmin_top,
                      FParcel &sb, FParcel &ml, FParcel &mu) {
    const float stop_below = tmin_top - kStopMargin;
    for (int it = it0; it < nt; ++it) {
        const float t = s.t_n1, td = s.td_n1;
        if (s.k_pf < nt) { s.t_n1 = rd.ldT(s.off); s.td_n1 = rd.ldTd(s.off); }
        if (s.k_pf + kL2Ahead < nt) rd.prefetch(s.off + kL2Ahead * s.ls);
        s.off += s.ls; ++s.k_pf;
        const float p_cur = s.lp[0], x_cur = s.lp[1], pk_cur = s.lp[2];
        s.lp += 4;
        const float b_cur = f_tv(t, f_mixing_ratio(f_es(t), f_es(td), p_cur, 141));   // PF:839-843
        if (KINDS & 1u) step6<false>(sb, it, crow.at(sb.m), pk_cur, x_cur, s.x_prv, b_cur, s.b_prv);
        if (KINDS & 2u) step6<false>(ml, it, crow.at(ml.m), pk_cur, x_cur, s.x_prv, b_cur, s.b_prv);
        if (KINDS & 4u) step6<false>(mu, it, crow.at(mu.m), pk_cur, x_cur, s.x_prv, b_cur, s.b_prv);
        // the row just processed by a parcel above its LCL is level it-1: its curve there is dprev + b_prv
        bool done = true;
        if (KINDS & 1u) done = done && (sb.bad || (it > sb.ka && sb.dprev < 0.0f && sb.dprev + s.b_prv < stop_below));
        if (KINDS & 2u) done = done && (ml.bad || (it > ml.ka && ml.dprev < 0.0f && ml.dprev + s.b_prv < stop_below));
        if (KINDS & 4u) done = done && (mu.bad || (it > mu.ka && mu.dprev < 0.0f && mu.dprev + s.b_prv < stop_below));
        s.b_prv = b_cur; s.x_prv = x_cur;
        crow.advance();
        if (XP_WARP_ALL(done)) return true;
    }
    return false;
}

// The whole suite for one column, default options.  Interfaces as suite_column (xp_fast.cuh); `cf` must be
// the VIRTUAL-temperature table (compute_coef_tv).  Returns the mask of kinds for the exact path.
template <unsigned KINDS, class Rd, class Cf, class Stash>
XP_HD unsigned suite_column6(const Rd &rd, const Cf &cf, const Prep &pr, const Tables &tb, const Opts &o,
                             Stash &stash, FResult res[3]) {
    unsigned redo = 0;
    float nanacc = 0.0f;                   // becomes NaN if a T/Td read of the pre-pass is NaN or infinite
    const int nt = pr.n_table;
    const int n_low = max(1, max((KINDS & 4u) ? pr.K_mu : 0, (KINDS & 2u) ? pr.n_ml_w : 0));
    const uint32_t ls = rd.ls();
    // the stash is used only if it holds every pre-pass level (then the sweep starts from it too)
    const int n_stash = (stash.capacity() >= n_low) ? n_low : 0;
    // start the global prefetch pipeline of the sweep now: its first levels arrive during the pre-pass
    Sweep6 s;
    {
        const int k0 = (n_stash > 0) ? n_stash : 1;          // first level the sweep reads from global memory
        s.off = rd.off0() + (uint32_t)k0 * ls; s.ls = ls;
        s.t_n1 = s.td_n1 = 0.0f;
        if (k0 < nt) { s.t_n1 = rd.ldT(s.off); s.td_n1 = rd.ldTd(s.off); }
#pragma unroll
        for (int j = 1; j <= kL2Ahead; ++j)
            if (k0 + j < nt) rd.prefetch(s.off + (uint32_t)j * ls);
        s.off += ls; s.k_pf = k0 + 1;
    }
    // the levels above kTopCheckHpa are read right before the sweep (early-termination bound): ask L2 for them now
    for (int k = pr.k_top; k < nt; ++k) rd.prefetch(rd.off0() + (uint32_t)k * ls);
    // ---- pre-pass over the lowest levels: mixed-layer means (float64) and most-unstable argmax ----
    // (a rolled loop -- the float64 code below is long and every copy of it costs instruction-cache misses;
    //  the levels are asked for up front as L2 prefetches and then loaded one level ahead)
    double sum_th = 0.0, sum_w = 0.0;
    float best = -1e30f, second = -1e30f, mu_t = 0.0f, mu_td = 0.0f;
    int k_mu = 0;
    uint32_t off0 = rd.off0();
    for (int k = 1; k < n_low; ++k) rd.prefetch(off0 + (uint32_t)k * ls);
    const float t_sfc = rd.ldT(off0), td_sfc = rd.ldTd(off0);
    float t_nx = t_sfc, td_nx = td_sfc;
#pragma unroll(kPrepassUnroll)
    for (int k = 0; k < n_low; ++k) {
        const float t = t_nx, td = td_nx;
        off0 += ls;
        if (k + 1 < n_low) { t_nx = rd.ldT(off0); td_nx = rd.ldTd(off0); }
        if (k < n_stash) stash.put(k, t, td);
        nanacc = f_fma(t, 0.0f, f_fma(td, 0.0f, nanacc));
        const float p = pr.p[k];
        const float e = f_es(td);
        const float ipe = f_rcp(p - e);
        const float r = kEpsF * e * ipe;                 // saturation mixing ratio of the dewpoint (PF:258)
        if ((KINDS & 2u) && k < pr.n_ml_w) {
            // mixed_parcel PF:253-258 in float64 (see suite_column)
            // (reciprocals by float32-seeded Newton steps, ~1e-16 relative: no IEEE division)
            const double tdd = (double)td;
            const double e64 = kSat0 * exp64_fast(17.67 * (tdd - 273.15) * rcp64(tdd - 29.65));
            sum_th += pr.mlw[k] * ((double)t * pr.thfac[k]);
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
            if (v > best) { second = best; best = v; k_mu = k; mu_t = t; mu_td = td; }   // ties: larger p (PF:128)
            else if (v > second) second = v;
        }
    }
    // the top of the column (levels above kTopCheckHpa): the coldest environment temperature up there bounds what
    // a parcel can still meet (sweep_top6).  Read here, with the parcel set-up behind them to hide the latency
    // (L2 keeps the lines for the sweep); their NaNs must be seen even if the sweep stops below them.
    float tmin_top = 1e30f, tmax_top = -1e30f;
    for (int k = pr.k_top; k < nt; ++k) {
        const uint32_t o_ = rd.off0() + (uint32_t)k * ls;
        const float t = rd.ldT(o_), td = rd.ldTd(o_);
        nanacc = f_fma(t, 0.0f, f_fma(td, 0.0f, nanacc));
        tmin_top = fminf(tmin_top, t);
        tmax_top = fmaxf(tmax_top, t);
    }
    // the bound rests on virtual temperature >= temperature, i.e. on a non-negative mixing ratio
    // eps es(Td)/(p - es(T)) (PF:684-710): that needs es(T) < p at every one of these levels.  Where the
    // stratopause is warm enough for es(T) to reach the lowest pressure swept, the sweep runs to the top.
    // There the float32 mixing ratio also loses its accuracy as es(T) comes close to p (the reference's formula is
    // singular at es(T) = p): columns with a level within 5 % of that go to the exact path.
    bool singular = false;
    if (nt > pr.k_top && !(f_es(tmax_top) < 0.5f * pr.p[nt - 1])) {
        tmin_top = -1e30f;
        for (int k = pr.k_top; k < nt; ++k) {
            const float pk = pr.p[k];
            if (fabsf(pk - f_es(rd.ldT(rd.off0() + (uint32_t)k * ls))) < 0.05f * pk) singular = true;
        }
    }
    // ---- parcels: staged so that the LCL solves, the gathers and their consumers overlap --------------
    FParcel sb, ml, mu;
    Setup6 u_sb, u_ml, u_mu;
    auto lev = [&](int k, float &t, float &td) {
        if (k < n_stash) stash.get(k, t, td);
        else { const uint32_t o_ = rd.off0() + (uint32_t)k * ls; t = rd.ldT(o_); td = rd.ldTd(o_); }
    };
    double mp_t = 0.0, mp_td = 0.0;
    if (KINDS & 1u) setup6_a(pr.p0, (double)t_sfc, (double)td_sfc, sb, u_sb);
    if (KINDS & 2u) {
        mp_t = sum_th * pr.exner0;                                               // PF:268-269
        {   // dewpoint_from_e(vapor_pressure(p0, w)), PF:275-282, with the branch-free log and reciprocals
            const double val = log64_fast(pr.p0 * sum_w * rcp64(kEps + sum_w) * (1.0 / kSat0));
            mp_td = 243.5 * val * rcp64(17.67 - val) + kZeroC;
        }
        setup6_a(pr.p0, mp_t, mp_td, ml, u_ml);
    }
    if (KINDS & 4u) {
        if (!(best - second >= kThetaEMargin)) redo |= 4u;                       // argmax within float32 error
        setup6_a(pr.p64[k_mu], (double)mu_t, (double)mu_td, mu, u_mu);
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
    // ---- the sweep --------------------------------------------------------------------------------------
    s.lp = &pr.plk[1][0];
    s.b_prv = 0.0f; s.x_prv = pr.lnp[0];
    auto crow = cf.row(0);
    // segment bounds: the mixed-layer parcel joins at K_ml; most-unstable parcels have all started by K_mu;
    // levels below n_stash come from the stash
    const int it_a = (KINDS & 2u) ? min(pr.K_ml, nt) : 1;
    const int it_b = (KINDS & 4u) ? max(it_a, min(pr.K_mu, nt)) : it_a;
    const bool fs = n_stash > 0;
    // with a stash the first n_stash levels come from it (n_stash = n_low >= it_b), the rest from the
    // global prefetch pipeline, which was started at level n_stash
    const int it_c = fs ? max(it_b, min(n_stash, nt)) : it_b;
    sweep_segment6<KINDS & 5u, true>(rd, s, crow, stash, fs, 1, it_a, nt, sb, ml, mu);
    sweep_segment6<KINDS, true>(rd, s, crow, stash, fs, it_a, it_c, nt, sb, ml, mu);
    const int it_d = max(it_c, min(pr.k_top + 1, nt));
    sweep_segment6<KINDS, false>(rd, s, crow, stash, false, it_c, it_d, nt, sb, ml, mu);
    const bool stopped = sweep_top6<KINDS>(rd, s, crow, it_d, nt, tmin_top, sb, ml, mu);
    // last iteration: no level `nt`; every parcel that is not bound for the exact path is above its LCL
    if (!stopped) {
        const float big = 1e30f;
        if (KINDS & 1u) step6<false>(sb, nt, crow.at(sb.m), 0.0f, s.x_prv, s.x_prv, big, s.b_prv);
        if (KINDS & 2u) step6<false>(ml, nt, crow.at(ml.m), 0.0f, s.x_prv, s.x_prv, big, s.b_prv);
        if (KINDS & 4u) step6<false>(mu, nt, crow.at(mu.m), 0.0f, s.x_prv, s.x_prv, big, s.b_prv);
    }
    // ---- results ---------------------------------------------------------------------------------------
    // A NaN/Inf T or Td in the pre-pass levels poisons nanacc; one in the swept levels makes the area
    // sums of every parcel that sweeps it non-finite (each parcel sweeps every level above its start).
    bool nan_seen = !(nanacc == 0.0f) || singular;
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
