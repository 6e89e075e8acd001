// xp_fast.cuh -- the float32 fast path of the SB+ML+MU suite for columns on a SHARED pressure axis
// (ERA5-style pressure levels: BASELINE.json configs[3], the configuration the metric is quoted on).
//
// Design (DESIGN.md "fast path"):
//  * one thread per column; T/Td of a level are read once (coalesced) and the environment
//    virtual temperature is shared by the three parcels;
//  * the reference's moist-adiabat lookup (PF:525-607: nearest 0.5 hPa x 0.02 K cell -> adiabat
//    number -> np.interp on that adiabat) keeps its EXACT cell selection -- the LCL is polished
//    in float64 and the uint16 index grid is gathered from global memory, one 2-byte load per
//    parcel -- while the adiabat temperature at the (shared) level pressures comes from a table
//    in SHARED memory: per level, 4-point Lagrange cubics across every 64th adiabat, built per
//    call from the float32 curve table (max deviation from curve i: 2.5e-4 K, rms 9e-6 K);
//  * every DECISION of the reference (sign of parcel-minus-environment at a level, LCL
//    bracketing, most-unstable argmax, table cell) is either taken with a margin larger than
//    the float32 error bound or the column is handed to the float64 exact kernel
//    (xp_column.cuh) through a compact list -- so level indices and crossing brackets match the
//    reference and only VALUES carry float32 rounding (CAPE/CIN errors ~1e-2 J/kg).
//
// The same source compiles for the host (tests/hostsim) to check the logic against the oracle.
#pragma once
#include "xp_parcels.cuh"

namespace xp {
namespace fast {

constexpr int kNodeStride = 64;                                         // adiabats per cubic interval
constexpr int kNI = (kNAdiabats + kNodeStride - 1) / kNodeStride;       // 224 intervals
constexpr int kFirstInterval = 1, kLastInterval = (kNAdiabats - 1) / kNodeStride - 2;   // full 4-node stencils
constexpr int kMaxLevels = 56;                                          // 56*224*16 B = 196 KB of smem
constexpr float kDecisionEps = 6e-4f;       // K: |parcel - environment| below this is "uncertain"
constexpr float kThetaEMargin = 4e-6f;      // ln(theta_e) gap below this is a most-unstable tie
constexpr float kCrossSlope = 0.5f;         // K per unit ln p: a crossing with |d0 - d1| < kCrossSlope * dx is too
                                            // shallow to place within 1e-3 relative in pressure in float32
constexpr double kSaturationMargin = 2e-3;  // K: T - Td below this -> exact path (LCL snap, PF:644 isclose)

struct Coef { float c0, c1, c2, c3; };      // T(f) = c0 + f (c1 + f (c2 + f c3)), f in [0, 1)

// Per-call constants of the shared pressure axis, computed by the prep kernel (one thread).
struct Prep {
    int ok;                 // 1: the axis qualifies (finite, strictly decreasing, <= 1100 hPa, L <= kMaxLevels, ...)
    int L;                  // levels
    int n_table;            // levels 0..n_table-1 have 2.5 <= p (inside the adiabat table); the rest give NaN parcels
    int K_ml;               // number of levels in the mixed layer (p >= bottom - depth); first kept level of the ML column
    int n_ml_w;             // number of levels with a non-zero mixed-layer weight (K_ml or K_ml + 1)
    int K_mu;               // number of levels in the most-unstable search layer
    int pad0, pad1;
    double exner0;          // (p[0]/1000)^kappa
    double p0;              // p[0]
    double mlw[kMaxLevels];       // mixed-layer trapezoid weights / depth (PF:137-162 with get_layer PF:63-100)
    double thfac[kMaxLevels];     // (1000/p)^kappa  (potential temperature factor, PF:253)
    double p64[kMaxLevels];
    float p[kMaxLevels], lnp[kMaxLevels], pk[kMaxLevels];   // p, ln p, p^kappa
};

// ---- per-call constants of the shared pressure axis ----------------------------------------------------
XP_HD void compute_prep(const float *p, int64_t pls, int L, const Opts &o, Prep &pr) {
    pr.L = L;
    bool ok = (L >= 3 && L <= kMaxLevels);
    if (ok) {
        for (int k = 0; k < L; ++k) {
            const double pk = (double)p[(int64_t)k * pls];
            pr.p64[k] = pk;
            if (!(pk > 0.0) || !isfinite(pk) || (k > 0 && !(pk < pr.p64[k - 1]))) ok = false;
        }
    }
    if (ok && !(pr.p64[0] <= 1100.0)) ok = false;
    if (!ok) { pr.ok = 0; return; }
    int n_table = 0;
    for (int k = 0; k < L; ++k) {
        const double pk = pr.p64[k];
        if (pk >= 2.5) n_table = k + 1;
        pr.p[k] = (float)pk;
        pr.lnp[k] = (float)log(pk);
        pr.pk[k] = (float)pow(pk, kKappa);
        pr.thfac[k] = 1.0 / exner(pk);                                   // PF:253 theta = T / exner(p)
        pr.mlw[k] = 0.0;
    }
    pr.n_table = n_table;
    pr.p0 = pr.p64[0];
    pr.exner0 = exner(pr.p64[0]);
    // mixed layer: get_layer(interpolate=True) PF:63-100 + trapz(x='pressure') PF:186-198, as weights
    const double bottom = pr.p64[0];
    const double top = bottom - o.ml_depth;                              // PF:84
    int K_ml = 0;
    while (K_ml < L && pr.p64[K_ml] >= top) ++K_ml;
    pr.K_ml = K_ml;
    pr.n_ml_w = K_ml;
    if (K_ml < 1 || K_ml >= L) ok = false;
    if (ok) {
        for (int k = 1; k < K_ml; ++k) {
            const double dx = fabs(pr.p64[k] - pr.p64[k - 1]);
            pr.mlw[k - 1] += dx / 2; pr.mlw[k] += dx / 2;
        }
        const double pp = pr.p64[K_ml - 1];
        if (pp != top) {                                                 // interpolate the layer top in ln p (PF:85-90)
            const double pa = pr.p64[K_ml];
            const double g = (log(top) - log(pp)) / (log(pa) - log(pp));
            const double dx = fabs(top - pp);
            pr.mlw[K_ml - 1] += dx / 2 * (2.0 - g);
            pr.mlw[K_ml] += dx / 2 * g;
            pr.n_ml_w = K_ml + 1;
        }
        const double depth = fabs(top - bottom);                         // PF:158-159
        for (int k = 0; k < pr.n_ml_w; ++k) pr.mlw[k] *= (1. / depth);
    }
    // most-unstable layer: get_layer(interpolate=False) with bound_pressure PF:208-227
    const double bound = bottom - o.mu_depth;
    double best = fabs(pr.p64[0] - bound);
    int kt = 0;
    for (int k = 1; k < L; ++k) {
        const double d = fabs(pr.p64[k] - bound);
        if (d < best) { best = d; kt = k; }                              // ties keep the larger pressure
    }
    pr.K_mu = kt + 1;
    if (pr.K_mu > n_table || K_ml >= n_table || n_table < 3) ok = false;
    pr.ok = ok ? 1 : 0;
}

// coef[k][m]: 4-point Lagrange cubic through adiabats (m-1, m, m+1, m+2) * 64 (0-based) evaluated at level k
// exactly as the reference evaluates a single adiabat (np.interp on the 0.5 hPa nodes, PF:585-592).
XP_HD Coef compute_coef(const Prep &pr, const float *curves, int k, int m) {
    double y[4];
    for (int j = 0; j < 4; ++j) {
        int a = (m - 1 + j) * kNodeStride;
        a = min(max(a, 0), kNAdiabats - 1);
        y[j] = adiabat_temperature(curves + (size_t)a * kNP, pr.p64[k]);
    }
    Coef c;
    c.c0 = (float)y[1];
    c.c1 = (float)(-y[0] / 3 - y[1] / 2 + y[2] - y[3] / 6);
    c.c2 = (float)(y[0] / 2 - y[1] + y[2] / 2);
    c.c3 = (float)(-y[0] / 6 + y[1] / 2 - y[2] / 2 + y[3] / 6);
    return c;
}

// ---- float32 primitives (MUFU on the device) -------------------------------------------------
#if defined(__CUDACC__)
XP_HD float f_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
XP_HD float f_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
XP_HD float f_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
XP_HD float f_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
XP_HD float f_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#else
XP_HD float f_ex2(float x) { return exp2f(x); }
XP_HD float f_lg2(float x) { return log2f(x); }
XP_HD float f_rcp(float x) { return 1.0f / x; }
XP_HD float f_sqrt(float x) { return sqrtf(x); }
XP_HD float f_fma(float a, float b, float c) { return fmaf(a, b, c); }
#endif

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kEpsF = (float)kEps;
constexpr float kEsC = 17.67f * kLog2e;     // es = 6.112 * 2^(kEsC * (T - 273.15)/(T - 29.65))

// Bolton saturation vapour pressure, float32.  (T - 273.15)/(T - 29.65) = 1 - 243.5/(T - 29.65).
XP_HD float f_es(float t) {
    const float u = f_fma(-243.5f, f_rcp(t - 29.65f), 1.0f);
    return 6.112f * f_ex2(kEsC * u);
}
// PF:684-710 mixing_ratio(T, Td, p) from the two saturation vapour pressures.
//   MetPy 1.4.1: rh * ws          = eps * es(Td) / (p - es(T))
//   MetPy 1.6.2: eps ws rh / (eps + ws (1 - rh)) = eps * es(Td) / (p - es(Td))      (algebraically)
XP_HD float f_mixing_ratio(float es_t, float es_td, float p, int compat) {
    return kEpsF * es_td * f_rcp(p - (compat == 162 ? es_td : es_t));
}
XP_HD float f_tv(float t, float w) { return t * f_fma(0.608f, w, 1.0f); }

// ---- LCL (metpy.calc.lcl fixed point, PF:644) ----------------------------------------------------
// With w = eps es(Td)/(p0 - es(Td)) the vapour pressure of the lifted parcel is e(p) = es(Td) p/p0, so
//   v(q) = ln(e/6.112) = v0 + ln q,   v0 = 17.67 (Td - 273.15)/(Td - 29.65),  q = p/p0
//   tdp(v) = 243.5 v/(17.67 - v) + 273.15
// and the fixed point is F(q) = q - (tdp(v(q))/T)^3.5 = 0.  Newton in float32 from q = 1, then ONE
// Newton step in float64: the result agrees with the fully converged iteration to ~1e-12 relative,
// which is what makes the table-cell selection below identical to the reference's.
XP_HD void lcl_fast(double p0, double t, double td, double &lcl_p, double &lcl_t) {
    const double v0 = 17.67 * (td - 273.15) / (td - 29.65);
    // float32 Newton
    const float v0f = (float)v0, rt = f_rcp((float)t);
    float q = 1.0f;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const float v = f_fma(kLn2, f_lg2(q), v0f);
        const float iv = f_rcp(17.67f - v);
        const float tdp = f_fma(243.5f * v, iv, 273.15f);
        const float r = tdp * rt;
        const float r35 = r * r * r * f_sqrt(r);
        const float dtdp = 243.5f * 17.67f * iv * iv;                   // d tdp / d v
        const float dF = 1.0f - 3.5f * r35 * dtdp * f_rcp(tdp * q);     // d/dq: r35 * 3.5 * (dtdp/tdp) * (1/q)
        q = q - (q - r35) * f_rcp(dF);
    }
    // float64 polish
    double qd = (double)q;
    {
        const double v = v0 + log(qd);
        const double iv = 1.0 / (17.67 - v);
        const double tdp = 243.5 * v * iv + 273.15;
        const double r = tdp / t;
        const double r35 = r * r * r * sqrt(r);
        const double dtdp = 243.5 * 17.67 * iv * iv;
        const double dF = 1.0 - 3.5 * r35 * dtdp / (tdp * qd);
        const double dq = (qd - r35) / dF;
        qd = qd - dq;
        // tdp at the polished q, first order (|dq| ~ 1e-6: the second-order term is < 1e-11 K)
        lcl_t = tdp - dtdp * dq / qd;
    }
    lcl_p = p0 * qd;
}

// ---- sweep state of one parcel (float32 version of Sweep in xp_column.cuh) -------------------------
struct FSweep {
    float xprev, dprev, aprev;          // ln p, parcel - environment, parcel curve at the previous row
    float pos, tot;                     // running sums of positive areas and of all areas (ln p units)
    float lcl_pos, lcl_tot;
    float lfc_pos, lfc_tot, lfc_x, lfc_t;
    float el_pos, el_tot, el_x, el_t;
    int row;                            // index of the next row
    int lfc_row, el_row;                // upper row of the interval holding the LFC / EL crossing (-1: none)
    bool after_lcl, any_inc, pos_parcel, el_above, unc;

    XP_HD void init(float x0, float a0) {
        xprev = x0; dprev = 0.0f; aprev = a0;      // first row: parcel == environment (PF:1117-1120)
        pos = tot = lcl_pos = lcl_tot = 0.0f;
        lfc_pos = lfc_tot = lfc_x = lfc_t = 0.0f;
        el_pos = el_tot = el_x = el_t = 0.0f;
        row = 1; lfc_row = el_row = -1;
        after_lcl = any_inc = pos_parcel = el_above = unc = false;
    }

    // Next row at ln p = x with parcel curve a and environment curve b (find_intersections
    // PF:992-1064 + trap_around_zeros PF:1200-1289 + trapz PF:164-206 for the interval below it).
    XP_HD void step(float x, float a, float b, bool is_lcl) {
        const float d = a - b;
        const float dx = xprev - x;
        const bool cross = dprev * d < 0.0f;
        const float frac = cross ? dprev * f_rcp(dprev - d) : 1.0f;    // zero at xprev - frac dx
        const float h = 0.5f * dx;
        const float a_lo = h * dprev * frac;
        const float a_hi = h * d * (cross ? 1.0f - frac : 1.0f);
        pos += fmaxf(a_lo, 0.0f); tot += a_lo;
        if (cross) {
            unc = unc || (fabsf(dprev - d) < kCrossSlope * dx);
            const float ix = f_fma(-frac, dx, xprev);
            const float iy = f_fma(frac, a - aprev, aprev);
            if (d > 0.0f) {                                            // increasing (PF:1058)
                any_inc = true;
                if (after_lcl && lfc_row < 0) {                        // max-pressure one above the LCL (PF:1127-1132)
                    lfc_row = row; lfc_pos = pos; lfc_tot = tot; lfc_x = ix; lfc_t = iy;
                }
            } else {                                                   // decreasing: min pressure wins (PF:1136)
                el_row = row; el_pos = pos; el_tot = tot; el_x = ix; el_t = iy; el_above = after_lcl;
            }
        }
        pos += fmaxf(a_hi, 0.0f); tot += a_hi;
        if (is_lcl) { lcl_pos = pos; lcl_tot = tot; after_lcl = true; }
        else if (after_lcl && d > 0.0f) pos_parcel = true;             // PF:1166-1169
        unc = unc || !(fabsf(d) >= kDecisionEps);                      // also catches NaN
        xprev = x; dprev = d; aprev = a; ++row;
    }
};

struct FResult {
    float cape, cin, lcl_p, lcl_t, lcl_tv, lfc_p, lfc_t, el_p, el_t, par_p, par_t, par_td;
    int shift;
};

XP_HD float f_qnan() {
#if defined(__CUDACC__)
    return __int_as_float(0x7fffffff);
#else
    return std::nanf("");
#endif
}

// Sweep.finish of xp_column.cuh (lfc_el PF:1140-1185, cape_cin_base PF:1329-1388) on the float32 state.
XP_HD void finish(const FSweep &s, float lcl_p, float lcl_targ, const Opts &o, FResult &r) {
    const bool top_colder = s.dprev <= 0.0f;                            // PF:1151
    const bool el_exists = top_colder && s.el_row >= 0 && s.el_above;   // PF:1152-1153
    const bool lfc_missing = !s.any_inc;                                // PF:1161
    const bool lfc_found = s.lfc_row >= 0;
    const bool replace = (s.pos_parcel && lfc_missing) || (!lfc_missing && !lfc_found && el_exists);
    const bool have_lfc = lfc_found || replace;
    float l_pos = s.lfc_pos, l_tot = s.lfc_tot;
    r.lfc_p = lfc_found ? f_ex2(s.lfc_x * kLog2e) : f_qnan();
    r.lfc_t = lfc_found ? s.lfc_t : f_qnan();
    if (replace) { r.lfc_p = lcl_p; r.lfc_t = lcl_targ; l_pos = s.lcl_pos; l_tot = s.lcl_tot; }
    r.el_p = el_exists ? f_ex2(s.el_x * kLog2e) : f_qnan();
    r.el_t = el_exists ? s.el_t : f_qnan();
    float cape = 0.0f, cin = 0.0f;
    if (have_lfc) {
        const float e_pos = el_exists ? s.el_pos : s.pos;
        const float e_tot = el_exists ? s.el_tot : s.tot;
        // EL below the LFC (PF:1352-1353 leaves no level between them): compare by interval index
        const bool el_below_lfc = el_exists && lfc_found && !replace && s.el_row < s.lfc_row;
        if (o.pos_neg) {
            cin = l_tot - l_pos;
            cape = el_below_lfc ? 0.0f : (e_pos - l_pos);
        } else {
            cin = l_tot;
            cape = el_below_lfc ? 0.0f : (e_tot - l_tot);
        }
    }
    cape *= (float)kRd; cin *= (float)kRd;
    if (o.post_zero && !(cin <= 0.0f)) cin = 0.0f;
    r.cape = cape; r.cin = cin;
}

// One parcel: everything that is constant along the column.
struct FParcel {
    float c_dry, w_parcel;      // dry adiabat T = c_dry * p^kappa (PF:291-316); parcel mixing ratio (PF:748)
    float f;                    // position inside the cubic interval
    int m;                      // cubic interval of the adiabat
    int ka;                     // first level above the LCL (p < lcl_p)
    int kfirst;                 // first level of the lifted column that is swept (after the start row)
    float x_lcl, a_lcl, b_lcl;  // the inserted LCL row (PF:858-931)
    float lcl_p, lcl_t, lcl_tv;
    bool bad;                   // this parcel must go to the exact path
};

// Set up one parcel after its (p0, T0, Td0) are known.  `kstart` is the level of the start row (the
// parcel level; for the mixed layer the start row is the prepended parcel itself and the column
// continues at `knext`).  Td/T of the bracketing levels come through `rd`.
template <class Rd>
XP_HD void setup_parcel(const Rd &rd, const Prep &pr, const Tables &tb, const Opts &o, double p0, double t0,
                        double td0, int kstart, int knext, bool start_is_virtual, FParcel &pc) {
    pc.bad = false;
    pc.kfirst = knext;
    // saturated / supersaturated / NaN parcels: the LCL snaps to the parcel level (np.isclose in
    // metpy.calc.lcl) or lies below it -- zero-width intervals, exact equality tests: exact path.
    if (!(t0 - td0 >= kSaturationMargin) || !(p0 > 0.0)) { pc.bad = true; t0 = 280.0; td0 = 270.0; p0 = 1000.0; }
    double lp, lt;
    lcl_fast(p0, t0, td0, lp, lt);
    const int adiabat = adiabat_lookup(tb, lp, lt);                    // PF:554-557, exact cell
    const int a0 = adiabat - 1;
    pc.m = a0 / kNodeStride;
    if (adiabat <= 0 || pc.m < kFirstInterval || pc.m > kLastInterval) { pc.bad = true; pc.m = kFirstInterval; }
    pc.f = (float)(a0 - pc.m * kNodeStride) * (1.0f / kNodeStride);
    const float p0f = (float)p0, t0f = (float)t0, td0f = (float)td0;
    const float lpf = (float)lp, ltf = (float)lt;
    pc.lcl_p = lpf; pc.lcl_t = ltf;
    const float es_l = f_es(ltf);
    pc.lcl_tv = f_tv(ltf, f_mixing_ratio(es_l, es_l, lpf, o.compat));           // PF:653-657
    pc.w_parcel = f_mixing_ratio(f_es(t0f), f_es(td0f), p0f, o.compat);         // PF:748
    pc.c_dry = t0f * f_rcp(pr.pk[kstart]);                                      // p0 == p[kstart]
    // LCL position among the levels of the lifted column (insert_level PF:965-966): exact in float64
    int ka = knext;
    while (ka < pr.L && pr.p64[ka] >= lp) ++ka;
    pc.ka = ka;
    // bracketing levels for the environment at the LCL (PF:1774-1806): "before" = last level with
    // p >= lcl_p (the start row when the LCL is below the first swept level), "after" = level ka.
    const int kb = ka - 1;
    const bool before_is_start = (ka == knext);
    if (ka >= pr.n_table || (!before_is_start && pr.p64[kb] == lp) || (before_is_start && p0 == lp)) {
        pc.bad = true;      // LCL above the table top / exactly on a level: exact path
        pc.ka = pr.L; pc.x_lcl = pr.lnp[pr.L - 1]; pc.a_lcl = pc.b_lcl = 0.0f;
        return;
    }
    float tb_, tdb, xb;
    if (before_is_start) { tb_ = t0f; tdb = td0f; xb = o.log_interp ? pr.lnp[kstart] : pr.p[kstart]; }
    else { tb_ = rd.T(kb); tdb = rd.Td(kb); xb = o.log_interp ? pr.lnp[kb] : pr.p[kb]; }
    if (start_is_virtual && before_is_start) xb = o.log_interp ? pr.lnp[0] : pr.p[0];
    const float ta = rd.T(ka), tda = rd.Td(ka);
    const float xa = o.log_interp ? pr.lnp[ka] : pr.p[ka];
    const float x_l = kLn2 * f_lg2(lpf);
    const float at = o.log_interp ? x_l : lpf;
    const float g = (at - xb) * f_rcp(xa - xb);
    const float te = f_fma(ta - tb_, g, tb_), tde = f_fma(tda - tdb, g, tdb);  // PF:1802
    const float etv = f_tv(te, f_mixing_ratio(f_es(te), f_es(tde), lpf, o.compat));     // PF:916-920
    pc.x_lcl = x_l;
    pc.a_lcl = o.vtc ? pc.lcl_tv : ltf;
    pc.b_lcl = o.vtc ? etv : te;
    if (!(te == te) || !(tde == tde)) pc.bad = true;
}

// The whole suite for one column.  Rd: float T(k), Td(k).  Cf: Coef at(k, m) from shared memory.
// Returns the mask of parcel kinds (bit 0 SB, 1 ML, 2 MU) that must be recomputed by the exact path.
template <class Rd, class Cf>
XP_HD unsigned suite_column(const Rd &rd, const Cf &cf, const Prep &pr, const Tables &tb, const Opts &o,
                            unsigned kinds, FResult res[3]) {
    unsigned redo = 0;
    bool nan_seen = false;
    // ---- pre-pass over the lowest levels: mixed-layer means (float64) and most-unstable argmax ----
    double sum_th = 0.0, sum_w = 0.0;
    float best = -1e30f, second = -1e30f, mu_t = 0.0f, mu_td = 0.0f;
    int k_mu = 0;
    const int n_pre = max((kinds & 4u) ? pr.K_mu : 0, (kinds & 2u) ? pr.n_ml_w : 0);
    for (int k = 0; k < n_pre; ++k) {
        const float t = rd.T(k), td = rd.Td(k);
        nan_seen = nan_seen || !(t == t) || !(td == td);
        if ((kinds & 2u) && k < pr.n_ml_w) {
            // mixed_parcel PF:253-258: theta and saturation mixing ratio of the dewpoint, float64
            const double e = sat_vapor_pressure((double)td);
            sum_th += pr.mlw[k] * ((double)t * pr.thfac[k]);
            sum_w += pr.mlw[k] * (kEps * e / (pr.p64[k] - e));
        }
        if ((kinds & 4u) && k < pr.K_mu) {
            // ln(theta_e), Bolton (1980) eq. 39 as in metpy.calc.equivalent_potential_temperature (PF:123)
            const float p = pr.p[k];
            const float e = f_es(td);
            const float ipe = f_rcp(p - e);
            const float r = kEpsF * e * ipe;
            const float l2t = f_lg2(t), l2td = f_lg2(td);
            const float t_l = 56.0f + f_rcp(f_rcp(td - 56.0f) + (l2t - l2td) * (kLn2 / 800.0f));
            const float it_l = f_rcp(t_l);
            float v = l2t * kLn2;                                                   // ln T
            v = f_fma((float)kKappa * kLn2, f_lg2(1000.0f * ipe), v);               // + kappa ln(1000/(p-e))
            v = f_fma(0.28f * r * kLn2, l2t - f_lg2(t_l), v);                       // + 0.28 r ln(T/t_l)
            v = f_fma(r * f_fma(0.448f, r, 1.0f), f_fma(3036.0f, it_l, -1.78f), v);
            if (v > best) { second = best; best = v; k_mu = k; mu_t = t; mu_td = td; }   // ties: larger p (PF:128)
            else if (v > second) second = v;
            if (!(v == v)) nan_seen = true;
        }
    }
    // ---- parcels --------------------------------------------------------------------------------
    FParcel pc[3];
    FSweep sw[3];
    const float t_sfc = rd.T(0), td_sfc = rd.Td(0);
    nan_seen = nan_seen || !(t_sfc == t_sfc) || !(td_sfc == td_sfc);
    if (kinds & 1u) {
        setup_parcel(rd, pr, tb, o, pr.p0, (double)t_sfc, (double)td_sfc, 0, 1, false, pc[0]);
        res[0].par_p = pr.p[0]; res[0].par_t = t_sfc; res[0].par_td = td_sfc; res[0].shift = 0;
        sw[0].init(pr.lnp[0], o.vtc ? f_tv(t_sfc, pc[0].w_parcel) : t_sfc);
    }
    if (kinds & 2u) {
        const double mp_t = sum_th * pr.exner0;                                  // PF:268-269
        const double mp_td = dewpoint_from_e(vapor_pressure(pr.p0, sum_w));      // PF:275-282
        setup_parcel(rd, pr, tb, o, pr.p0, mp_t, mp_td, 0, pr.K_ml, true, pc[1]);
        res[1].par_p = pr.p[0]; res[1].par_t = (float)mp_t; res[1].par_td = (float)mp_td; res[1].shift = pr.K_ml;
        sw[1].init(pr.lnp[0], o.vtc ? f_tv((float)mp_t, pc[1].w_parcel) : (float)mp_t);
    }
    if (kinds & 4u) {
        if (!(best - second >= kThetaEMargin)) redo |= 4u;                       // argmax within float32 error
        setup_parcel(rd, pr, tb, o, pr.p64[k_mu], (double)mu_t, (double)mu_td, k_mu, k_mu + 1, false, pc[2]);
        res[2].par_p = pr.p[k_mu]; res[2].par_t = mu_t; res[2].par_td = mu_td; res[2].shift = k_mu;
        sw[2].init(pr.lnp[k_mu], o.vtc ? f_tv(mu_t, pc[2].w_parcel) : mu_t);
    }
    // ---- the sweep over the levels -------------------------------------------------------------------
    for (int k = 1; k < pr.n_table; ++k) {
        const float t = rd.T(k), td = rd.Td(k);
        nan_seen = nan_seen || !(t == t) || !(td == td);
        const float p = pr.p[k], x = pr.lnp[k];
        const float es_t = f_es(t), es_td = f_es(td);
        const float b = o.vtc ? f_tv(t, f_mixing_ratio(es_t, es_td, p, o.compat)) : t;      // PF:839-843
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            if (!((kinds >> q) & 1u) || k < pc[q].kfirst) continue;
            const FParcel &c = pc[q];
            if (k == c.ka) sw[q].step(c.x_lcl, c.a_lcl, c.b_lcl, true);          // the inserted LCL row
            float tp, w;
            if (k < c.ka) { tp = c.c_dry * pr.pk[k]; w = c.w_parcel; }           // PF:742, 767, 773
            else {
                const Coef cc = cf.at(k, c.m);
                tp = f_fma(f_fma(f_fma(cc.c3, c.f, cc.c2), c.f, cc.c1), c.f, cc.c0);        // PF:585-592
                const float es = f_es(tp);
                w = kEpsF * es * f_rcp(p - es);                                  // PF:760
            }
            const float a = o.vtc ? f_tv(tp, w) : tp;
            sw[q].step(x, a, b, false);
        }
    }
    // ---- results ---------------------------------------------------------------------------------------
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        if (!((kinds >> q) & 1u)) continue;
        const FParcel &c = pc[q];
        FResult &r = res[q];
        r.lcl_p = c.lcl_p; r.lcl_t = c.lcl_t; r.lcl_tv = c.lcl_tv;
        finish(sw[q], c.lcl_p, o.vtc ? c.lcl_tv : c.lcl_t, o, r);
        if (c.bad || sw[q].unc || nan_seen || c.ka >= pr.n_table) redo |= (1u << q);
    }
    return redo & kinds;
}

}  // namespace fast
}  // namespace xp
