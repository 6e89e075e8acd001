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
#ifndef XP_TOP_CHECK_HPA
#define XP_TOP_CHECK_HPA 125.0
#endif
constexpr double kTopCheckHpa = XP_TOP_CHECK_HPA;   // v6 sweep: early termination is considered above this pressure
constexpr float kStopMargin = 1.0f;          // K: early-termination margin below the coldest environment level
constexpr unsigned kRedoRowsOk = 16u;        // << kind (bits 4-6), profile calls only: the column is on the list for a crossing
                                             // decision alone, its float32 profile rows stand and the fix-up rewrites scalars only
constexpr unsigned kRedoMuIsSb = 8u;         // redo-mask bit (== kListMuIsSb): write the SB exact result to the MU outputs too
constexpr double kSaturationMargin = 2e-3;
// margins for a parcel that is itself only float32-accurate (`approx`; currently unused: all parcels are
// exact float32 data or float64 means): |dT_lcl| < 4e-5 K, |dp_lcl| < 5e-4 hPa
constexpr float kApproxCellEdge = 4e-3f;    // cell units: 8e-5 K of 0.02 K, 2e-3 hPa of 0.5 hPa
constexpr double kApproxLclP = 2e-3;        // hPa  // K: T - Td below this -> exact path (LCL snap, PF:644 isclose)

struct Coef { float c0, c1, c2, c3; };      // T(f) = c0 + f (c1 + f (c2 + f c3)), f in [0, 1)

// Per-call constants of the shared pressure axis, computed by the prep kernel (one thread).
struct Prep {
    int ok;                 // 1: the axis qualifies (finite, strictly decreasing, <= 1100 hPa, L <= kMaxLevels, ...)
    int L;                  // levels
    int n_table;            // levels 0..n_table-1 have 2.5 <= p (inside the adiabat table); the rest give NaN parcels
    int K_ml;               // number of levels in the mixed layer (p >= bottom - depth); first kept level of the ML column
    int n_ml_w;             // number of levels with a non-zero mixed-layer weight (K_ml or K_ml + 1)
    int K_mu;               // number of levels in the most-unstable search layer
    int k_top;              // first level above kTopCheckHpa (n_table when there are fewer than 3): from here on the
                            // v6 sweep may stop once no parcel of the warp can meet the environment again
    int pad1;
    double exner0;          // (p[0]/1000)^kappa
    double p0;              // p[0]
    double mlw[kMaxLevels];       // mixed-layer trapezoid weights / depth (PF:137-162 with get_layer PF:63-100)
    double thfac[kMaxLevels];     // (1000/p)^kappa  (potential temperature factor, PF:253)
    double p64[kMaxLevels];
    float p[kMaxLevels], lnp[kMaxLevels], pk[kMaxLevels];   // p, ln p, p^kappa
    float plk[kMaxLevels][4];                               // the same three, packed per level (one 16-byte load), +
                                                            // [3] = 0.5 (ln p[k-1] - ln p[k]): half-width of the interval below level k
};

// ---- per-call constants of the shared pressure axis ----------------------------------------------------
// Level part (independent per level; the prep kernel runs it with one thread per level).
template <typename TP>
XP_HD void compute_prep_level(const TP *p, int64_t pls, int k, Prep &pr) {
    const double pk = (double)p[(int64_t)k * pls];
    pr.p64[k] = pk;
    pr.p[k] = (float)pk;
    pr.lnp[k] = (float)log(pk);
    pr.pk[k] = (float)pow(pk, kKappa);
    pr.plk[k][0] = pr.p[k]; pr.plk[k][1] = pr.lnp[k]; pr.plk[k][2] = pr.pk[k]; pr.plk[k][3] = 0.0f;
    pr.thfac[k] = 1.0 / exner(pk);                                   // PF:253 theta = T / exner(p)
    pr.mlw[k] = 0.0;
}

// Axis part (after every level is done).
XP_HD void compute_prep_axis(int L, const Opts &o, Prep &pr) {
    pr.L = L;
    bool ok = (L >= 3 && L <= kMaxLevels);
    if (ok) {
        for (int k = 0; k < L; ++k) {
            const double pk = pr.p64[k];
            if (!(pk > 0.0) || !isfinite(pk) || (k > 0 && !(pk < pr.p64[k - 1]))) ok = false;
        }
    }
    if (ok && !(pr.p64[0] <= 1100.0)) ok = false;
    if (!ok) { pr.ok = 0; return; }
    int n_table = 0;
    for (int k = 0; k < L; ++k)
        if (pr.p64[k] >= 2.5) n_table = k + 1;
    pr.n_table = n_table;
    for (int k = 1; k < L; ++k) pr.plk[k][3] = 0.5f * (pr.lnp[k - 1] - pr.lnp[k]);   // float32, as the sweeps compute it
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
    {
        int k_top = n_table;
        for (int k = n_table - 1; k >= 0 && pr.p64[k] < kTopCheckHpa; --k) k_top = k;
        pr.k_top = (n_table - k_top >= 3) ? k_top : n_table;
    }
    if (pr.K_mu > n_table || K_ml >= n_table || n_table < 3) ok = false;
    pr.ok = ok ? 1 : 0;
}

template <typename TP>
XP_HD void compute_prep(const TP *p, int64_t pls, int L, const Opts &o, Prep &pr) {
    for (int k = 0; k < L && k < kMaxLevels; ++k) compute_prep_level(p, pls, k, pr);
    compute_prep_axis(L, o, pr);
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

// warp-uniform vote (the host simulation runs one column at a time)
#if defined(__CUDACC__)
#define XP_WARP_ALL(x) __all_sync(0xffffffffu, (x))
#else
#define XP_WARP_ALL(x) (x)
#endif

// Profile-output policies (parcel_profile_with_lcl rows, PF:806-931).  kEnabled = false compiles the row
// bookkeeping away; the kernels provide a writer with
//   put(kind, row, p, parcel T, parcel Tv, environment T, environment Tv, environment Td).
struct NoProfile {
    static constexpr bool kEnabled = false;
    XP_HD void put(int, int, float, float, float, float, float, float) const {}
};

// Environment-curve policies of suite_column (see there).
struct EnvRecompute {                 // environment recomputed in the sweep, no early termination
    static constexpr bool kStaged = false, kFullPass = false;
    XP_HD void put(int, float) {}
    XP_HD float get(int) const { return 0.0f; }
};
struct EnvMinOnly {                   // first pass over all levels for the coldest environment level only
    static constexpr bool kStaged = false, kFullPass = true;
    XP_HD void put(int, float) {}
    XP_HD float get(int) const { return 0.0f; }
};

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

// (rcp64 / sqrt64 / log64_fast / exp64_fast -- the branch-free float64 helpers -- live in xp_math.cuh: the exact
//  fix-up over the uncertain-column list uses them too)
using xp::exp64_fast;
using xp::pow_kappa64;
using xp::log64_fast;
using xp::rcp64;
using xp::sqrt64;

// ---- specific humidity -> dewpoint on load (xp_columns.dewpoint_is_specific_humidity; PF:1889, 1969) -------------------
// metpy.calc.dewpoint_from_specific_humidity in the form of MetPy `compat` (see dewpoint_from_q in xp_math.cuh).
// float32 for the levels of the sweep (|error| ~1e-5 K, far inside the decision margins: d Tv / d Td ~ 2e-4) ...
XP_HD float f_td_from_q(float p, float t, float q, int compat) {
    const float w = q * f_rcp(1.0f - q);
    const float e = (compat == 162) ? p * w * f_rcp(kEpsF + w) : w * (p - f_es(t)) * (1.0f / kEpsF);
    const float v = kLn2 * f_lg2(e * (1.0f / 6.112f));
    return f_fma(243.5f * v, f_rcp(17.67f - v), 273.15f);
}
// ... float64 (branch-free helpers, ~3 ulp) for what the table cell of a parcel's LCL hangs on: the vapour pressure of
// a level (mixed-layer mean) and the dewpoint of a parcel level.
XP_HD double e64_from_q_fast(double p, double t, double q, int compat) {
    const double w = q * rcp64(1.0 - q);
    if (compat == 162) return p * w * rcp64(kEps + w);
    const double es_t = kSat0 * exp64_fast(17.67 * (t - 273.15) * rcp64(t - 29.65));
    return w * (p - es_t) * (1.0 / kEps);
}
XP_HD double td64_from_q_fast(double p, double t, double q, int compat) {
    const double val = log64_fast(e64_from_q_fast(p, t, q, compat) * (1.0 / kSat0));
    return 243.5 * val * rcp64(17.67 - val) + kZeroC;
}

// ---- LCL (metpy.calc.lcl fixed point, PF:644), cheaper float64 polish -------------------------------------------
// Same scheme as lcl_fast (xp_fast.cuh): float32 Newton on F(q) = q - (tdp(v0 + ln q)/T)^3.5, then ONE float64
// Newton step.  Here only the RESIDUAL F is evaluated in float64 (one log, reciprocals and a square root by
// float32-seeded Newton iterations, no IEEE division); its derivative is the float32 one, whose 1e-6 relative
// error enters the step (|dq| ~ 1e-7 q) at second order.
XP_HD void lcl_fast6(double p0, double t, double td, double &lcl_p, double &lcl_t) {
    const double v0 = 17.67 * (td - 273.15) * rcp64(td - 29.65);
    const float v0f = (float)v0, rt = f_rcp((float)t);
    // start from Bolton's (1980) LCL temperature, good to ~0.1 K: q0 = (t_l / T)^3.5 is within ~1e-3 of the root,
    // two Newton steps reach float32 rounding
    float q, dF = 1.0f, dtdp = 0.0f;
    {
        const float tf = (float)t, tdf = (float)td;
        const float l2t = f_lg2(tf);
        const float t_l = 56.0f + f_rcp(f_rcp(tdf - 56.0f) + (l2t - f_lg2(tdf)) * (kLn2 / 800.0f));
        q = f_ex2(3.5f * (f_lg2(t_l) - l2t));
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const float v = f_fma(kLn2, f_lg2(q), v0f);
        const float iv = f_rcp(17.67f - v);
        const float tdp = f_fma(243.5f * v, iv, 273.15f);
        const float r = tdp * rt;
        const float r35 = r * r * r * f_sqrt(r);
        dtdp = 243.5f * 17.67f * iv * iv;                               // d tdp / d v
        dF = 1.0f - 3.5f * r35 * dtdp * f_rcp(tdp * q);                 // d/dq: r35 * 3.5 * (dtdp/tdp) * (1/q)
        q = q - (q - r35) * f_rcp(dF);
    }
    // (dF, dtdp belong to the previous iterate: they differ from the ones at q by ~1e-6 relative)
    const double qd = (double)q;
    const double v = v0 + log64_fast(qd);
    const double tdp = 243.5 * v * rcp64(17.67 - v) + 273.15;
    const double r = tdp * rcp64(t);
    const double r35 = r * r * r * sqrt64(r);
    const double dq = (qd - r35) * (double)f_rcp(dF);
    // tdp at the polished q, first order (|dq| ~ 1e-7: the second-order term is < 1e-11 K)
    lcl_t = tdp - (double)(dtdp * f_rcp(q)) * dq;
    lcl_p = p0 * (qd - dq);
}

// (the first version of the solver -- four Newton steps from q = 1, libm float64 polish -- gave the same LCL to
// 2e-11; every fast path now uses the cheaper one)
XP_HD void lcl_fast(double p0, double t, double td, double &lcl_p, double &lcl_t) { lcl_fast6(p0, t, td, lcl_p, lcl_t); }

// mixed_parcel's last steps (PF:268-282) with the branch-free float64 helpers: temperature = theta * exner(p0),
// dewpoint = dewpoint(vapor_pressure(p0, w)).
XP_HD void mixed_parcel_t_td(double p0, double theta, double w, double &t, double &td) {
    t = theta * exp64_fast(kKappa * log64_fast(p0 * 1e-3));                      // PF:268-269
    const double val = log64_fast(p0 * w * rcp64(kEps + w) * (1.0 / kSat0));     // PF:275-282
    td = 243.5 * val * rcp64(17.67 - val) + kZeroC;
}

// ---- table cell of the LCL (PF:554-557 .sel(method='nearest')) --------------------------------------------
// Nearest 0.5 hPa x 0.02 K cell, ties to the larger node, clamped -- as adiabat_lookup() in
// xp_column.cuh but with the node values taken as exact multiples (they are the doubles nearest to
// them, so the two differ only for an LCL within 3e-14 of a cell edge).  `edge` returns the distance
// (in cell units, 0..0.5) of the LCL from the nearest cell edge, so that callers whose LCL is only
// approximately known can hand edge cases to the exact path.
XP_HD int adiabat_cell(const Tables &tb, double p, double t, float &edge) {
    const double xs = (p - 2.5) * 2.0, ts = (t - 173.0) * 50.0;
    const double fp = floor(xs + 0.5), ft = floor(ts + 0.5);
    const int ip = min(max((int)fp, 0), kNP - 1), it = min(max((int)ft, 0), kNT - 1);
    edge = (float)fmin(0.5 - fabs(xs - fp), 0.5 - fabs(ts - ft));
    return (int)XP_LDG(tb.index_grid + (size_t)(kNP - 1 - ip) * kNT + it);
}

// ---- one parcel: constants + sweep state, all in registers ------------------------------------------------
//
// Row schedule.  The profile of a parcel is: start row (parcel == environment), the levels below
// the LCL, the inserted LCL row (PF:858-931), the levels above it.  The sweep runs over ITERATIONS
// it = 1 .. n_table shared by the three parcels: at iteration `it` a parcel whose first level above
// the LCL is `ka` processes
//      it <  ka : level it        (dry adiabat)
//      it == ka : the LCL row
//      it >  ka : level it - 1    (moist adiabat; the parcel "lags" one level behind)
// so every lane executes the same instruction stream whatever its LCL height, the environment of
// level it / it-1 is shared by the parcels, and "above the LCL" is simply `it > ka`.
struct FParcel {
    // constants
    float c_dryv;               // dry-adiabat curve value = c_dryv * p^kappa (virtual temperature folded in)
    float f;                    // position inside the cubic interval of the adiabat
    int m;                      // cubic interval
    int ka;                     // first level above the LCL = iteration of the LCL row
    int kfirst;                 // first iteration at which this parcel processes a row
    float x_lcl, a_lcl, b_lcl;  // the inserted LCL row
    float lcl_p, lcl_t, lcl_tv;
    bool bad;                   // must go to the exact path
    // profile output only (dead otherwise): unfolded dry adiabat, parcel mixing ratio, environment at the LCL
    float c_dry, w_par, lcl_env_t, lcl_env_td, lcl_env_tv;
    // sweep state (float32 version of Sweep in xp_column.cuh)
    float xprev, dprev, aprev;  // ln p, parcel - environment, parcel curve at the previous row
    float pos, tot;             // running sums of the positive areas and of all areas (ln p units)
    float lcl_pos, lcl_tot;
    float lfc_pos, lfc_tot, lfc_x, lfc_t;
    float el_pos, el_tot, el_x, el_t;
    int lfc_it, el_it;          // iteration of the upper row of the interval with the LFC / EL crossing (0: none)
    int n_inc;                  // increasing crossings seen (PF:1161)
    float min_abs_d;            // min |parcel - environment| over the rows    -> decision margin
    float min_slope;            // min over crossings of |d0 - d1| - kCrossSlope dx -> crossing placement margin
    float max_d_above;          // max (parcel - environment) over rows above the LCL (PF:1166-1169)
};

XP_HD float f_qnan() {
#if defined(__CUDACC__)
    return __int_as_float(0x7fffffff);
#else
    return std::nanf("");
#endif
}

XP_HD void sweep_init(FParcel &c, float x0, float a0) {
    c.xprev = x0; c.dprev = 0.0f; c.aprev = a0;        // start row: parcel == environment (PF:1117-1120)
    c.pos = c.tot = c.lcl_pos = c.lcl_tot = 0.0f;
    c.lfc_pos = c.lfc_tot = c.lfc_x = c.lfc_t = 0.0f;
    c.el_pos = c.el_tot = c.el_x = c.el_t = 0.0f;
    c.lfc_it = c.el_it = 0; c.n_inc = 0;
    c.min_abs_d = 1e30f; c.min_slope = 1e30f; c.max_d_above = -1e30f;
}

// One row (ln p = x, parcel curve a, environment curve b): find_intersections PF:992-1064 +
// trap_around_zeros PF:1200-1289 + trapz PF:164-206 for the interval below it.  Branch-free.
template <int MODE>
XP_HD void sweep_step(FParcel &c, int it, float x, float a, float b, bool is_lcl, bool above) {
    const float d = a - b;
    const float dx = c.xprev - x;
    const float den = c.dprev - d;
    const bool cross = c.dprev * d < 0.0f;
    const float fr = c.dprev * f_rcp(den);
    const float frac = cross ? fr : 1.0f;                        // zero at xprev - frac dx
    const float g = cross ? 1.0f - fr : 1.0f;
    const float h = 0.5f * dx;
    const float a_lo = h * c.dprev * frac;
    const float a_hi = h * d * g;
    c.pos += fmaxf(a_lo, 0.0f); c.tot += a_lo;
    const float ix = f_fma(-frac, dx, c.xprev);
    const float iy = f_fma(frac, a - c.aprev, c.aprev);
    const bool inc = cross && d > 0.0f;                          // PF:1058
    const bool dec = cross && !(d > 0.0f);                       // PF:1060
    c.n_inc += inc ? 1 : 0;
    // LFC: max-pressure increasing crossing above the LCL (PF:1127-1132) = the first one met
    const bool take_lfc = inc && above && c.lfc_it == 0;
    c.lfc_it = take_lfc ? it : c.lfc_it;
    c.lfc_pos = take_lfc ? c.pos : c.lfc_pos; c.lfc_tot = take_lfc ? c.tot : c.lfc_tot;
    c.lfc_x = take_lfc ? ix : c.lfc_x; c.lfc_t = take_lfc ? iy : c.lfc_t;
    // EL: min-pressure decreasing crossing (PF:1136) = the last one met
    c.el_it = dec ? it : c.el_it;
    c.el_pos = dec ? c.pos : c.el_pos;
    if (MODE != 1) c.el_tot = dec ? c.tot : c.el_tot;         // only needed without pos_cape_neg_cin
    c.el_x = dec ? ix : c.el_x; c.el_t = dec ? iy : c.el_t;
    c.pos += fmaxf(a_hi, 0.0f); c.tot += a_hi;
    c.lcl_pos = is_lcl ? c.pos : c.lcl_pos; c.lcl_tot = is_lcl ? c.tot : c.lcl_tot;
    c.max_d_above = fmaxf(c.max_d_above, above ? d : -1e30f);
    c.min_abs_d = fminf(c.min_abs_d, fabsf(d));
    c.min_slope = fminf(c.min_slope, cross ? f_fma(-kCrossSlope, dx, fabsf(den)) : 1e30f);
    c.xprev = x; c.dprev = d; c.aprev = a;
}

struct FResult {
    float cape, cin, lcl_p, lcl_t, lcl_tv, lfc_p, lfc_t, el_p, el_t, par_p, par_t, par_td;
    int shift;
};

// Sweep.finish of xp_column.cuh (lfc_el PF:1140-1185, cape_cin_base PF:1329-1388) on the float32 state.
template <int MODE>
XP_HD void sweep_finish(const FParcel &s, const Opts &o, FResult &r) {
    const float lcl_targ = o.vtc ? s.lcl_tv : s.lcl_t;
    const bool top_colder = s.dprev <= 0.0f;                            // PF:1151
    const bool el_exists = top_colder && s.el_it > s.ka;                // PF:1152-1153 (EL above the LCL)
    const bool lfc_missing = s.n_inc == 0;                              // PF:1161
    const bool lfc_found = s.lfc_it != 0;
    const bool pos_parcel = s.max_d_above > 0.0f;
    const bool replace = (pos_parcel && lfc_missing) || (!lfc_missing && !lfc_found && el_exists);
    const bool have_lfc = lfc_found || replace;
    float l_pos = s.lfc_pos, l_tot = s.lfc_tot;
    r.lfc_p = lfc_found ? f_ex2(s.lfc_x * kLog2e) : f_qnan();
    r.lfc_t = lfc_found ? s.lfc_t : f_qnan();
    if (replace) { r.lfc_p = s.lcl_p; r.lfc_t = lcl_targ; l_pos = s.lcl_pos; l_tot = s.lcl_tot; }
    r.el_p = el_exists ? f_ex2(s.el_x * kLog2e) : f_qnan();
    r.el_t = el_exists ? s.el_t : f_qnan();
    float cape = 0.0f, cin = 0.0f;
    if (have_lfc) {
        const float e_pos = el_exists ? s.el_pos : s.pos;
        const float e_tot = el_exists ? s.el_tot : s.tot;
        // EL below the LFC (PF:1352-1353 leaves no level between them): compare by interval
        const bool el_below_lfc = el_exists && lfc_found && !replace && s.el_it < s.lfc_it;
        if (MODE == 1 || o.pos_neg) {
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
    r.lcl_p = s.lcl_p; r.lcl_t = s.lcl_t; r.lcl_tv = s.lcl_tv;
}

// Set up one parcel once its (p0, T0, Td0) are known.  `kstart` is the level of the start row (the
// parcel level; for the mixed layer the start row is the prepended parcel at p[0]) and the lifted
// column continues at level `knext`.  T/Td of the levels bracketing the LCL come through `rd`.
// `approx`: the parcel itself carries float32-level errors (mixed-layer means of float32 mixing
// ratios): LCL decisions closer than kApproxCellEdge / kApproxLclP to an edge go to the exact path.
template <class Rd>
XP_HD void setup_parcel(const Rd &rd, const Prep &pr, const Tables &tb, const Opts &o, double p0, double t0,
                        double td0, int kstart, int knext, bool approx, FParcel &pc) {
    pc.bad = false;
    pc.kfirst = knext;
    // saturated / supersaturated / NaN parcels: the LCL snaps to the parcel level (np.isclose in
    // metpy.calc.lcl) or lies below it -- zero-width intervals, exact equality tests: exact path.
    if (!(t0 - td0 >= kSaturationMargin) || !(p0 > 0.0)) { pc.bad = true; t0 = 280.0; td0 = 270.0; p0 = 1000.0; }
    double lp, lt;
    lcl_fast(p0, t0, td0, lp, lt);
    float edge;
    const int adiabat = adiabat_cell(tb, lp, lt, edge);                // PF:554-557
    if (approx && edge < kApproxCellEdge) pc.bad = true;
    const int a0 = adiabat - 1;
    pc.m = a0 / kNodeStride;
    if (adiabat <= 0 || pc.m < kFirstInterval || pc.m > kLastInterval) { pc.bad = true; pc.m = kFirstInterval; }
    pc.f = (float)(a0 - pc.m * kNodeStride) * (1.0f / kNodeStride);
    const float p0f = (float)p0, t0f = (float)t0, td0f = (float)td0;
    const float lpf = (float)lp, ltf = (float)lt;
    pc.lcl_p = lpf; pc.lcl_t = ltf;
    const float es_l = f_es(ltf);
    pc.lcl_tv = f_tv(ltf, f_mixing_ratio(es_l, es_l, lpf, o.compat));           // PF:653-657
    const float w_parcel = f_mixing_ratio(f_es(t0f), f_es(td0f), p0f, o.compat);   // PF:748
    const float c_dry = t0f * f_rcp(pr.pk[kstart]);                             // p0 == p[kstart]; PF:291-316
    pc.c_dryv = o.vtc ? c_dry * f_fma(0.608f, w_parcel, 1.0f) : c_dry;          // PF:767, 773, 775
    sweep_init(pc, pr.lnp[kstart], o.vtc ? f_tv(t0f, w_parcel) : t0f);
    // LCL position among the levels of the lifted column (insert_level PF:965-966): exact in float64
    int ka = knext;
    while (ka < pr.L && pr.p64[ka] >= lp) ++ka;
    pc.ka = ka;
    if (approx && ((ka < pr.L && fabs(pr.p64[ka] - lp) < kApproxLclP) ||
                   (ka > knext && fabs(pr.p64[ka - 1] - lp) < kApproxLclP))) pc.bad = true;
    // bracketing levels for the environment at the LCL (PF:1774-1806): "before" = last level with
    // p >= lcl_p (the start row when the LCL is below the first swept level), "after" = level ka.
    const int kb = ka - 1;
    const bool before_is_start = (ka == knext);
    pc.x_lcl = pr.lnp[kstart]; pc.a_lcl = pc.b_lcl = 0.0f;
    if (ka >= pr.n_table || (!before_is_start && pr.p64[kb] == lp) || (before_is_start && p0 == lp)) {
        pc.bad = true;      // LCL above the table top / exactly on a level: exact path
        pc.ka = pr.n_table;
        return;
    }
    float tb_, tdb, xb;
    if (before_is_start) { tb_ = t0f; tdb = td0f; xb = o.log_interp ? pr.lnp[kstart] : pr.p[kstart]; }
    else { tb_ = rd.T(kb); tdb = rd.Td(kb); xb = o.log_interp ? pr.lnp[kb] : pr.p[kb]; }
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

// One parcel, one iteration of the shared sweep (see the row schedule above FParcel).
// `cprev` = coefficient row of level it-1; *_cur / *_prv = level it / it-1.
template <int MODE, class CoefRow>
XP_HD void parcel_iteration(FParcel &c, int it, bool last, const CoefRow &cprev, float pk_cur, float p_prv,
                            float x_cur, float x_prv, float b_cur, float b_prv, bool vtc) {
    const bool above = it > c.ka;          // the row is above the LCL (and lags one level)
    const bool is_lcl = it == c.ka;
    if (it < c.kfirst || (last && !above)) return;
    // moist adiabat at level it-1 (PF:585-592) and its saturation mixing ratio (PF:760)
    const Coef cc = cprev.at(c.m);
    const float tm = f_fma(f_fma(f_fma(cc.c3, c.f, cc.c2), c.f, cc.c1), c.f, cc.c0);
    const float es = f_es(tm);
    const float a_m = vtc ? f_tv(tm, kEpsF * es * f_rcp(p_prv - es)) : tm;
    const float a_d = c.c_dryv * pk_cur;                                        // PF:742
    const float a = is_lcl ? c.a_lcl : (above ? a_m : a_d);
    const float b = is_lcl ? c.b_lcl : (above ? b_prv : b_cur);
    const float x = is_lcl ? c.x_lcl : (above ? x_prv : x_cur);
    sweep_step<MODE>(c, it, x, a, b, is_lcl, above);
}

// The whole suite for one column.  Rd: float T(k), Td(k).  Cf: row(k) -> object with Coef at(m).
// KINDS: bit 0 SB, 1 ML, 2 MU.  MODE 1: the reference's default options (virtual temperature
// correction, MetPy 1.4.1 formulas, pos_cape_neg_cin) folded in at compile time; MODE 0: run time.
// Returns the mask of kinds that the exact path must recompute.
//
// Env: where the environment curve lives between the first pass and the sweep.
//   Env::kStaged = true : `env.put(k, b)` / `env.get(k)` is a per-thread column in shared memory; the first
//     pass runs over ALL levels (reading T/Td once), stores the environment curve and its minimum, and the
//     sweep (a) reads it back instead of recomputing it and (b) stops -- warp-uniformly -- as soon as every
//     parcel of every lane is above its LCL and colder than the coldest environment level by
//     kStopMargin: no crossing and no positive area can follow, so with pos_cape_neg_cin (MODE 1)
//     nothing the outputs depend on changes any more.
//   Env::kStaged = false: the environment is recomputed in the sweep from T/Td (no shared memory needed).
template <unsigned KINDS, int MODE, class Rd, class Cf, class Env>
XP_HD unsigned suite_column(const Rd &rd, const Cf &cf, const Prep &pr, const Tables &tb, const Opts &o,
                            Env &env, FResult res[3]) {
    unsigned redo = 0;
    float nanacc = 0.0f;                   // becomes NaN if any T/Td read is NaN or infinite
    const bool vtc = (MODE == 1) ? true : (o.vtc != 0);
    const int compat = (MODE == 1) ? 141 : o.compat;
    const int nt = pr.n_table;
    float b_min = 1e30f;
    // ---- pre-pass over the lowest levels: mixed-layer means (float64) and most-unstable argmax ----
    double sum_th = 0.0, sum_w = 0.0;
    float best = -1e30f, second = -1e30f, mu_t = 0.0f, mu_td = 0.0f;
    int k_mu = 0;
    const int n_low = max((KINDS & 4u) ? pr.K_mu : 0, (KINDS & 2u) ? pr.n_ml_w : 0);
    const int n_pre = Env::kFullPass ? nt : n_low;
    const float *tp0 = rd.tptr(0), *tdp0 = rd.tdptr(0);
    const int64_t ls = rd.stride();
    // T/Td are fetched kPre levels ahead (this pass has little arithmetic per level to hide the
    // HBM latency behind): a ring of registers refilled one chunk at a time
    constexpr int kPre = 4;
    float tq[kPre], tdq[kPre], tn[kPre], tdn[kPre];
#pragma unroll
    for (int j = 0; j < kPre; ++j) {
        tn[j] = tdn[j] = 0.0f;
        if (j < n_pre) { tn[j] = Rd::ld(tp0 + (int64_t)j * ls); tdn[j] = Rd::ld(tdp0 + (int64_t)j * ls); }
    }
    for (int k0 = 0; k0 < n_pre; k0 += kPre) {
#pragma unroll
        for (int j = 0; j < kPre; ++j) { tq[j] = tn[j]; tdq[j] = tdn[j]; }
        tp0 += (int64_t)kPre * ls; tdp0 += (int64_t)kPre * ls;
#pragma unroll
        for (int j = 0; j < kPre; ++j)
            if (k0 + kPre + j < n_pre) { tn[j] = Rd::ld(tp0 + (int64_t)j * ls); tdn[j] = Rd::ld(tdp0 + (int64_t)j * ls); }
#pragma unroll
      for (int j = 0; j < kPre; ++j) {
        const int k = k0 + j;
        if (k >= n_pre) break;
        const float t = tq[j], td = tdq[j];
        nanacc = f_fma(t, 0.0f, f_fma(td, 0.0f, nanacc));
        const float p = pr.p[k];
        const float e = f_es(td);
        if (Env::kFullPass) {                                                     // environment curve, PF:839-843
            const float b = vtc ? f_tv(t, f_mixing_ratio(f_es(t), e, p, compat)) : t;
            if (Env::kStaged) env.put(k, b);
            b_min = fminf(b_min, b);
        }
        if (k >= n_low) continue;
        const float ipe = f_rcp(p - e);
        const float r = kEpsF * e * ipe;                 // saturation mixing ratio of the dewpoint (PF:258)
        if ((KINDS & 2u) && k < pr.n_ml_w) {
            // mixed_parcel PF:253-258: theta and saturation mixing ratio of the dewpoint in float64, so
            // that the table cell of this parcel's LCL is the reference's (a float32 mixing ratio would
            // put ~1.7 % of the mixed-layer parcels within its error of a cell edge)
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
    if (KINDS & 1u) {
        setup_parcel(rd, pr, tb, o, pr.p0, (double)t_sfc, (double)td_sfc, 0, 1, false, sb);
        res[0].par_p = pr.p[0]; res[0].par_t = t_sfc; res[0].par_td = td_sfc; res[0].shift = 0;
    }
    if (KINDS & 2u) {
        const double mp_t = sum_th * pr.exner0;                                  // PF:268-269
        const double mp_td = dewpoint_from_e(vapor_pressure(pr.p0, sum_w));      // PF:275-282
        setup_parcel(rd, pr, tb, o, pr.p0, mp_t, mp_td, 0, pr.K_ml, false, ml);
        res[1].par_p = pr.p[0]; res[1].par_t = (float)mp_t; res[1].par_td = (float)mp_td; res[1].shift = pr.K_ml;
    }
    if (KINDS & 4u) {
        if (!(best - second >= kThetaEMargin)) redo |= 4u;                       // argmax within float32 error
        setup_parcel(rd, pr, tb, o, pr.p64[k_mu], (double)mu_t, (double)mu_td, k_mu, k_mu + 1, false, mu);
        res[2].par_p = pr.p[k_mu]; res[2].par_t = mu_t; res[2].par_td = mu_td; res[2].shift = k_mu;
    }
    // ---- the sweep --------------------------------------------------------------------------------------
    const float *lp_p = pr.p + 1, *lp_x = pr.lnp + 1, *lp_k = pr.pk + 1;      // level `it` of the axis constants
    float b_prv = 0.0f, x_prv = pr.lnp[0], p_prv = pr.p[0];
    const float *tp = rd.tptr(1), *tdp = rd.tdptr(1);
    float t_nxt = 0.0f, td_nxt = 0.0f;
    if (!Env::kStaged) { t_nxt = Rd::ld(tp); td_nxt = Rd::ld(tdp); }
    auto crow = cf.row(0);
    const float stop_below = b_min - kStopMargin;
    for (int it = 1; it <= nt; ++it) {
        const bool last = (it == nt);
        float b_cur = 0.0f, x_cur = x_prv, pk_cur = 0.0f, p_cur = p_prv;
        if (Env::kStaged) {
            if (!last) { p_cur = *lp_p++; x_cur = *lp_x++; pk_cur = *lp_k++; b_cur = env.get(it); }
        } else {
            const float t = t_nxt, td = td_nxt;
            tp += ls; tdp += ls;
            if (it + 1 < nt) { t_nxt = Rd::ld(tp); td_nxt = Rd::ld(tdp); }      // prefetch the next level
            if (!last) {
                nanacc = f_fma(t, 0.0f, f_fma(td, 0.0f, nanacc));
                p_cur = *lp_p++; x_cur = *lp_x++; pk_cur = *lp_k++;
                b_cur = vtc ? f_tv(t, f_mixing_ratio(f_es(t), f_es(td), p_cur, compat)) : t;   // PF:839-843
            }
        }
        if (KINDS & 1u) parcel_iteration<MODE>(sb, it, last, crow, pk_cur, p_prv, x_cur, x_prv, b_cur, b_prv, vtc);
        if (KINDS & 2u) parcel_iteration<MODE>(ml, it, last, crow, pk_cur, p_prv, x_cur, x_prv, b_cur, b_prv, vtc);
        if (KINDS & 4u) parcel_iteration<MODE>(mu, it, last, crow, pk_cur, p_prv, x_cur, x_prv, b_cur, b_prv, vtc);
        b_prv = b_cur; x_prv = x_cur; p_prv = p_cur;
        crow.advance();
        if (Env::kFullPass && MODE == 1) {
            // a parcel is finished when it is above its LCL and colder than every environment level
            // (its curve only cools with height); parcels already bound for the exact path do not count
            bool done = true;
            if (KINDS & 1u) done = done && (sb.bad || (it > sb.ka && sb.aprev < stop_below));
            if (KINDS & 2u) done = done && (ml.bad || (it > ml.ka && ml.aprev < stop_below));
            if (KINDS & 4u) done = done && (mu.bad || (it > mu.ka && it > mu.kfirst && mu.aprev < stop_below));
#if defined(XP_HOST_SIM) && defined(XP_DEBUG_STOP)
            if (done) { XP_DEBUG_STOP(it); }
#endif
            if ((it & o.vote_mask) == 0 && XP_WARP_ALL(done)) break;
        }
    }
    // ---- results ---------------------------------------------------------------------------------------
    const bool nan_seen = !(nanacc == 0.0f);
    auto wrap = [&](const FParcel &c, FResult &r, unsigned bit) {
        sweep_finish<MODE>(c, o, r);
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
