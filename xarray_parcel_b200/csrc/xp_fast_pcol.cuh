// xp_fast_pcol.cuh -- float32 fast path for columns with PER-COLUMN pressure (model levels:
// BASELINE.json configs[1], [2]; the layout of the reference's own test_data.nc).
//
// Same design as xp_fast.cuh (lagged-row sweep shared by the parcels, float64-polished LCL, exact
// table cell, decisions never taken inside the float32 margin, uncertain columns handed to the
// float64 exact kernel) with two differences forced by the per-column pressure:
//  * ln p and p^kappa of a level are computed per thread (2 MUFU), and the mixed-layer / most-
//    unstable layer bounds are found per column in float64;
//  * the moist adiabat of a parcel is read straight from the float32 curve table in global memory
//    (L2-resident), exactly as the reference evaluates it: two neighbouring 0.5 hPa nodes of adiabat
//    i and a linear interpolation in p (PF:585-592) -- no approximation beyond float32 rounding.
#pragma once
#include "xp_fast.cuh"

#ifndef XP_WARP_MAX_INT
#if defined(__CUDACC__)
#define XP_WARP_MAX_INT(x) __reduce_max_sync(__activemask(), (x))
#else
namespace xp { namespace fast { inline int &host_warp_max_floor() { static int v = 0; return v; } } }
#define XP_WARP_MAX_INT(x) ((x) > xp::fast::host_warp_max_floor() ? (x) : xp::fast::host_warp_max_floor())
#endif
#endif

namespace xp {
namespace fast {

// ---- the adiabat family as a table in SHARED memory, for arbitrary level pressures ----------------------------------------
// Per-column pressure has no per-level row of cubics (xp_fast.cuh); the generic kernels gather two 0.5 hPa nodes of the
// parcel's adiabat from the 125 MB curve table per (level, parcel) -- L2-latency-bound (r1c profiles: 43-52 % issue
// utilisation, 3.5-5 long-scoreboard stalls per issue).  PTab is the same family on a grid that fits shared memory:
// virtual temperature of the saturated parcel (PF:760/775, default options) as
//     cubic across every 64th adiabat (as compute_coef_tv)  x  4-point Lagrange across pressure NODES that are uniform
//     in xi = p^kappa (the dry adiabat is linear in xi, the moist one nearly so aloft),
// two segments: A = 100..1100 hPa, 50 nodes (max |table - reference evaluation| 3.7e-4 K, for pseudo-adiabats hotter
// than 312 K at 1100 hPa next to 100 hPa; ~1e-4 K otherwise); B = 2.5..100 hPa, 24 nodes (<= 2.3e-3 K down to 20 hPa,
// <= 0.05 K above: nothing is decided that finely up there).  The decision margin of a row scales with its segment
// (kPTabMarginA / B / Top); tests/test_fast_hostsim.py::test_ptab_accuracy holds the table to those bounds.  Adiabat intervals below
// kPTabFirstInterval (pseudo-adiabats colder than 230 K at 1100 hPa) are left out: such parcels take the exact path.
constexpr int kPTabNodesA = 50, kPTabNodesB = 24, kPTabNodes = kPTabNodesA + kPTabNodesB;
constexpr int kPTabFirstInterval = 89;                              // (230 - 173) / 0.01 / 64
constexpr int kPTabIntervals = kLastInterval + 1 - kPTabFirstInterval;   // 133 -> 64 x 133 x 16 B = 136 KB
// (the node grids overlap around the 100 hPa split so that the four-node stencils next to it stay centred)
constexpr double kPTabSplit = 100.0, kPTabTop = 2.5, kPTabBottom = 1100.0, kPTabLowA = 88.0, kPTabHighB = 125.0;
constexpr float kPTabMarginA = 8e-4f, kPTabMarginB = 5e-3f, kPTabMarginTop = 0.1f;   // K (kDecisionEps is 6e-4)

struct PTabDesc {               // uniform-in-xi node grids of the two segments
    float x0a, idxa, x0b, idxb, xsplit, xtop20;
};
XP_HD PTabDesc ptab_desc() {
    PTabDesc d;
    const double xa0 = pow(kPTabLowA, kKappa), xa1 = pow(kPTabBottom, kKappa);
    const double xb0 = pow(kPTabTop, kKappa), xb1 = pow(kPTabHighB, kKappa);
    d.x0a = (float)xa0; d.idxa = (float)((kPTabNodesA - 1) / (xa1 - xa0));
    d.x0b = (float)xb0; d.idxb = (float)((kPTabNodesB - 1) / (xb1 - xb0));
    d.xsplit = (float)pow(kPTabSplit, kKappa); d.xtop20 = (float)pow(20.0, kKappa);
    return d;
}
// pressure of node `j` (0 .. kPTabNodes-1; segment B first, ascending pressure)
XP_HD double ptab_node_pressure(int j) {
    const double xa0 = pow(kPTabLowA, kKappa), xa1 = pow(kPTabBottom, kKappa);
    const double xb0 = pow(kPTabTop, kKappa), xb1 = pow(kPTabHighB, kKappa);
    if (j < kPTabNodesB) {
        if (j == 0) return kPTabTop;
        if (j == kPTabNodesB - 1) return kPTabHighB;
        return pow(xb0 + (xb1 - xb0) * j / (kPTabNodesB - 1), kInvKappa);
    }
    j -= kPTabNodesB;
    if (j == 0) return kPTabLowA;
    if (j == kPTabNodesA - 1) return kPTabBottom;
    return pow(xa0 + (xa1 - xa0) * j / (kPTabNodesA - 1), kInvKappa);
}
// coefficient (node j, interval m) of the table: as compute_coef_tv, at the node's pressure
XP_HD Coef compute_ptab_coef(const float *curves, int j, int m) {
    const double p = ptab_node_pressure(j);
    double y[4];
    for (int q = 0; q < 4; ++q) {
        int a = (m - 1 + q) * kNodeStride;
        a = min(max(a, 0), kNAdiabats - 1);
        const double t = adiabat_temperature(curves + (size_t)a * kNP, p);
        y[q] = virtual_temperature(t, sat_mixing_ratio(p, t));                  // PF:760, 775
    }
    Coef c;
    c.c0 = (float)y[1];
    c.c1 = (float)(-y[0] / 3 - y[1] / 2 + y[2] - y[3] / 6);
    c.c2 = (float)(y[0] / 2 - y[1] + y[2] / 2);
    c.c3 = (float)(-y[0] / 6 + y[1] / 2 - y[2] / 2 + y[3] / 6);
    return c;
}

// What a level contributes to every parcel's table evaluation: the first of the four node rows and the Lagrange
// weights in xi, plus the factor that scales |parcel - environment| before it is compared with kDecisionEps.
struct PTabLevel {
    const Coef *row;            // node j-1 (rows are kPTabIntervals apart)
    float w0, w1, w2, w3;
    float margin_scale;
};
struct PTabView {
    static constexpr bool kTable = true;
    const Coef *base;           // [kPTabNodes][kPTabIntervals]
    PTabDesc d;
    XP_HD PTabLevel level(float xi) const {
        const bool seg_a = xi >= d.xsplit;
        const float s = (xi - (seg_a ? d.x0a : d.x0b)) * (seg_a ? d.idxa : d.idxb);
        const int n = seg_a ? kPTabNodesA : kPTabNodesB;
        const int j = min(max((int)s, 1), n - 3);                  // the stencil j-1 .. j+2 stays inside the segment
        const float u = s - (float)j;
        PTabLevel lv;
        lv.row = base + (size_t)((seg_a ? kPTabNodesB : 0) + j - 1) * kPTabIntervals;
        const float um = u - 1.0f, up = u + 1.0f, u2 = u - 2.0f;
        lv.w0 = (-1.0f / 6.0f) * u * um * u2;
        lv.w1 = 0.5f * up * um * u2;
        lv.w2 = -0.5f * up * u * u2;
        lv.w3 = (1.0f / 6.0f) * up * u * um;
        lv.margin_scale = seg_a ? kDecisionEps / kPTabMarginA
                                : ((xi >= d.xtop20) ? kDecisionEps / kPTabMarginB : kDecisionEps / kPTabMarginTop);
        return lv;
    }
    // virtual temperature of the saturated parcel on adiabat (interval m, position f) at the level
    XP_HD float eval(const PTabLevel &lv, int m, float f) const {
        const Coef *r = lv.row + (m - kPTabFirstInterval);
#if defined(XP_BOUNDS_CHECK) && defined(__CUDA_ARCH__)
        if (m < kPTabFirstInterval || m > kLastInterval || r < base ||
            r + 3 * kPTabIntervals >= base + (size_t)kPTabNodes * kPTabIntervals) {
            printf("XP_BOUNDS_CHECK failed: PTabView::eval m=%d row offset %ld\n", m, (long)(lv.row - base));
            __trap();
        }
#endif
        const Coef c0 = r[0], c1 = r[kPTabIntervals], c2 = r[2 * kPTabIntervals], c3 = r[3 * kPTabIntervals];
        const float v0 = f_fma(f_fma(f_fma(c0.c3, f, c0.c2), f, c0.c1), f, c0.c0);
        const float v1 = f_fma(f_fma(f_fma(c1.c3, f, c1.c2), f, c1.c1), f, c1.c0);
        const float v2 = f_fma(f_fma(f_fma(c2.c3, f, c2.c2), f, c2.c1), f, c2.c0);
        const float v3 = f_fma(f_fma(f_fma(c3.c3, f, c3.c2), f, c3.c1), f, c3.c0);
        return f_fma(lv.w3, v3, f_fma(lv.w2, v2, f_fma(lv.w1, v1, lv.w0 * v0)));
    }
};
struct NoPTab {                 // the generic kernels: adiabats gathered from the curve table in global memory
    static constexpr bool kTable = false;
    XP_HD PTabLevel level(float) const { return PTabLevel(); }
    XP_HD float eval(const PTabLevel &, int, float) const { return 0.0f; }
};

// np.interp(p, P_ascending, curve) in float32 (PF:585-600); p inside [2.5, 1100] is a precondition.
XP_HD float adiabat_temperature_f32(const float *__restrict__ curve, int j, float w) {
    const float f0 = XP_LDG(curve + j), f1 = XP_LDG(curve + j + 1);
    return f_fma(f1 - f0, w, f0);
}

struct PColParcel : FParcel {
    const float *curve;         // row of this parcel's adiabat in the curve table (ascending pressure)
    float f0, f1;               // the two table nodes bracketing the pressure of the previous level,
                                // gathered one iteration ahead so that the L2 latency is hidden
};

// Per-parcel set-up for per-column pressure.  Rd: float P(k), T(k), Td(k).
// qm != 0: the dewpoint array holds specific humidity in that MetPy form (levels converted as they are read).
template <class Rd>
XP_HD void setup_parcel_pcol(const Rd &rd, int L, const Tables &tb, const Opts &o, double p0, double t0,
                             double td0, float x_start, int knext, PColParcel &pc, int qm = 0, bool ptab = false) {
    pc.bad = false;
    pc.kfirst = knext;
    if (!(t0 - td0 >= kSaturationMargin) || !(p0 > 0.0)) { pc.bad = true; t0 = 280.0; td0 = 270.0; p0 = 1000.0; }
    double lp, lt;
    lcl_fast(p0, t0, td0, lp, lt);
    float edge;
    const int adiabat = adiabat_cell(tb, lp, lt, edge);                // PF:554-557
    if (adiabat <= 0) pc.bad = true;
    pc.curve = tb.curves + (size_t)(adiabat > 0 ? adiabat - 1 : 0) * kNP;
    pc.m = 0; pc.f = 0.0f; pc.f0 = pc.f1 = 0.0f;
    if (ptab) {                 // shared-memory table (PTabView): cubic interval and position, as setup6_c
        const int a0 = adiabat - 1;
        pc.m = a0 / kNodeStride;
        if (adiabat <= 0 || pc.m < kPTabFirstInterval || pc.m > kLastInterval) { pc.bad = true; pc.m = kPTabFirstInterval; }
        pc.f = (float)(a0 - pc.m * kNodeStride) * (1.0f / kNodeStride);
    }
    const float p0f = (float)p0, t0f = (float)t0, td0f = (float)td0;
    const float lpf = (float)lp, ltf = (float)lt;
    pc.lcl_p = lpf; pc.lcl_t = ltf;
    const float es_l = f_es(ltf);
    pc.lcl_tv = f_tv(ltf, f_mixing_ratio(es_l, es_l, lpf, o.compat));           // PF:653-657
    const float w_parcel = f_mixing_ratio(f_es(t0f), f_es(td0f), p0f, o.compat);   // PF:748
    const float c_dry = t0f * f_ex2(-(float)kKappa * f_lg2(p0f));               // T0 / p0^kappa, PF:291-316
    pc.c_dryv = o.vtc ? c_dry * f_fma(0.608f, w_parcel, 1.0f) : c_dry;
    pc.c_dry = c_dry; pc.w_par = w_parcel;
    pc.lcl_env_t = pc.lcl_env_td = pc.lcl_env_tv = f_qnan();
    sweep_init(pc, x_start, o.vtc ? f_tv(t0f, w_parcel) : t0f);
    // LCL position among the levels of the lifted column (insert_level PF:965-966): exact in float64
    // (two memory round trips instead of one per level: eight independent probes four levels apart bracket the
    //  LCL, four more loads resolve the bracket; columns whose LCL lies more than 32 levels up continue in chunks)
    int ka = knext;
    double pka = 0.0, pkb = p0;
    bool found = false;
    {
        float pq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pq[j] = (ka + 4 * j < L) ? rd.P(ka + 4 * j) : 0.0f;
        int jc = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) jc += (jc == j && ka + 4 * j < L && (double)pq[j] >= lp) ? 1 : 0;
        // probes 0 .. jc-1 are at or below the LCL (p >= lcl_p); the LCL level is among the four levels after probe jc-1
        if (jc == 0) {
            if (ka < L) { pka = (double)pq[0]; found = true; }
        } else {
            float pl = pq[0];
#pragma unroll
            for (int j = 1; j < 8; ++j) pl = (j == jc - 1) ? pq[j] : pl;
            ka += 4 * (jc - 1);                               // a level with p >= lcl_p
            pkb = (double)pl; ++ka;
        }
    }
    while (ka < L && !found) {
        float pq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pq[j] = (ka + j < L) ? rd.P(ka + j) : 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (found || ka >= L) continue;
            pka = (double)pq[j];
            if (!(pka >= lp)) found = true;
            else { pkb = pka; ++ka; }
        }
    }
    pc.ka = ka;
    const bool before_is_start = (ka == knext);
    pc.x_lcl = x_start; pc.a_lcl = pc.b_lcl = 0.0f;
    if (ka >= L || pkb == lp) { pc.bad = true; pc.ka = L; return; }            // LCL above the top / on a level
    float tb_, tdb;
    if (before_is_start) { tb_ = t0f; tdb = td0f; }
    else { tb_ = rd.T(ka - 1); tdb = rd.Td(ka - 1); if (qm) tdb = f_td_from_q((float)pkb, tb_, tdb, qm); }
    const float ta = rd.T(ka);
    float tda = rd.Td(ka);
    if (qm) tda = f_td_from_q((float)pka, ta, tda, qm);                // the array holds specific humidity
    const float x_l = kLn2 * f_lg2(lpf);
    float xb, xa, at;
    if (o.log_interp) { xb = kLn2 * f_lg2((float)pkb); xa = kLn2 * f_lg2((float)pka); at = x_l; }
    else { xb = (float)pkb; xa = (float)pka; at = lpf; }
    const float g = (at - xb) * f_rcp(xa - xb);
    const float te = f_fma(ta - tb_, g, tb_), tde = f_fma(tda - tdb, g, tdb);  // PF:1802
    const float etv = f_tv(te, f_mixing_ratio(f_es(te), f_es(tde), lpf, o.compat));     // PF:916-920
    pc.x_lcl = x_l;
    pc.a_lcl = o.vtc ? pc.lcl_tv : ltf;
    pc.b_lcl = o.vtc ? etv : te;
    pc.lcl_env_t = te; pc.lcl_env_td = tde; pc.lcl_env_tv = etv;
    if (!(te == te) || !(tde == tde)) pc.bad = true;
}

// Per-thread ring of (p, T, Td) levels for the RE-BASED profile sweep (one lifted kind + profile rows, see
// suite_column_pcol): there iteration `it` of a lane consumes level it + shift with a per-lane shift, so direct loads
// put the 32 lanes of a warp on up to 32 different lines (ncu, 10 M x 90 most-unstable + rows: 38.6 GB read for
// 10.8 GB of input).  With a ring every lane loads the SAME level at the same time -- one coalesced line per array --
// into its own slots and consumes it `wmax - shift` iterations later; only the thread that wrote a slot reads it, so
// there is nothing to synchronise.  kRingLead levels are fetched beyond the furthest lane to cover the load latency.
constexpr int kRingLead = 4;
struct NoRing {
    static constexpr bool kEnabled = false;
    XP_HD int capacity() const { return 0; }
    XP_HD void put(int, float, float, float) const {}
    XP_HD void get(int, float &p, float &t, float &td) const { p = t = td = 0.0f; }
};

// Environment of a level as the profile rows need it.
struct EnvLevel { float p, t, td, tv; };

// One parcel, one iteration (row schedule of xp_fast.cuh).  j_cur: table node of the pressure of level
// `it` (gathered for the next iteration); w_prv: interpolation weight of level it-1.
template <int MODE, class Prof, class PTab>
XP_HD void parcel_iteration_pcol(PColParcel &c, int it, bool last, int j_cur, float w_prv, float pk_cur,
                                 float p_prv, float x_cur, float x_prv, float b_cur, float b_prv, bool vtc,
                                 const EnvLevel &e_cur, const EnvLevel &e_prv, int q, Prof &prof,
                                 const PTab &ptab, const PTabLevel &lv_prv) {
    const bool above = it > c.ka;
    const bool is_lcl = it == c.ka;
    float tm = 0.0f, tv_m;
    if (PTab::kTable) {
        if (it < c.kfirst || (last && !above)) return;
        tv_m = ptab.eval(lv_prv, c.m, c.f);                                    // shared-memory table (default options)
    } else {
        const float f0 = c.f0, f1 = c.f1;
        c.f0 = XP_LDG(c.curve + j_cur); c.f1 = XP_LDG(c.curve + j_cur + 1);     // for the next iteration
        if (it < c.kfirst || (last && !above)) return;
        tm = f_fma(f1 - f0, w_prv, f0);                                        // np.interp, PF:585-592
        const float es = f_es(tm);
        tv_m = f_tv(tm, kEpsF * es * f_rcp(p_prv - es));                        // PF:760, 775
    }
    const float a_m = vtc ? tv_m : tm;
    const float a_d = c.c_dryv * pk_cur;                                        // PF:742
    const float a = is_lcl ? c.a_lcl : (above ? a_m : a_d);
    const float b = is_lcl ? c.b_lcl : (above ? b_prv : b_cur);
    const float x = is_lcl ? c.x_lcl : (above ? x_prv : x_cur);
    if (Prof::kEnabled) {                                                       // PF:806-931 row of this iteration
        const int row = it - c.kfirst + 1;
        if (is_lcl) prof.put(q, row, c.lcl_p, c.lcl_t, c.lcl_tv, c.lcl_env_t, c.lcl_env_tv, c.lcl_env_td);
        else if (above) prof.put(q, row, e_prv.p, tm, tv_m, e_prv.t, e_prv.tv, e_prv.td);
        else { const float tp = c.c_dry * pk_cur; prof.put(q, row, e_cur.p, tp, f_tv(tp, c.w_par), e_cur.t, e_cur.tv, e_cur.td); }
    }
    sweep_step<MODE>(c, it, x, a, b, is_lcl, above);
    // rows read from the coarser part of the table (above 100 hPa) are decided with a wider margin
    if (PTab::kTable && above) c.min_abs_d = fminf(c.min_abs_d, fabsf(a - b) * lv_prv.margin_scale);
}

// The sweep of the shared-memory-table kernels (default options, scalar outputs): v7 steps (xp_fast7.cuh) on
// per-column pressure.  Defined in xp_fast_pcol7.cuh.
template <unsigned KINDS, class Rd, class PTab>
XP_HD unsigned ptab7_sweep(const Rd &rd, int L, const Opts &o, const PTab &ptab, PColParcel &sb, PColParcel &ml,
                           PColParcel &mu, FResult res[3], unsigned redo, float nanacc, bool bad_axis, float best,
                           float second, int k_mu, int qm, float p_sfc, float t_sfc, float td_sfc, float x_sfc);

// The suite for one column with its own pressure profile.  Returns the redo mask (see suite_column).
// QIN: the dewpoint array may hold specific humidity (o.qmode); false compiles the conversion away.
template <unsigned KINDS, int MODE, bool QIN, class Rd, class Prof, class Ring, class PTab>
XP_HD unsigned suite_column_pcol(const Rd &rd, int L, const Tables &tb, const Opts &o, Prof &prof, Ring &ring,
                                 const PTab &ptab, FResult res[3]) {
    static_assert(!PTab::kTable || (MODE == 1 && !Prof::kEnabled), "the shared-memory table holds the virtual temperature only");
    unsigned redo = 0, rows_exact = 0;       // rows_exact: kinds whose profile rows the exact path must rewrite too
    float nanacc = 0.0f;
    bool bad_axis = false;                 // pressure not finite / not strictly decreasing / outside the table
    // qm != 0: the dewpoint array holds specific humidity (xp_columns.dewpoint_is_specific_humidity): levels are
    // converted as they are loaded (float32), parcel dewpoints and mixed-layer vapour pressures in float64
    const int qm = QIN ? o.qmode : 0;
    const float p_sfc = rd.P(0), t_sfc = rd.T(0), raw_sfc = rd.Td(0);
    const double bottom = (double)p_sfc;                                         // PF:80 (pressure decreases upward)
    const double td_sfc64 = qm ? td64_from_q_fast(bottom, (double)t_sfc, (double)raw_sfc, qm) : (double)raw_sfc;
    const float td_sfc = (float)td_sfc64;
    nanacc = f_fma(p_sfc, 0.0f, f_fma(t_sfc, 0.0f, f_fma(td_sfc, 0.0f, nanacc)));
    if (!(p_sfc <= 1100.0f)) bad_axis = true;
    // ---- pre-pass: mixed-layer means (float64) and most-unstable argmax over the lowest levels ------------
    const double top_ml = bottom - o.ml_depth;                                   // PF:84, PF:1636
    const double bound_mu = bottom - o.mu_depth;                                 // PF:92
    double sum_th = 0.0, sum_w = 0.0, pp = bottom, thp = 0.0, wp = 0.0;
    int K_ml = L;
    bool ml_done = !(KINDS & 2u), mu_done = !(KINDS & 4u);
    float best = -1e30f, second = -1e30f, mu_t = 0.0f, mu_td = 0.0f, mu_p = p_sfc, mu_raw = raw_sfc;
    int k_mu = 0;
    float p_n1 = p_sfc, t_n1 = t_sfc, td_n1 = raw_sfc, p_n2 = 0.0f, t_n2 = 0.0f, td_n2 = 0.0f;
    if (1 < L) { p_n2 = rd.P(1); t_n2 = rd.T(1); td_n2 = rd.Td(1); }
    for (int k = 0; k < L && !(ml_done && mu_done); ++k) {
        // levels are read two iterations ahead (this loop's float64 body hides one memory round trip, not two)
        const float pf = p_n1, t = t_n1, raw = td_n1;
        const float td = qm ? f_td_from_q(pf, t, raw, qm) : raw;
        p_n1 = p_n2; t_n1 = t_n2; td_n1 = td_n2;
        if (k + 2 < L) { p_n2 = rd.P(k + 2); t_n2 = rd.T(k + 2); td_n2 = rd.Td(k + 2); }
        nanacc = f_fma(pf, 0.0f, f_fma(t, 0.0f, f_fma(td, 0.0f, nanacc)));
        const double p = (double)pf;
        if (k > 0 && !(p < pp)) bad_axis = true;
        if ((KINDS & 2u) && !ml_done) {
            // mixed_parcel PF:229-289 on get_layer(interpolate=True) PF:63-100, trapz in p PF:186-198
            // (branch-free float64 exp / log and Newton reciprocals of xp_fast.cuh: ~3 ulp, a third of the
            //  instructions of the libm calls)
            const double tdd = (double)td;
            const double e = qm ? e64_from_q_fast(p, (double)t, (double)raw, qm)
                                : kSat0 * exp64_fast(17.67 * (tdd - 273.15) * rcp64(tdd - 29.65));
            const double th = (double)t * pow_kappa64(1000.0 * rcp64(p));               // PF:253: T (1000 / p)^kappa
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
                if (v > best) { second = best; best = v; k_mu = k; mu_t = t; mu_td = td; mu_p = pf; mu_raw = raw; }
                else if (v > second) second = v;
            }
        }
        pp = p;
    }
    // ---- parcels ----------------------------------------------------------------------------------------
    PColParcel sb, ml, mu;
    const float x_sfc = kLn2 * f_lg2(p_sfc);
    if (KINDS & 1u) {
        setup_parcel_pcol(rd, L, tb, o, bottom, (double)t_sfc, td_sfc64, x_sfc, 1, sb, qm, PTab::kTable);
        res[0].par_p = p_sfc; res[0].par_t = t_sfc; res[0].par_td = td_sfc; res[0].shift = 0;
    }
    if (KINDS & 2u) {
        const double depth = fabs(top_ml - bottom);                              // PF:158-159
        double mp_t, mp_td;
        mixed_parcel_t_td(bottom, (1. / depth) * sum_th, (1. / depth) * sum_w, mp_t, mp_td);   // PF:161, 268-282
        if (!ml_done || K_ml < 1) { redo |= 2u; rows_exact |= 2u; K_ml = max(K_ml, 1); }   // no level above / NaN layer: exact path
        setup_parcel_pcol(rd, L, tb, o, bottom, mp_t, mp_td, x_sfc, K_ml, ml, qm, PTab::kTable);
        res[1].par_p = p_sfc; res[1].par_t = (float)mp_t; res[1].par_td = (float)mp_td; res[1].shift = K_ml;
    }
    if (KINDS & 4u) {
        if (!(best - second >= kThetaEMargin)) { redo |= 4u; rows_exact |= 4u; }
        const double mu_td64 = qm ? td64_from_q_fast((double)mu_p, (double)mu_t, (double)mu_raw, qm) : (double)mu_td;
        if (qm) mu_td = (float)mu_td64;
        setup_parcel_pcol(rd, L, tb, o, (double)mu_p, (double)mu_t, mu_td64, kLn2 * f_lg2(mu_p), k_mu + 1, mu, qm, PTab::kTable);
        res[2].par_p = mu_p; res[2].par_t = mu_t; res[2].par_td = mu_td; res[2].shift = k_mu;
    }
    // ---- the sweep ----------------------------------------------------------------------------------------
    if constexpr (PTab::kTable)
        return ptab7_sweep<KINDS>(rd, L, o, ptab, sb, ml, mu, res, redo, nanacc, bad_axis, best, second, k_mu, qm,
                                  p_sfc, t_sfc, td_sfc, x_sfc);
    const bool vtc = (MODE == 1) ? true : (o.vtc != 0);
    const int compat = (MODE == 1) ? 141 : o.compat;
    if (Prof::kEnabled) {       // start rows: the parcel level itself (environment == parcel there)
        auto row0 = [&](const PColParcel &c, int q, const FResult &r) {
            const float tv0 = f_tv(r.par_t, c.w_par);
            prof.put(q, 0, r.par_p, r.par_t, tv0, r.par_t, tv0, r.par_td);
        };
        if (KINDS & 1u) row0(sb, 0, res[0]);
        if (KINDS & 2u) row0(ml, 1, res[1]);
        if (KINDS & 4u) row0(mu, 2, res[2]);
    }
    // Profile rows of ONE lifted kind (the reference's mixed_layer_cape_cin / most_unstable_cape_cin return the
    // profile from the parcel's own start): row r of a column is written at iteration kfirst - 1 + r, and kfirst
    // differs from lane to lane -- at any one time the 32 lanes of a warp would write 32 DIFFERENT rows, 4 bytes
    // per 128-byte line, which the L2 turns into read-modify-write traffic (measured: 5.5 x the algorithmic reads).
    // So the sweep of such a call is re-based per lane: iteration `it` visits level it + shift (shift =
    // kfirst - 1), every lane writes row `it` at iteration `it` -- coalesced -- and the level reads are the ones
    // that diverge (they only cost cache lines that the neighbouring lanes use a few iterations later).
    int shift = 0;
    if (Prof::kEnabled && (KINDS == 2u || KINDS == 4u)) {
        PColParcel &c = (KINDS == 2u) ? ml : mu;
        shift = min(max(c.kfirst - 1, 0), L - 1);
        c.ka -= shift; c.kfirst -= shift;
    }
    const int Lq = L - shift;                         // levels from the first swept one's predecessor up
    float p_s = p_sfc, t_s = t_sfc, td_s = td_sfc;
    if (shift > 0) {
        p_s = rd.P(shift); t_s = rd.T(shift); td_s = rd.Td(shift);
        if (qm) td_s = f_td_from_q(p_s, t_s, td_s, qm);
    }
    EnvLevel e_prv = {p_s, t_s, td_s, 0.0f};
    float b_prv = 0.0f, x_prv = (shift > 0) ? kLn2 * f_lg2(p_s) : x_sfc, p_prv = p_s;
    float w_prv;
    {   // node/weight of the pressure below the first swept level and the first gathers
        const float s0 = (p_s - 2.5f) * 2.0f;
        const int j0 = min(max((int)s0, 0), kNP - 2);
        w_prv = s0 - (float)j0;
        if (!PTab::kTable) {
            if (KINDS & 1u) { sb.f0 = XP_LDG(sb.curve + j0); sb.f1 = XP_LDG(sb.curve + j0 + 1); }
            if (KINDS & 2u) { ml.f0 = XP_LDG(ml.curve + j0); ml.f1 = XP_LDG(ml.curve + j0 + 1); }
            if (KINDS & 4u) { mu.f0 = XP_LDG(mu.curve + j0); mu.f1 = XP_LDG(mu.curve + j0 + 1); }
        }
    }
    PTabLevel lv_prv = ptab.level(f_ex2((float)kKappa * f_lg2(p_s)));
    const int k1 = min(1 + shift, L - 1);
    const float *ppp = rd.pptr(k1), *tp = rd.tptr(k1), *tdp = rd.tdptr(k1);
    const int64_t ls = rd.stride(), pls = rd.pstride();
    // re-based sweep through the ring (see NoRing): warp-uniform window of levels [it + wmin, it + wmax + lead]
    bool use_ring = false;
    int wmax = 0, g_next = 0;
    if (Ring::kEnabled && Prof::kEnabled && (KINDS == 2u || KINDS == 4u)) {
        wmax = XP_WARP_MAX_INT(shift);
        const int wmin = max(0, -(XP_WARP_MAX_INT(-shift)));
        use_ring = (wmax - wmin + 1 + kRingLead) <= ring.capacity();
        g_next = wmin + 1;
    }
    float p_nxt = 0.0f, t_nxt = 0.0f, td_nxt = 0.0f;
    if (!use_ring) { p_nxt = Rd::ld(ppp); t_nxt = Rd::ld(tp); td_nxt = Rd::ld(tdp); }
    for (int it = 1; it <= Lq; ++it) {
        const bool last = (it == Lq);
        float p_cur0 = p_nxt, t = t_nxt, td = td_nxt;
        if (use_ring) {
            const int need = min(it + wmax + kRingLead, L - 1);       // the same for every lane still in the loop
            for (; g_next <= need; ++g_next) ring.put(g_next % ring.capacity(), rd.P(g_next), rd.T(g_next), rd.Td(g_next));
            if (!last) ring.get((it + shift) % ring.capacity(), p_cur0, t, td);
        }
        if (qm && !last) td = f_td_from_q(p_cur0, t, td, qm);
        ppp += pls; tp += ls; tdp += ls;
        if (!use_ring && it + 1 < Lq) { p_nxt = Rd::ld(ppp); t_nxt = Rd::ld(tp); td_nxt = Rd::ld(tdp); }
        float b_cur = 0.0f, x_cur = x_prv, pk_cur = 0.0f, p_cur = p_prv, w_cur = w_prv;
        PTabLevel lv_cur = lv_prv;
        EnvLevel e_cur = e_prv;
        int j_cur = 0;
        if (!last) {
            nanacc = f_fma(p_cur0, 0.0f, f_fma(t, 0.0f, f_fma(td, 0.0f, nanacc)));
            p_cur = p_cur0;
            if (!(p_cur < p_prv) || !(p_cur >= 2.5f)) bad_axis = true;
            // table node and weight of this level's pressure (shared by the parcels; used by the
            // lagging rows of the next iteration), PF:585-592
            const float s_cur = (p_cur - 2.5f) * 2.0f;
            j_cur = min(max((int)s_cur, 0), kNP - 2);
            w_cur = s_cur - (float)j_cur;
            const float l2p = f_lg2(p_cur);
            x_cur = kLn2 * l2p; pk_cur = f_ex2((float)kKappa * l2p);
            if (PTab::kTable) lv_cur = ptab.level(pk_cur);
            if (vtc || Prof::kEnabled) {
                const float es_t = f_es(t), es_td = f_es(td);
                e_cur.tv = f_tv(t, f_mixing_ratio(es_t, es_td, p_cur, compat));  // PF:839-843
            }
            b_cur = vtc ? e_cur.tv : t;
            e_cur.p = p_cur; e_cur.t = t; e_cur.td = td;
        }
        if (KINDS & 1u) parcel_iteration_pcol<MODE>(sb, it, last, j_cur, w_prv, pk_cur, p_prv, x_cur, x_prv, b_cur, b_prv, vtc, e_cur, e_prv, 0, prof, ptab, lv_prv);
        if (KINDS & 2u) parcel_iteration_pcol<MODE>(ml, it, last, j_cur, w_prv, pk_cur, p_prv, x_cur, x_prv, b_cur, b_prv, vtc, e_cur, e_prv, 1, prof, ptab, lv_prv);
        if (KINDS & 4u) parcel_iteration_pcol<MODE>(mu, it, last, j_cur, w_prv, pk_cur, p_prv, x_cur, x_prv, b_cur, b_prv, vtc, e_cur, e_prv, 2, prof, ptab, lv_prv);
        b_prv = b_cur; x_prv = x_cur; p_prv = p_cur; w_prv = w_cur; e_prv = e_cur; lv_prv = lv_cur;
    }
    if (Prof::kEnabled) {       // rows above the lifted column are NaN (PF:1552, 1637: levels dropped below)
        const float qn = f_qnan();
        auto pad = [&](const PColParcel &c, int q) {
            for (int row = Lq - c.kfirst + 2; row <= L; ++row) prof.put(q, row, qn, qn, qn, qn, qn, qn);
        };
        if (KINDS & 1u) pad(sb, 0);
        if (KINDS & 2u) pad(ml, 1);
        if (KINDS & 4u) pad(mu, 2);
    }
    // ---- results ----------------------------------------------------------------------------------------------
    const bool nan_seen = !(nanacc == 0.0f) || bad_axis;
    auto wrap = [&](const PColParcel &c, FResult &r, unsigned bit) {
        sweep_finish<MODE>(c, o, r);
        const bool unc = !(c.min_abs_d >= kDecisionEps) || !(c.min_slope >= 0.0f);
        if (c.bad || unc || nan_seen) redo |= bit;
        if (c.bad || nan_seen) rows_exact |= bit;
    };
    if (KINDS & 1u) wrap(sb, res[0], 1u);
    if (KINDS & 2u) wrap(ml, res[1], 2u);
    if (KINDS & 4u) wrap(mu, res[2], 4u);
    if (!Prof::kEnabled && (KINDS & 5u) == 5u && (redo & 4u) && k_mu == 0 && !nan_seen &&
        (best - second >= kThetaEMargin))
        redo = (redo & ~4u) | 1u | kRedoMuIsSb;
    if (Prof::kEnabled) redo |= ((redo & 7u) & ~rows_exact) * kRedoRowsOk;
    return redo;
}

}  // namespace fast
}  // namespace xp
