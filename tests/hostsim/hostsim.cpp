// hostsim.cpp -- TEST-ONLY host build of the per-column kernel code.
//
// Compiles xarray_parcel_b200/csrc/xp_parcels.cuh (the exact functions the sm_100a kernels
// inline) with g++ so that tests/test_hostsim_vs_oracle.py can check the kernel *logic*
// against the oracle on a machine without a GPU.  This library is built into
// tests/hostsim/_build, is never imported by the product package and is not a CPU fallback:
// the product path (libxparcel.so) has no host implementation.
#include <cstdint>

#include <vector>

#include "../../xarray_parcel_b200/csrc/xp_fast_pcol.cuh"
#include "../../xarray_parcel_b200/csrc/xp_fast6.cuh"
#include "../../xarray_parcel_b200/csrc/xp_fast7.cuh"
#include "../../xarray_parcel_b200/csrc/xp_fast_pcol6.cuh"
#include "../../xarray_parcel_b200/csrc/xp_fast_pcol7.cuh"
#include "../../xarray_parcel_b200/csrc/xp_layers.cuh"
#include "../../xarray_parcel_b200/csrc/xp_levels.cuh"

namespace {

// != 0 (141 / 162): the dewpoint arrays handed to the entry points hold specific humidity (hostsim_set_qmode)
static int g_qmode = 0;

struct HostReader {
    const double *p, *t, *td;
    int64_t ls, pls;
    int L;
    double P(int k) const { return p[(int64_t)k * pls]; }
    double Tk(int k) const { return t[(int64_t)k * ls]; }
    double Td(int k) const {
        const double raw = td[(int64_t)k * ls];
        return g_qmode ? xp::dewpoint_from_q(P(k), Tk(k), raw, g_qmode) : raw;
    }
};

struct HostProf {
    double *base;      // [6][L+1][n] or null
    int64_t n, col;
    int L;
    void put(int v, const xp::ProfileRow &r) const {
        if (!base) return;
        const double vals[6] = {r.p, r.t, r.tv, r.env_t, r.env_tv, r.env_td};
        for (int f = 0; f < 6; ++f) base[((int64_t)f * (L + 1) + v) * n + col] = vals[f];
    }
};

}  // namespace

extern "C" void hostsim_cape_cin(const double *p, const double *t, const double *td, int64_t n,
                                 int L, int p1d, int kind, const double *ex, const int *iopts,
                                 double ml_depth, double mu_depth, const uint16_t *index_grid,
                                 const float *curves, double *out /*[12][n]*/, int32_t *shift,
                                 double *prof, uint32_t *flags) {
    xp::Tables tb = {index_grid, curves};
    xp::Opts o;
    o.vtc = iopts[0]; o.log_interp = iopts[1]; o.pos_neg = iopts[2]; o.post_zero = iopts[3];
    o.compat = iopts[4]; o.exact_only = 0; o.vote_mask = 0; o.ml_depth = ml_depth; o.mu_depth = mu_depth;
    uint32_t fl = 0;
    for (int64_t c = 0; c < n; ++c) {
        HostReader rd = {p1d ? p : p + c, t + c, td + c, n, p1d ? 1 : n, L};
        HostProf pw = {prof, n, c, L};
        xp::ParcelResult r;
        double p0, t0, td0;
        int sh;
        double ep = ex ? ex[c] : 0, et = ex ? ex[n + c] : 0, etd = ex ? ex[2 * n + c] : 0;
        xp::run_column(rd, kind, tb, o, ep, et, etd, r, p0, t0, td0, sh, pw);
        fl |= r.flags;
        const double vals[12] = {r.cape, r.cin, r.lcl_p, r.lcl_t, r.lcl_tv, r.lfc_p,
                                 r.lfc_t, r.el_p, r.el_t, p0, t0, td0};
        for (int f = 0; f < 12; ++f) out[(int64_t)f * n + c] = vals[f];
        if (shift) shift[c] = sh;
    }
    if (flags) *flags = fl;
}

// The float32 fast path of the suite (xp_fast.cuh) on a shared pressure axis.
// out: [3 kinds][12 fields][n] float32, shift [3][n], redo [n] (mask of kinds the fast path hands to
// the exact kernel).  Returns Prep.ok.
namespace {
struct HostRdF {
    const float *t, *td;
    int64_t ls;
    float T(int k) const { return t[(int64_t)k * ls]; }
    float Td(int k) const { return td[(int64_t)k * ls]; }
    const float *tptr(int k) const { return t + (int64_t)k * ls; }
    const float *tdptr(int k) const { return td + (int64_t)k * ls; }
    int64_t stride() const { return ls; }
    static float ld(const float *p) { return *p; }
    static void prefetch(const float *) {}
};
struct HostRd6 {                       // v6 sweep: 32-bit element offsets from the array bases
    static constexpr bool kDouble = false;
    double ldT64(uint32_t off) const { return (double)t[off]; }
    double ldTd64(uint32_t off) const { return (double)td[off]; }
    const float *t, *td;
    uint32_t col, lstride;
    uint32_t off0() const { return col; }
    uint32_t ls() const { return lstride; }
    float ldT(uint32_t off) const { return t[off]; }
    float ldTd(uint32_t off) const { return td[off]; }
    void prefetch(uint32_t) const {}
};
struct HostEnv {
    static constexpr bool kStaged = true, kFullPass = true;
    float v[xp::fast::kMaxLevels];
    void put(int k, float x) { v[k] = x; }
    float get(int k) const { return v[k]; }
};
struct HostStash {
    float t[xp::fast::kMaxLevels], td[xp::fast::kMaxLevels];
    int capacity() const { return xp::fast::kMaxLevels; }
    void put(int k, float a, float b) { t[k] = a; td[k] = b; }
    void get(int k, float &a, float &b) const { a = t[k]; b = td[k]; }
};
struct HostCoefRow {
    const xp::fast::Coef *row;
    void advance() { row += xp::fast::kNI; }
    xp::fast::Coef at(int m) const { return row[m]; }
};
struct HostCoef {
    const xp::fast::Coef *base;
    HostCoefRow row(int k) const { return HostCoefRow{base + k * xp::fast::kNI}; }
};
}  // namespace

// Which sweep the columns 2, 3 mod 4 of hostsim_fast_suite run under the default options (6: xp_fast6.cuh, 7:
// xp_fast7.cuh) and, for v7, the stand-in for the other lanes' LCL rows (see host_warp_max_floor).
extern "C" void hostsim_set_qmode(int qmode) { g_qmode = qmode; }
static int g_fast_sweep = 7;
extern "C" void hostsim_set_fast_sweep(int version, int ka_floor) {
    g_fast_sweep = version;
    xp::fast::host_warp_max_floor() = ka_floor;
}

extern "C" int hostsim_fast_suite(const float *p, const float *t, const float *td, int64_t n, int L,
                                  const int *iopts, double ml_depth, double mu_depth,
                                  const uint16_t *index_grid, const float *curves, float *out,
                                  int32_t *shift, uint32_t *redo) {
    xp::Tables tb = {index_grid, curves};
    xp::Opts o;
    o.vtc = iopts[0]; o.log_interp = iopts[1]; o.pos_neg = iopts[2]; o.post_zero = iopts[3];
    o.compat = iopts[4]; o.exact_only = 0; o.vote_mask = 0; o.ml_depth = ml_depth; o.mu_depth = mu_depth;
    o.qmode = g_qmode;
    static xp::fast::Prep pr;
    xp::fast::compute_prep(p, 1, L, o, pr);
    if (!pr.ok) return 0;
    std::vector<xp::fast::Coef> coef((size_t)L * xp::fast::kNI);
    for (int k = 0; k < pr.n_table; ++k)
        for (int m = 0; m < xp::fast::kNI; ++m) coef[(size_t)k * xp::fast::kNI + m] = xp::fast::compute_coef(pr, curves, k, m);
    HostCoef cf = {coef.data()};
    // default options: the v6 sweep on the virtual-temperature table (xp_fast6.cuh) -- what the kernel runs
    const bool m1 = o.vtc && o.compat == 141 && o.pos_neg;
    if (g_qmode && !m1) return 0;          // specific-humidity input: the v7 sweep only (fast_eligible in xp_fast.cu)
    std::vector<xp::fast::Coef> coef_tv;
    if (m1) {
        coef_tv.resize((size_t)L * xp::fast::kNI);
        for (int k = 0; k < pr.n_table; ++k)
            for (int m = 0; m < xp::fast::kNI; ++m)
                coef_tv[(size_t)k * xp::fast::kNI + m] = xp::fast::compute_coef_tv(pr, curves, k, m);
    }
    HostCoef cf_tv = {coef_tv.data()};
    for (int64_t c = 0; c < n; ++c) {
        HostRdF rd = {t + c, td + c, n};
        xp::fast::FResult r[3];
        // default options: v6 sweep on columns 2, 3 mod 4 (the generic sweep keeps 0, 1 mod 4);
        // other options: even columns environment staged + early termination, odd columns recomputed
        if (m1 && ((c & 3) >= 2 || g_qmode)) {
            // column 2 mod 4: T/Td of the lowest levels stashed by the pre-pass; 3 mod 4: no stash
            const HostRd6 rd6 = {t, td, (uint32_t)c, (uint32_t)n};
            if ((c & 3) == 2) {
                HostStash st;
                redo[c] = (g_fast_sweep == 7) ? xp::fast::suite_column7<7u, true>(rd6, cf_tv, pr, tb, o, st, r)
                                              : xp::fast::suite_column6<7u>(rd6, cf_tv, pr, tb, o, st, r);
            } else {
                xp::fast::NoStash st;
                redo[c] = (g_fast_sweep == 7) ? xp::fast::suite_column7<7u, true>(rd6, cf_tv, pr, tb, o, st, r)
                                              : xp::fast::suite_column6<7u>(rd6, cf_tv, pr, tb, o, st, r);
            }
        } else if (c & 1) {
            xp::fast::EnvRecompute env;
            redo[c] = (o.vtc && o.compat == 141 && o.pos_neg) ? xp::fast::suite_column<7u, 1>(rd, cf, pr, tb, o, env, r)
                                                              : xp::fast::suite_column<7u, 0>(rd, cf, pr, tb, o, env, r);
        } else {
            HostEnv env;
            redo[c] = (o.vtc && o.compat == 141 && o.pos_neg) ? xp::fast::suite_column<7u, 1>(rd, cf, pr, tb, o, env, r)
                                                              : xp::fast::suite_column<7u, 0>(rd, cf, pr, tb, o, env, r);
        }
        for (int q = 0; q < 3; ++q) {
            const float vals[12] = {r[q].cape, r[q].cin, r[q].lcl_p, r[q].lcl_t, r[q].lcl_tv, r[q].lfc_p,
                                    r[q].lfc_t, r[q].el_p, r[q].el_t, r[q].par_p, r[q].par_t, r[q].par_td};
            for (int f = 0; f < 12; ++f) out[((int64_t)q * 12 + f) * n + c] = vals[f];
            shift[(int64_t)q * n + c] = r[q].shift;
        }
    }
    return 1;
}

// The float32 fast path for per-column pressure (xp_fast_pcol.cuh).  Same outputs as hostsim_fast_suite.
namespace {
struct HostRdP {
    const float *p, *t, *td;
    int64_t ls;
    float P(int k) const { return p[(int64_t)k * ls]; }
    float T(int k) const { return t[(int64_t)k * ls]; }
    float Td(int k) const { return td[(int64_t)k * ls]; }
    const float *pptr(int k) const { return p + (int64_t)k * ls; }
    const float *tptr(int k) const { return t + (int64_t)k * ls; }
    const float *tdptr(int k) const { return td + (int64_t)k * ls; }
    int64_t stride() const { return ls; }
    int64_t pstride() const { return ls; }
    static float ld(const float *q) { return *q; }
};
}  // namespace

namespace {
struct HostRdP6 {                      // pcol v6 sweep: 32-bit element offsets from the array bases
    const float *p, *t, *td;
    uint32_t col, lstride;
    uint32_t off0() const { return col; }
    uint32_t ls() const { return lstride; }
    uint32_t pls() const { return lstride; }
    float ldP(uint32_t off) const { return p[off]; }
    float ldT(uint32_t off) const { return t[off]; }
    float ldTd(uint32_t off) const { return td[off]; }
    void prefetch(uint32_t, uint32_t) const {}
};
struct HostStash3 {
    float p[128], t[128], td[128];
    int cap;
    int capacity() const { return cap; }
    void put(int k, float a, float b, float c) { p[k] = a; t[k] = b; td[k] = c; }
    void get(int k, float &a, float &b, float &c) const { a = p[k]; b = t[k]; c = td[k]; }
};
struct HostProfW {
    static constexpr bool kEnabled = true;
    float *base;       // [3 kinds][6 fields][L+1][n]
    int64_t n, col;
    int L;
    void put(int q, int row, float p, float tp, float tv, float et, float etv, float etd) const {
        const float v[6] = {p, tp, tv, et, etv, etd};
        if (row < 0 || row > L) __builtin_trap();          // the device writer drops such rows; here they are bugs
        for (int f = 0; f < 6; ++f) base[(((int64_t)q * 6 + f) * (L + 1) + row) * n + col] = v[f];
    }
};
}  // namespace

// 0: all three kinds at once (KINDS = 7); 2 / 4: ONE lifted kind (mixed layer / most unstable) with profile rows -- the
// re-based sweep -- read through a per-thread ring of `g_pcol_ring` levels (0: direct loads).
static int g_pcol_kind = 0, g_pcol_ring = 0, g_pcol_table = 0;
extern "C" void hostsim_set_pcol_kind(int kind, int ring_levels) { g_pcol_kind = kind; g_pcol_ring = ring_levels; }
// 1: default options without profile rows read the adiabat family from the two-segment table (fast::PTabView) -- what
// suite_fast_ptab_kernel runs
extern "C" void hostsim_set_pcol_table(int on) { g_pcol_table = on; }
// max |table - reference evaluation| over a sample of (adiabat, pressure) pairs: out[0] segment A (100..1100 hPa),
// out[1] 20..100 hPa, out[2] 2.5..20 hPa
extern "C" void hostsim_ptab_error(const float *curves, int n_samples, double *out) {
    std::vector<xp::fast::Coef> tab((size_t)xp::fast::kPTabNodes * xp::fast::kPTabIntervals);
    for (int j = 0; j < xp::fast::kPTabNodes; ++j)
        for (int m = 0; m < xp::fast::kPTabIntervals; ++m)
            tab[(size_t)j * xp::fast::kPTabIntervals + m] = xp::fast::compute_ptab_coef(curves, j, m + xp::fast::kPTabFirstInterval);
    const xp::fast::PTabView pt{tab.data(), xp::fast::ptab_desc()};
    for (int i = 0; i < 9; ++i) out[i] = 0.0;
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; };
    for (int i = 0; i < n_samples; ++i) {
        const int a0 = xp::fast::kPTabFirstInterval * xp::fast::kNodeStride +
                       (int)(rnd() * ((xp::fast::kLastInterval + 1 - xp::fast::kPTabFirstInterval) * xp::fast::kNodeStride - 1));
        const int seg = i % 3;
        const double p = seg == 0 ? 100.0 + rnd() * 1000.0 : (seg == 1 ? 20.0 + rnd() * 80.0 : 2.5 + rnd() * 17.5);
        const int m = a0 / xp::fast::kNodeStride;
        const float f = (float)(a0 - m * xp::fast::kNodeStride) * (1.0f / xp::fast::kNodeStride);
        const float xi = xp::fast::f_ex2((float)xp::kKappa * xp::fast::f_lg2((float)p));
        const float got = pt.eval(pt.level(xi), m, f);
        const double t = xp::adiabat_temperature(curves + (size_t)a0 * xp::kNP, (double)(float)p);
        const double want = xp::virtual_temperature(t, xp::sat_mixing_ratio((double)(float)p, t));
        const double e = fabs((double)got - want);
        if (e > out[seg]) { out[seg] = e; out[3 + 2 * seg] = (double)a0; out[4 + 2 * seg] = p; }
    }
}
namespace {
struct HostRing {
    static constexpr bool kEnabled = true;
    int cap;
    float p[64], t[64], td[64];
    int capacity() const { return cap; }
    void put(int s, float a, float b, float c) { if (s < 0 || s >= cap) __builtin_trap(); p[s] = a; t[s] = b; td[s] = c; }
    void get(int s, float &a, float &b, float &c) const { if (s < 0 || s >= cap) __builtin_trap(); a = p[s]; b = t[s]; c = td[s]; }
};
}  // namespace

extern "C" int hostsim_fast_suite_pcol(const float *p, const float *t, const float *td, int64_t n, int L,
                                       const int *iopts, double ml_depth, double mu_depth,
                                       const uint16_t *index_grid, const float *curves, float *out,
                                       int32_t *shift, uint32_t *redo, float *prof) {
    xp::Tables tb = {index_grid, curves};
    xp::Opts o;
    o.vtc = iopts[0]; o.log_interp = iopts[1]; o.pos_neg = iopts[2]; o.post_zero = iopts[3];
    o.compat = iopts[4]; o.exact_only = 0; o.vote_mask = 0; o.ml_depth = ml_depth; o.mu_depth = mu_depth;
    o.qmode = g_qmode;
    for (int64_t c = 0; c < n; ++c) {
        HostRdP rd = {p + c, t + c, td + c, n};
        xp::fast::FResult r[3];
        const bool m1 = o.vtc && o.compat == 141 && o.pos_neg;
        xp::fast::NoRing nr;
        if (prof && (g_pcol_kind == 2 || g_pcol_kind == 4)) {
            HostProfW pw = {prof, n, c, L};
            HostRing ring; ring.cap = g_pcol_ring;
            for (int q = 0; q < 3; ++q) r[q] = xp::fast::FResult();
            if (g_pcol_kind == 2)
                redo[c] = g_pcol_ring ? xp::fast::suite_column_pcol<2u, 1, true>(rd, L, tb, o, pw, ring, xp::fast::NoPTab(), r)
                                      : xp::fast::suite_column_pcol<2u, 1, true>(rd, L, tb, o, pw, nr, xp::fast::NoPTab(), r);
            else
                redo[c] = g_pcol_ring ? xp::fast::suite_column_pcol<4u, 1, true>(rd, L, tb, o, pw, ring, xp::fast::NoPTab(), r)
                                      : xp::fast::suite_column_pcol<4u, 1, true>(rd, L, tb, o, pw, nr, xp::fast::NoPTab(), r);
        } else if (prof) {
            HostProfW pw = {prof, n, c, L};
            redo[c] = m1 ? xp::fast::suite_column_pcol<7u, 1, true>(rd, L, tb, o, pw, nr, xp::fast::NoPTab(), r)
                         : xp::fast::suite_column_pcol<7u, 0, true>(rd, L, tb, o, pw, nr, xp::fast::NoPTab(), r);
        } else if (m1 && (c % 3) != 0 && !g_qmode && !g_pcol_table) {
            // default options, no profile: the v6 sweep (xp_fast_pcol6.cuh) on two columns out of three, with a
            // stash of 36 levels (the kernel's), of 5 levels (search and sweep cross its end) or none
            const HostRdP6 rd6 = {p, t, td, (uint32_t)c, (uint32_t)n};
            if ((c % 3) == 1) {
                HostStash3 st; st.cap = (c % 2) ? 36 : 5;
                redo[c] = xp::fast::suite_column_pcol6<7u>(rd6, L, tb, o, st, r);
            } else {
                xp::fast::NoStash3 st;
                redo[c] = xp::fast::suite_column_pcol6<7u>(rd6, L, tb, o, st, r);
            }
        } else if (m1 && g_pcol_table) {
            static std::vector<xp::fast::Coef> tab;
            static const float *tab_of = nullptr;
            if (tab_of != curves) {
                tab.resize((size_t)xp::fast::kPTabNodes * xp::fast::kPTabIntervals);
                for (int j = 0; j < xp::fast::kPTabNodes; ++j)
                    for (int m = 0; m < xp::fast::kPTabIntervals; ++m)
                        tab[(size_t)j * xp::fast::kPTabIntervals + m] =
                            xp::fast::compute_ptab_coef(curves, j, m + xp::fast::kPTabFirstInterval);
                tab_of = curves;
            }
            const xp::fast::PTabView pt{tab.data(), xp::fast::ptab_desc()};
            xp::fast::NoProfile np;
            redo[c] = xp::fast::suite_column_pcol<7u, 1, true>(rd, L, tb, o, np, nr, pt, r);
        } else {
            xp::fast::NoProfile np;
            redo[c] = m1 ? xp::fast::suite_column_pcol<7u, 1, true>(rd, L, tb, o, np, nr, xp::fast::NoPTab(), r)
                         : xp::fast::suite_column_pcol<7u, 0, true>(rd, L, tb, o, np, nr, xp::fast::NoPTab(), r);
        }
        for (int q = 0; q < 3; ++q) {
            const float vals[12] = {r[q].cape, r[q].cin, r[q].lcl_p, r[q].lcl_t, r[q].lcl_tv, r[q].lfc_p,
                                    r[q].lfc_t, r[q].el_p, r[q].el_t, r[q].par_p, r[q].par_t, r[q].par_td};
            for (int f = 0; f < 12; ++f) out[((int64_t)q * 12 + f) * n + c] = vals[f];
            shift[(int64_t)q * n + c] = r[q].shift;
        }
    }
    return 1;
}

// LCL solvers of the fast paths on n parcels: which = 0 lcl_fast (xp_fast.cuh), 1 lcl_fast6 (xp_fast6.cuh).
extern "C" void hostsim_lcl_fast(const double *p, const double *t, const double *td, int64_t n, int which,
                                 double *lcl_p, double *lcl_t) {
    for (int64_t i = 0; i < n; ++i) {
        if (which == 0) xp::fast::lcl_fast(p[i], t[i], td[i], lcl_p[i], lcl_t[i]);
        else xp::fast::lcl_fast6(p[i], t[i], td[i], lcl_p[i], lcl_t[i]);
    }
}

// Branch-free float64 log / exp of the v6 fast path (xp_fast6.cuh): which = 0 log, 1 exp.
extern "C" void hostsim_fast_math64(const double *x, int64_t n, int which, double *y) {
    for (int64_t i = 0; i < n; ++i)
        y[i] = which == 0 ? xp::fast::log64_fast(x[i]) : (which == 1 ? xp::fast::exp64_fast(x[i]) : xp::pow_kappa64(x[i]));
}

// Layer primitives (xp_layers.cuh): mixed_layer of n_fields variables x [n_fields][L][n], mixed_parcel (out
// [6][n]) and get_layer's bounds, column by column.
extern "C" void hostsim_mixed_layer(const double *p, int p1d, const double *x, int n_fields, int64_t n, int L,
                                    double depth, int pressure_field, double *out /*[n_fields][n]*/) {
    for (int64_t c = 0; c < n; ++c) {
        const double *pc = p1d ? p : p + c;
        const int64_t pls = p1d ? 1 : n;
        auto pressure_at = [&](int k) { return pc[(int64_t)k * pls]; };
        for (int f0 = 0; f0 < n_fields; f0 += 4) {
            auto load = [&](int k, double (&v)[4]) {
                for (int f = 0; f < 4; ++f)
                    v[f] = f0 + f < n_fields ? x[((int64_t)(f0 + f) * L + k) * n + c] : 0.0;
            };
            double bottom, top, mean[4];
            xp::mixed_layer_means<4>(L, pressure_at, load, depth, bottom, top, mean,
                                     pressure_field >= f0 && pressure_field < f0 + 4 ? pressure_field - f0 : -1);
            for (int f = 0; f < 4 && f0 + f < n_fields; ++f) out[(int64_t)(f0 + f) * n + c] = mean[f];
        }
    }
}

extern "C" void hostsim_mixed_parcel(const double *p, int p1d, const double *t, const double *td, int64_t n, int L,
                                     double depth, double *out /*[6][n]*/) {
    for (int64_t c = 0; c < n; ++c) {
        const double *pc = p1d ? p : p + c;
        const int64_t pls = p1d ? 1 : n;
        auto pressure_at = [&](int k) { return pc[(int64_t)k * pls]; };
        auto t_at = [&](int k) { return t[(int64_t)k * n + c]; };
        auto td_at = [&](int k) { return td[(int64_t)k * n + c]; };
        double o[6];
        xp::mixed_parcel_full(L, pressure_at, t_at, td_at, depth, o);
        for (int f = 0; f < 6; ++f) out[(int64_t)f * n + c] = o[f];
    }
}

extern "C" void hostsim_layer_bounds(const double *p, int p1d, int64_t n, int L, double depth, int interpolate,
                                     double *bottom, double *top) {
    for (int64_t c = 0; c < n; ++c) {
        const double *pc = p1d ? p : p + c;
        const int64_t pls = p1d ? 1 : n;
        auto pressure_at = [&](int k) { return pc[(int64_t)k * pls]; };
        xp::layer_bounds(L, pressure_at, depth, interpolate != 0, bottom[c], top[c]);
    }
}

// Level primitives (xp_levels.cuh), column by column; one variable v [L][n] per call.
extern "C" void hostsim_insert_level(const double *c, const double *v, const double *lev_c, const double *lev_v,
                                     int64_t n, int L, double *out /*[L+1][n]*/) {
    for (int64_t i = 0; i < n; ++i)
        for (int j = 0; j <= L; ++j) {
            const double c_j = j < L ? c[(int64_t)j * n + i] : xp::qnan(), c_jm = j >= 1 ? c[(int64_t)(j - 1) * n + i] : xp::qnan();
            const double v_j = j < L ? v[(int64_t)j * n + i] : xp::qnan(), v_jm = j >= 1 ? v[(int64_t)(j - 1) * n + i] : xp::qnan();
            out[(int64_t)j * n + i] = xp::insert_level_value(j, L, c_j, c_jm, v_j, v_jm, lev_c[i], lev_v[i]);
        }
}

extern "C" void hostsim_shift_out_nans(const double *ref, const double *v, int64_t n, int L, double *out, int32_t *shift) {
    for (int64_t i = 0; i < n; ++i) {
        auto ref_at = [&](int k) { return ref[(int64_t)k * n + i]; };
        const int k0 = xp::leading_nans(L, ref_at);
        shift[i] = k0;
        for (int j = 0; j < L; ++j) out[(int64_t)j * n + i] = j + k0 < L ? v[(int64_t)(j + k0) * n + i] : xp::qnan();
    }
}

extern "C" void hostsim_trapz(const double *x, const double *v, const uint8_t *mask, int64_t n, int L, int sign, double *out) {
    for (int64_t i = 0; i < n; ++i) {
        auto x_at = [&](int k) { return x[(int64_t)k * n + i]; };
        auto v_at = [&](int k) { return v[(int64_t)k * n + i]; };
        auto use = [&](int k) { return !mask || mask[(int64_t)k * n + i] != 0; };
        out[i] = xp::trapz_column(L, x_at, v_at, use, sign);
    }
}

extern "C" void hostsim_find_intersections(const double *x, const double *a, const double *b, int64_t n, int L, int log_x,
                                           double *out /*[3][L-1][n]: ix, iy, sign*/) {
    for (int64_t i = 0; i < n; ++i)
        for (int k = 1; k < L; ++k) {
            double x0 = x[(int64_t)(k - 1) * n + i], x1 = x[(int64_t)k * n + i];
            if (log_x) { x0 = log(x0); x1 = log(x1); }
            double ix, iy, sc;
            xp::interval_crossing(x0, x1, a[(int64_t)(k - 1) * n + i], a[(int64_t)k * n + i], b[(int64_t)(k - 1) * n + i],
                                  b[(int64_t)k * n + i], log_x != 0, ix, iy, sc);
            out[((int64_t)0 * (L - 1) + k - 1) * n + i] = ix;
            out[((int64_t)1 * (L - 1) + k - 1) * n + i] = iy;
            out[((int64_t)2 * (L - 1) + k - 1) * n + i] = sc;
        }
}

extern "C" void hostsim_interp1d(const double *at, const double *xp, const double *fp, int64_t m, int n, double *out) {
    auto xp_at = [&](int k) { return xp[k]; };
    auto fp_at = [&](int k) { return fp[k]; };
    for (int64_t i = 0; i < m; ++i) out[i] = xp::interp1d_point(at[i], n, xp_at, fp_at);
}

extern "C" void hostsim_trap_around_zeros(const double *x, const double *y, int64_t n, int L, int log_x,
                                          double *out /*[3][2L-1][n]: area, x, dx*/) {
    const int R = 2 * L - 1;
    for (int64_t i = 0; i < n; ++i) {
        for (int q = 0; q < 3; ++q) out[((int64_t)q * R + L - 1) * n + i] = xp::qnan();
        for (int k = 0; k + 1 < L; ++k) {
            double before[3], after[3];
            xp::zero_half_areas(x[(int64_t)k * n + i], x[(int64_t)(k + 1) * n + i], y[(int64_t)k * n + i],
                                y[(int64_t)(k + 1) * n + i], log_x != 0, before, after);
            for (int q = 0; q < 3; ++q) {
                out[((int64_t)q * R + k) * n + i] = before[q];
                out[((int64_t)q * R + L + k) * n + i] = after[q];
            }
        }
    }
}

extern "C" int hostsim_pressure_order(const double *p, int64_t n, int L) {
    int r = 0;
    for (int64_t i = 0; i < n; ++i) {
        auto pressure_at = [&](int k) { return p[(int64_t)k * n + i]; };
        r |= xp::pressure_order(L, pressure_at);
    }
    return r;
}
