// xp_math.cuh -- thermodynamic device functions of the parcel path.
//
// MetPy (<= 1.6) formulas that the reference calls (parcel_functions.py "PF" call sites in
// brackets), written so the operation order matches the Python expressions; the constants
// are computed from the same literals as metpy.constants so they are bit-identical.
#pragma once
#include <math.h>
#include <stdint.h>

// XP_HD marks the per-column functions.  Under nvcc they are device functions; the same
// sources also compile with a plain host C++ compiler (XP_HOST_SIM) -- used ONLY by
// tests/hostsim to check the kernel logic against the oracle on machines without a GPU.
// The product never runs this code on the CPU: libxparcel.so contains only the device path.
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define XP_HD __device__ __forceinline__
#define XP_LDG(ptr) __ldg(ptr)
#else
#include <algorithm>
#include <cmath>
#include <cstring>
#define XP_HOST_SIM 1
#define XP_HD inline
#define XP_LDG(ptr) (*(ptr))
using std::isfinite;
using std::isnan;
using std::max;
using std::min;
#endif

namespace xp {

// ---- metpy.constants ---------------------------------------------------------------------
constexpr double kRgas = 8.314462618;
constexpr double kMd = 28.96546e-3;
constexpr double kMw = 18.015268e-3;
constexpr double kRd = kRgas / kMd;                      // 287.04749097718457
constexpr double kEps = kMw / kMd;                       // 0.6219569100577033
constexpr double kGamma = 1.4;
constexpr double kCpd = kGamma * kRd / (kGamma - 1.0);   // 1004.6662184201462
constexpr double kKappa = kRd / kCpd;                    // 0.28571428571428564
constexpr double kInvKappa = 1.0 / kKappa;
constexpr double kLv = 2.50084e6;
constexpr double kSat0 = 6.112;
constexpr double kZeroC = 273.15;
constexpr double kVtEps = 0.608;                         // PF:782 (Doswell & Rasmussen 1994)

XP_HD double qnan() {
#if defined(__CUDACC__)
    return __longlong_as_double(0x7ff8000000000000LL);
#else
    return std::nan("");
#endif
}

// ---- branch-free float64 helpers (no IEEE division, no libm slow paths) --------------------------------------------
// Used by the float32 fast paths (float64 LCL polish, mixed-layer means) and, with XP_F64_FAST_MATH, by the exact
// fix-up over the uncertain-column list (xp_list.cu), whose run time is the latency of ONE item's float64 chain.
#if defined(__CUDACC__)
XP_HD float xp_rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#else
XP_HD float xp_rcp_approx(float x) { return 1.0f / x; }
#endif
// Reciprocal and square root by float32-seeded Newton steps (~1e-16 relative), branch-free log / exp below.
XP_HD double rcp64(double y) {
    double r = (double)xp_rcp_approx((float)y);
    r = fma(fma(-y, r, 1.0), r, r);
    return fma(fma(-y, r, 1.0), r, r);
}
XP_HD double sqrt64(double y) {               // y ~ 1
#if defined(__CUDACC__)
    double r = (double)rsqrtf((float)y);
#else
    double r = 1.0 / sqrt(y);
#endif
    r = r * fma(-0.5 * y, r * r, 1.5);
    r = r * fma(-0.5 * y, r * r, 1.5);
    return y * r;
}
// Branch-free float64 log / exp for the arguments of this path (no special cases: x finite, log: x in
// (1e-300, 1e300), exp: |x| < 700).  ~3 ulp -- the path needs 1e-12 relative -- in a third of the instructions
// of the libm versions and, having no slow-path branches, they let the compiler interleave the three parcels.
// Polynomial coefficients live in constant memory on the device: a DFMA takes a constant-bank operand directly,
// whereas a float64 literal costs two uniform moves per use.
#if defined(__CUDACC__)
#define XP_CONST_TABLE static __constant__ double
#else
#define XP_CONST_TABLE static const double
#endif
XP_CONST_TABLE kLogC[11] = {1.0 / 23.0, 1.0 / 21.0, 1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0,
                            1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0};
XP_CONST_TABLE kExpC[12] = {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,
                            1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5};
XP_CONST_TABLE kLn2Split[2] = {6.93147180369123816490e-01, 1.90821492927058770002e-10};

XP_HD double log64_fast(double x) {
#if defined(__CUDACC__)
    long long bits = __double_as_longlong(x);
#else
    long long bits; std::memcpy(&bits, &x, 8);
#endif
    int e = (int)(bits >> 52) - 1023;
    long long mb = (bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL;        // mantissa in [1, 2)
    const bool hi = (bits & 0x000fffffffffffffLL) > 0x0006a09e667f3bcdLL;       // > sqrt(2): use m/2, e+1
    mb = hi ? (mb - 0x0010000000000000LL) : mb;
    e = hi ? e + 1 : e;
#if defined(__CUDACC__)
    const double m = __longlong_as_double(mb);
#else
    double m; std::memcpy(&m, &mb, 8);
#endif
    const double f = m - 1.0;                                                   // [-0.293, 0.414]
    const double s = f * rcp64(2.0 + f);                                        // |s| <= 0.172
    const double z = s * s;
    // sum_{k=0..10} z^k / (2k + 3) by Estrin's scheme: dependent depth 4 instead of Horner's 10 (three more
    // instructions; these chains sit on the critical path of every parcel set-up)
    const double z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
    const double b0 = fma(kLogC[9], z, kLogC[10]), b1 = fma(kLogC[7], z, kLogC[8]), b2 = fma(kLogC[5], z, kLogC[6]);
    const double b3 = fma(kLogC[3], z, kLogC[4]), b4 = fma(kLogC[1], z, kLogC[2]);
    const double c0 = fma(b1, z2, b0), c1 = fma(b3, z2, b2), c2 = fma(kLogC[0], z2, b4);
    const double p = fma(c2, z8, fma(c1, z4, c0));
    const double lm = fma(2.0 * s * z, p, 2.0 * s);                             // ln m = 2 atanh(s)
    const double ed = (double)e;
    return fma(ed, kLn2Split[0], fma(ed, kLn2Split[1], lm));                    // + e ln2 (hi + lo)
}
XP_HD double exp64_fast(double x) {
    const double kMagic = 6755399441055744.0;                                   // 1.5 * 2^52: rint by addition
    const double tn = fma(x, 1.4426950408889634, kMagic);
    const double n = tn - kMagic;
    double r = fma(-n, kLn2Split[0], x);
    r = fma(-n, kLn2Split[1], r);                                               // |r| <= 0.3466
    // sum_{k=0..13} r^k / k! (kExpC = 1/13! ... 1/2!) by Estrin's scheme: dependent depth 4 instead of 14
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double b0 = 1.0 + r, b1 = fma(kExpC[10], r, kExpC[11]), b2 = fma(kExpC[8], r, kExpC[9]);
    const double b3 = fma(kExpC[6], r, kExpC[7]), b4 = fma(kExpC[4], r, kExpC[5]), b5 = fma(kExpC[2], r, kExpC[3]);
    const double b6 = fma(kExpC[0], r, kExpC[1]);
    const double c0 = fma(b1, r2, b0), c1 = fma(b3, r2, b2), c2 = fma(b5, r2, b4);
    double p = fma(fma(b6, r4, c2), r8, fma(c1, r4, c0));
#if defined(__CUDACC__)
    // (the exponent is added as an UNSIGNED shift: shifting a negative signed value left is undefined behaviour in
    //  C++17 -- UBSan flags it in the host build of this code -- while the two's-complement sum is what is wanted)
    const unsigned long long ni = (unsigned long long)(long long)__double2int_rn(n);
    return __longlong_as_double((long long)((unsigned long long)__double_as_longlong(p) + (ni << 52)));
#else
    unsigned long long pb; std::memcpy(&pb, &p, 8);
    pb += ((unsigned long long)(long long)n) << 52;
    double out; std::memcpy(&out, &pb, 8);
    return out;
#endif
}



// x^(2/7) (= x^kappa: the Exner function and potential temperature factors) in float64 without log / exp: a float32
// estimate y0 (two MUFU, ~3e-7 relative) and ONE Newton step on y^7 = x^2, whose error is 3 eps^2 ~ 3e-13 relative --
// a quarter of the instructions of exp64_fast(kappa * log64_fast(x)).  x must be positive and finite.
XP_HD double pow_kappa64(double x) {
#if defined(__CUDACC__)
    float l2, y0f;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"((float)x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y0f) : "f"(l2 * (2.0f / 7.0f)));
#else
    const float y0f = exp2f(log2f((float)x) * (2.0f / 7.0f));
#endif
    const double y0 = (double)y0f;
    const double y2 = y0 * y0, y3 = y2 * y0, y6 = y3 * y3, y7 = y6 * y0;
    // y1 = y0 + y0 (x^2 / y0^7 - 1) / 7
    return fma(y0 * (1.0 / 7.0), fma(x * x, rcp64(y7), -1.0), y0);
}

// exp / log / pow of the exact column code.  Default: libm, rounding like NumPy's to the last bits.  With
// XP_F64_FAST_MATH (xp_list.cu only): the branch-free versions above (~3 ulp) for ordinary arguments, libm for the
// special ones (NaN, zero, negative, overflow), so that NaN propagation and the reference's asserts behave the same.
#if defined(XP_F64_FAST_MATH)
XP_HD double xp_exp(double x) { return (fabs(x) < 700.0) ? exp64_fast(x) : exp(x); }
XP_HD double xp_log(double x) { return (x > 1e-300 && x < 1e300) ? log64_fast(x) : log(x); }
XP_HD double xp_pow(double x, double y) {
    if (!(x > 1e-300 && x < 1e300)) return pow(x, y);
    const double l = y * log64_fast(x);
    return (fabs(l) < 700.0) ? exp64_fast(l) : pow(x, y);
}
#else
XP_HD double xp_exp(double x) { return exp(x); }
XP_HD double xp_log(double x) { return log(x); }
XP_HD double xp_pow(double x, double y) { return pow(x, y); }
#endif

// Bolton (1980): metpy.calc.saturation_vapor_pressure [PF:258, 698, 760]
XP_HD double sat_vapor_pressure(double t) {
    return kSat0 * xp_exp(17.67 * (t - 273.15) / (t - 29.65));
}
// metpy.calc.dewpoint, degC -> K [PF:280-281, inside metpy.calc.lcl PF:644]
XP_HD double dewpoint_from_e(double e) {
    double val = xp_log(e / kSat0);
    return 243.5 * val / (17.67 - val) + kZeroC;
}
// metpy.calc.mixing_ratio(e, p)
XP_HD double mixing_ratio_ep(double e, double p) { return kEps * e / (p - e); }
// metpy.calc.saturation_mixing_ratio(p, T) [PF:258, 760]
XP_HD double sat_mixing_ratio(double p, double t) {
    return mixing_ratio_ep(sat_vapor_pressure(t), p);
}
// metpy.calc.vapor_pressure(p, w) [PF:275]
XP_HD double vapor_pressure(double p, double w) { return p * w / (kEps + w); }

// metpy.calc.dewpoint_from_specific_humidity(p, T, q) as the reference calls it [PF:1889, 1969; parcel_test.py:
// 432-436]: MetPy 1.4.1 goes through the relative humidity, MetPy >= 1.6 through the vapour pressure
// (environment_changes_eval.ipynb:278).  Used by the pointwise kernel and by every column reader when the
// dewpoint array of a call holds specific humidity (xp_columns.dewpoint_is_specific_humidity).
XP_HD double dewpoint_from_q(double p, double t, double q, int compat) {
    const double w = q / (1 - q);                                        // mixing_ratio_from_specific_humidity
    double e;
    if (compat == 162) {
        e = p * w / (kEps + w);                                          // vapor_pressure(p, w)
    } else {
        const double es_t = sat_vapor_pressure(t);
        const double rh = w / (kEps * es_t / (p - es_t));                // relative_humidity_from_mixing_ratio
        e = rh * es_t;
    }
    return dewpoint_from_e(e);
}

// PF:684-710 mixing_ratio(T, Td, p): RH from dewpoint, then w from RH.
XP_HD double mixing_ratio_t_td(double t, double td, double p, int compat) {
    double es_t = sat_vapor_pressure(t);
    double es_td = sat_vapor_pressure(td);
    double rh = es_td / es_t;
    double ws = mixing_ratio_ep(es_t, p);
    if (compat == 162) return kEps * ws * rh / (kEps + ws * (1.0 - rh));
    return rh * ws;
}
// PF:782-804
XP_HD double virtual_temperature(double t, double w) {
    return t * (1 + kVtEps * w);
}
XP_HD double exner(double p) { return xp_pow(p / 1000.0, kKappa); }   // [PF:269]
XP_HD double potential_temperature(double p, double t) {           // [PF:253]
    return t / exner(p);
}
// PF:291-316
XP_HD double dry_lapse(double p, double t0, double p0) {
    return t0 * xp_pow(p / p0, kKappa);
}
// metpy.calc.equivalent_potential_temperature, Bolton (1980) eq. 39 [PF:123]
XP_HD double theta_e(double p, double t, double td) {
    double r = sat_mixing_ratio(p, td);
    double e = sat_vapor_pressure(td);
    double t_l = 56 + 1. / (1. / (td - 56) + xp_log(t / td) / 800.);
    double th_l = potential_temperature(p - e, t) * xp_pow(t / t_l, 0.28 * r);
    return th_l * xp_exp(r * (1 + 0.448 * r) * (3036. / t_l - 1.78));
}
// dT/dp of the pseudo-adiabat (metpy.calc.moist_lapse) [PF:480]
XP_HD double moist_lapse_rhs(double p, double t) {
    double rs = sat_mixing_ratio(p, t);
    double frac = (kRd * t + kLv * rs) / (kCpd + (kLv * kLv * rs * kEps / (kRd * t * t)));
    return frac / p;
}

// metpy.calc.lcl [PF:644] per parcel, iterated to ITS OWN fixed point (the reference's
// SciPy fixed_point stops on an array-wide 1e-5 criterion; this is the limit it approaches,
// see oracle/thermo.py lcl(mode='converged')).  p, t, td must be finite.
XP_HD void lcl_solve(double p0, double t, double td, double &lcl_p,
                                          double &lcl_t) {
    const double w = mixing_ratio_ep(sat_vapor_pressure(td), p0);
    auto g = [&](double p) {
        double tdp = dewpoint_from_e(vapor_pressure(p, w));
        return p0 * xp_pow(tdp / t, kInvKappa);
    };
    double p = p0;
    bool bad = false;
    // Steffensen (Aitken delta^2) steps, as scipy.optimize.fixed_point(method='del2') does,
    // then plain iterations until the iterate stops moving.
    for (int it = 0; it < 6; ++it) {
        double p1 = g(p);
        double p2 = g(p1);
        if (isnan(p1) || isnan(p2)) { bad = true; break; }
        double d = p2 - 2.0 * p1 + p;
        double pn = (d != 0.0) ? p - (p1 - p) * (p1 - p) / d : p2;
        if (!(pn > 0.0) || !isfinite(pn)) pn = p2;
        bool done = fabs(pn - p) <= 1e-13 * fabs(pn);
        p = pn;
        if (done) break;
    }
    if (!bad) {
        for (int it = 0; it < 60; ++it) {
            double pn = g(p);
            if (isnan(pn)) { bad = true; break; }
            bool done = (pn == p) || fabs(pn - p) <= 4e-16 * fabs(p);
            p = pn;
            if (done) break;
        }
    }
    if (bad) p = qnan();
    // np.isclose(lcl_p, pressure): |a - b| <= atol + rtol * |b|
    if (fabs(p - p0) <= 1e-8 + 1e-5 * fabs(p0)) p = p0;
    lcl_p = p;
    lcl_t = dewpoint_from_e(vapor_pressure(p, w));
}

XP_HD double sign_of(double v) {   // np.sign: NaN stays NaN
    return (v > 0.0) ? 1.0 : ((v < 0.0) ? -1.0 : ((v == 0.0) ? 0.0 : qnan()));
}

}  // namespace xp
