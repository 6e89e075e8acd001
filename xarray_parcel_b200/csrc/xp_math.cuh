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

// Bolton (1980): metpy.calc.saturation_vapor_pressure [PF:258, 698, 760]
XP_HD double sat_vapor_pressure(double t) {
    return kSat0 * exp(17.67 * (t - 273.15) / (t - 29.65));
}
// metpy.calc.dewpoint, degC -> K [PF:280-281, inside metpy.calc.lcl PF:644]
XP_HD double dewpoint_from_e(double e) {
    double val = log(e / kSat0);
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
XP_HD double exner(double p) { return pow(p / 1000.0, kKappa); }   // [PF:269]
XP_HD double potential_temperature(double p, double t) {           // [PF:253]
    return t / exner(p);
}
// PF:291-316
XP_HD double dry_lapse(double p, double t0, double p0) {
    return t0 * pow(p / p0, kKappa);
}
// metpy.calc.equivalent_potential_temperature, Bolton (1980) eq. 39 [PF:123]
XP_HD double theta_e(double p, double t, double td) {
    double r = sat_mixing_ratio(p, td);
    double e = sat_vapor_pressure(td);
    double t_l = 56 + 1. / (1. / (td - 56) + log(t / td) / 800.);
    double th_l = potential_temperature(p - e, t) * pow(t / t_l, 0.28 * r);
    return th_l * exp(r * (1 + 0.448 * r) * (3036. / t_l - 1.78));
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
        return p0 * pow(tdp / t, kInvKappa);
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
