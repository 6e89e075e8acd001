// xp_derived.cu -- column kernels behind the reference's derived convective indices:
//   interp_levels_kernel : linear_interp / log_interp (PF:1758-1828) of up to 4 fields at one coordinate
//                          value per column  -> lifted_index (PF:1722), deep_convective_index (PF:1830),
//                          isobar_temperature (PF:2193), lapse_rate (PF:2102), wind_shear (PF:2216)
//   level_crossing_kernel: find_intersections (PF:992-1064, log_x = False) of a field against a constant,
//                          lowest crossing coordinate  -> freezing_level_height / melting_level_height
//                          (PF:2137-2191)
// One thread per column, float64 arithmetic in the reference's operation order (no FMA contraction),
// level-major inputs so each level read of a warp is one coalesced line.
#include "xp_kernels.cuh"
#include "xp_parcels.cuh"

namespace xp {

namespace {

template <typename T>
struct InterpParams {
    const T *coords;            // [L][N] or shared [L]
    int64_t cls;                // level stride of coords
    int c1d;
    const T *x[4];              // fields [L][N]
    T *out[4];                  // [N]
    int n_fields;
    int64_t ls;                 // level stride of the fields
    int L;
    int64_t n;
    const T *at;                // per-column coordinate [N] or null
    double at_scalar;
    int log_coords;
};

// linear_interp(extrapolate=False) PF:1774-1806: bracketing coordinates = min{c >= at} / max{c <= at}
// (NaN-skipping), field values = NaN-skipping mean over the levels that carry that coordinate.
template <typename T>
__global__ void interp_levels_kernel(const __grid_constant__ InterpParams<T> prm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.n) return;
    double at = prm.at ? (double)prm.at[i] : prm.at_scalar;
    if (prm.log_coords) at = log(at);                                    // PF:1828
    const T *c = prm.c1d ? prm.coords : prm.coords + i;
    double cb = qnan(), ca = qnan();
    double sb[4] = {0, 0, 0, 0}, sa[4] = {0, 0, 0, 0};
    int nb[4] = {0, 0, 0, 0}, na[4] = {0, 0, 0, 0};
    for (int k = 0; k < prm.L; ++k) {
        double ck = (double)c[(int64_t)k * prm.cls];
        if (prm.log_coords) ck = log(ck);
        const bool is_b = ck >= at, is_a = ck <= at;
        const bool new_b = is_b && !(ck >= cb), new_a = is_a && !(ck <= ca);   // strictly better (or first)
        if (new_b) { cb = ck; for (int f = 0; f < 4; ++f) { sb[f] = 0; nb[f] = 0; } }
        if (new_a) { ca = ck; for (int f = 0; f < 4; ++f) { sa[f] = 0; na[f] = 0; } }
        const bool acc_b = is_b && ck == cb, acc_a = is_a && ck == ca;
        if (acc_b || acc_a) {
            for (int f = 0; f < prm.n_fields; ++f) {
                const double v = (double)prm.x[f][(int64_t)k * prm.ls + i];
                if (isnan(v)) continue;                                   // .mean() skips NaN (PF:1798-1799)
                if (acc_b) { sb[f] += v; ++nb[f]; }
                if (acc_a) { sa[f] += v; ++na[f]; }
            }
        }
    }
    for (int f = 0; f < prm.n_fields; ++f) {
        const double xb = nb[f] ? sb[f] / nb[f] : qnan();
        const double xa = na[f] ? sa[f] / na[f] : qnan();
        double res = xb + (xa - xb) * ((at - cb) / (ca - cb));           // PF:1802
        if (xb == xa) res = xb;                                           // PF:1806
        prm.out[f][i] = (T)res;
    }
}

template <typename T>
struct CrossParams {
    const T *x;                 // coordinate, e.g. height [L][N] or shared [L]
    int64_t xls;
    int x1d;
    const T *a;                 // field [L][N]
    int64_t ls;
    int L;
    int64_t n;
    double level;               // the constant to intersect with
    T *out;                     // [N] lowest crossing coordinate
};

// find_intersections(x, a, b = level) PF:1019-1053 and .min over the crossings (PF:2153-2154).
template <typename T>
__global__ void level_crossing_kernel(const __grid_constant__ CrossParams<T> prm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.n) return;
    const T *xc = prm.x1d ? prm.x : prm.x + i;
    double best = qnan();
    double x0 = qnan(), a0 = qnan();
    for (int k = 0; k < prm.L; ++k) {
        const double x1 = (double)xc[(int64_t)k * prm.xls];
        const double a1 = (double)prm.a[(int64_t)k * prm.ls + i];
        if (k > 0) {
            const double d0 = a0 - prm.level, d1 = a1 - prm.level;
            const double s0 = sign_of(d0), s1 = sign_of(d1);
            // np.diff(np.sign(a - b)) != 0; NaN differences give NaN coordinates, which .min skips
            if ((s0 == s0) && (s1 == s1) && (s0 != s1)) {
                const double ix = (d1 * x0 - d0 * x1) / (d1 - d0);        // PF:1046
                if (!isnan(ix) && !(ix >= best)) best = ix;
            }
        }
        x0 = x1; a0 = a1;
    }
    prm.out[i] = (T)best;
}

// ---- pointwise kernels around the hot path ------------------------------------------------------
// One thread per point, float64 arithmetic in the reference's operation order, grid-stride.

// metpy.calc.dewpoint_from_specific_humidity as the reference calls it (PF:1889, 1969): MetPy 1.4.1 goes
// through the relative humidity, MetPy >= 1.6 through the vapour pressure (environment_changes_eval.ipynb:278).
template <typename T>
__global__ void dewpoint_from_q_kernel(const T *p, const T *t, const T *q, int64_t n, int compat, T *out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        out[i] = (T)dewpoint_from_q((double)p[i], (double)t[i], (double)q[i], compat);
    }
}

// metpy.calc.saturation_mixing_ratio(p, T) (PF:258; the most-unstable parcel's mixing ratio PF:2047-2053)
template <typename T>
__global__ void sat_mixing_ratio_kernel(const T *p, const T *t, int64_t n, T *out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (T)sat_mixing_ratio((double)p[i], (double)t[i]);
}

// dry_lapse (PF:291-316), mixing_ratio (PF:684-710), virtual_temperature (PF:782-804) as the reference exposes them
template <typename T>
__global__ void dry_lapse_kernel(const T *p, const T *t0, const T *p0, int64_t n, T *out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (T)dry_lapse((double)p[i], (double)t0[i], (double)p0[i]);
}
template <typename T>
__global__ void mixing_ratio_kernel(const T *t, const T *td, const T *p, int64_t n, int compat, T *out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (T)mixing_ratio_t_td((double)t[i], (double)td[i], (double)p[i], compat);
}
template <typename T>
__global__ void virtual_temperature_kernel(const T *t, const T *w, int64_t n, double epsilon, T *out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (T)((double)t[i] * (1 + epsilon * (double)w[i]));
}

// wet_bulb_temperature by Normand's rule (PF:389-445): lift every point to its LCL (PF:609-682), then bring
// it back down the moist adiabat of the lookup tables to its own pressure (moist_lapse PF:525-607).
template <typename T>
__global__ void wet_bulb_kernel(const T *p, const T *t, const T *td, int64_t n, Tables tb, T *out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double pp = (double)p[i], tt = (double)t[i], dd = (double)td[i];
        double r = qnan();
        if (!(isnan(pp) || isnan(tt) || isnan(dd))) {                    // PF:627-634, 680
            double lp, lt;
            lcl_solve(pp, tt, dd, lp, lt);
            const int adiabat = adiabat_lookup(tb, lp, lt);              // PF:554-557
            if (adiabat > 0) r = adiabat_temperature(tb.curves + (size_t)(adiabat - 1) * kNP, pp);
        }
        out[i] = (T)r;
    }
}

// significant_hail_parameter (PF:2261-2306), xarray .where semantics: a failed comparison (also with NaN)
// takes the `other` branch.
__device__ __forceinline__ double ship_value(double mucape, double mr, double lapse, double t500, double shear,
                                             double flh) {
    mr = mr * 1e3;
    lapse = -lapse;
    t500 = t500 - 273.15;
    if (!(shear >= 7)) shear = qnan();
    if (!(shear <= 27)) shear = qnan();
    if (!(mr >= 11)) mr = qnan();
    if (!(mr <= 13.6)) mr = qnan();
    if (!(t500 <= -5.5)) t500 = -5.5;
    double ship = mucape * mr * lapse * -t500 * shear / 42000000;
    if (!(mucape >= 1300)) ship = ship * (mucape / 1300);
    if (!(lapse >= 5.8)) ship = ship * (lapse / 5.8);
    if (!(flh >= 2400)) ship = ship * (flh / 2400);
    return ship;
}

template <typename T>
__global__ void ship_kernel(const T *mucape, const T *mr, const T *lapse, const T *t500, const T *shear, const T *flh,
                            int64_t n, T *out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (T)ship_value((double)mucape[i], (double)mr[i], (double)lapse[i], (double)t500[i], (double)shear[i],
                               (double)flh[i]);
}

// storm_proxies (PF:2323-2407).  Inputs in the order of ProxyIn; outputs: 9 flags (uint8) in the order of the
// reference's `proxies` dict, and SHIP.
template <typename T>
struct ProxyParams {
    const T *in[13];
    uint8_t *flag[9];
    T *ship;
    int64_t n;
};
enum ProxyIn { kMl100Cape, kMl50Cape, kMuCape, kS06, kMl100Li, kMl100Dci, kPosShear, kMl50Cin, kMl100Cin, kLapse,
               kMuMr, kT500, kFlh };

template <typename T>
__global__ void storm_proxies_kernel(const __grid_constant__ ProxyParams<T> prm) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < prm.n; i += (int64_t)gridDim.x * blockDim.x) {
        double v[13];
#pragma unroll
        for (int k = 0; k < 13; ++k) v[k] = (double)prm.in[k][i];
        // negative CAPE is ignored (PF:2337-2339)
        const double ml100 = (v[kMl100Cape] >= 0) ? v[kMl100Cape] : qnan();
        const double ml50 = (v[kMl50Cape] >= 0) ? v[kMl50Cape] : qnan();
        const double mu = (v[kMuCape] >= 0) ? v[kMuCape] : qnan();
        const double s06 = v[kS06];
        const bool pos_shear = !(v[kPosShear] == 0.0);                   // numpy truthiness (NaN is true)
        const bool craven = (ml100 * s06) >= 20000;                                          // PF:2345
        const bool kunz = (v[kMl100Li] <= -2.07) || ((mu >= 1474) || (v[kMl100Dci] >= 25.7));   // PF:2348-2350
        bool trapp = ((ml100 * s06) >= 10000) && (ml100 >= 100);                             // PF:2353-2357
        trapp = trapp && (s06 >= 5);
        trapp = trapp && pos_shear;
        const bool marsh = (ml100 * s06) >= 10000;                                           // PF:2360
        const bool allen11 = (ml50 * pow(s06, 1.67)) >= 25000;                               // PF:2363
        bool allen14 = allen11 && (v[kMl50Cin] > -25);                                       // PF:2366-2371
        allen14 = allen14 && (s06 > 7.5);
        allen14 = allen14 && (v[kLapse] < -6.5);
        const bool eccel = ((ml100 * s06) > 10000) && (v[kMl100Cin] > -50);                  // PF:2374-2375
        bool mohr = (v[kMl100Li] <= -1.6) || (ml100 >= 439);                                 // PF:2378-2381
        mohr = mohr || (v[kMl100Dci] >= 26.4);
        const double ship = ship_value(mu, v[kMuMr], v[kLapse], v[kT500], s06, v[kFlh]);     // PF:2384-2389
        const bool flags[9] = {craven, kunz, trapp, marsh, allen11, allen14, eccel, mohr, ship > 0.1};
#pragma unroll
        for (int k = 0; k < 9; ++k)
            if (prm.flag[k]) prm.flag[k][i] = flags[k] ? 1 : 0;
        if (prm.ship) prm.ship[i] = (T)ship;
    }
}

inline unsigned pointwise_grid(int64_t n) {
    const int64_t g = (n + 255) / 256;
    return (unsigned)(g < 148 * 16 ? g : 148 * 16);
}

}  // namespace

template <typename T>
void launch_dewpoint_from_q(const T *p, const T *t, const T *q, int64_t n, int compat, T *out, cudaStream_t stream) {
    if (n > 0) dewpoint_from_q_kernel<T><<<pointwise_grid(n), 256, 0, stream>>>(p, t, q, n, compat, out);
}
template <typename T>
void launch_sat_mixing_ratio(const T *p, const T *t, int64_t n, T *out, cudaStream_t stream) {
    if (n > 0) sat_mixing_ratio_kernel<T><<<pointwise_grid(n), 256, 0, stream>>>(p, t, n, out);
}
template <typename T>
void launch_dry_lapse(const T *p, const T *t0, const T *p0, int64_t n, T *out, cudaStream_t stream) {
    if (n > 0) dry_lapse_kernel<T><<<pointwise_grid(n), 256, 0, stream>>>(p, t0, p0, n, out);
}
template <typename T>
void launch_mixing_ratio(const T *t, const T *td, const T *p, int64_t n, int compat, T *out, cudaStream_t stream) {
    if (n > 0) mixing_ratio_kernel<T><<<pointwise_grid(n), 256, 0, stream>>>(t, td, p, n, compat, out);
}
template <typename T>
void launch_virtual_temperature(const T *t, const T *w, int64_t n, double epsilon, T *out, cudaStream_t stream) {
    if (n > 0) virtual_temperature_kernel<T><<<pointwise_grid(n), 256, 0, stream>>>(t, w, n, epsilon, out);
}
template <typename T>
void launch_wet_bulb(const T *p, const T *t, const T *td, int64_t n, const Tables &tb, T *out, cudaStream_t stream) {
    if (n > 0) wet_bulb_kernel<T><<<pointwise_grid(n), 256, 0, stream>>>(p, t, td, n, tb, out);
}
template <typename T>
void launch_ship(const T *const *in6, int64_t n, T *out, cudaStream_t stream) {
    if (n > 0) ship_kernel<T><<<pointwise_grid(n), 256, 0, stream>>>(in6[0], in6[1], in6[2], in6[3], in6[4], in6[5], n, out);
}
template <typename T>
void launch_storm_proxies(const T *const *in13, uint8_t *const *flags9, T *ship, int64_t n, cudaStream_t stream) {
    if (n <= 0) return;
    ProxyParams<T> p;
    for (int k = 0; k < 13; ++k) p.in[k] = in13[k];
    for (int k = 0; k < 9; ++k) p.flag[k] = flags9[k];
    p.ship = ship; p.n = n;
    storm_proxies_kernel<T><<<pointwise_grid(n), 256, 0, stream>>>(p);
}
#define XP_INST_POINTWISE(T)                                                                                          \
    template void launch_dewpoint_from_q<T>(const T *, const T *, const T *, int64_t, int, T *, cudaStream_t);        \
    template void launch_sat_mixing_ratio<T>(const T *, const T *, int64_t, T *, cudaStream_t);                       \
    template void launch_dry_lapse<T>(const T *, const T *, const T *, int64_t, T *, cudaStream_t);                   \
    template void launch_mixing_ratio<T>(const T *, const T *, const T *, int64_t, int, T *, cudaStream_t);           \
    template void launch_virtual_temperature<T>(const T *, const T *, int64_t, double, T *, cudaStream_t);            \
    template void launch_wet_bulb<T>(const T *, const T *, const T *, int64_t, const Tables &, T *, cudaStream_t);    \
    template void launch_ship<T>(const T *const *, int64_t, T *, cudaStream_t);                                       \
    template void launch_storm_proxies<T>(const T *const *, uint8_t *const *, T *, int64_t, cudaStream_t);
XP_INST_POINTWISE(float)
XP_INST_POINTWISE(double)
#undef XP_INST_POINTWISE

template <typename T>
void launch_interp_levels(const T *coords, int64_t cls, int c1d, const T *const *x, T *const *out, int n_fields,
                          int64_t ls, int L, int64_t n, const T *at, double at_scalar, int log_coords,
                          cudaStream_t stream) {
    if (n <= 0 || n_fields <= 0) return;
    InterpParams<T> p;
    p.coords = coords; p.cls = cls; p.c1d = c1d; p.n_fields = n_fields; p.ls = ls; p.L = L; p.n = n;
    p.at = at; p.at_scalar = at_scalar; p.log_coords = log_coords;
    for (int f = 0; f < 4; ++f) { p.x[f] = f < n_fields ? x[f] : nullptr; p.out[f] = f < n_fields ? out[f] : nullptr; }
    interp_levels_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(p);
}

template <typename T>
void launch_level_crossing(const T *x, int64_t xls, int x1d, const T *a, int64_t ls, int L, int64_t n,
                           double level, T *out, cudaStream_t stream) {
    if (n <= 0) return;
    CrossParams<T> p;
    p.x = x; p.xls = xls; p.x1d = x1d; p.a = a; p.ls = ls; p.L = L; p.n = n; p.level = level; p.out = out;
    level_crossing_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(p);
}

template void launch_interp_levels<float>(const float *, int64_t, int, const float *const *, float *const *, int,
                                          int64_t, int, int64_t, const float *, double, int, cudaStream_t);
template void launch_interp_levels<double>(const double *, int64_t, int, const double *const *, double *const *,
                                           int, int64_t, int, int64_t, const double *, double, int, cudaStream_t);
template void launch_level_crossing<float>(const float *, int64_t, int, const float *, int64_t, int, int64_t,
                                           double, float *, cudaStream_t);
template void launch_level_crossing<double>(const double *, int64_t, int, const double *, int64_t, int, int64_t,
                                            double, double *, cudaStream_t);

}  // namespace xp
