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

}  // namespace

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
