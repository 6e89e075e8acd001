// xp_layers.cu -- the layer primitives of the parcel selectors as stand-alone column kernels:
//   mixed_layer_kernel  : mixed_layer (PF:137-162) = (1 / pressure depth) * trapz(x = 'pressure') (PF:164-206)
//                         over get_layer(interpolate=True) (PF:63-100: layer top = bottom - depth, interpolated in
//                         ln p by log_interp PF:1813-1828 and inserted by insert_level PF:933-990), for up to 4
//                         variables of a column at once
//   mixed_parcel_kernel : mixed_parcel (PF:229-289) with ALL the variables the reference returns (theta,
//                         mixing_ratio, temperature, vapour_pressure, dewpoint, pressure)
//   layer_bounds_kernel : get_layer's bottom / top pressures for both variants (interpolate=True: bottom - depth;
//                         interpolate=False: bound_pressure PF:208-227, the level closest to it, ties -> the larger)
// One thread per column, float64 arithmetic in the reference's operation order (no FMA contraction); level-major
// inputs, so each level read of a warp is one coalesced line.  The same loop runs fused inside the lifting
// kernels (xp_parcels.cuh mixed_parcel); these entry points expose it for the callers of PF that use the
// primitives on their own.
#include "xp_kernels.cuh"
#include "xp_layers.cuh"

namespace xp {

namespace {

template <typename T>
struct MixedLayerParams {
    const T *p;                 // [L][N] or shared [L]
    int64_t pls;
    int p1d;
    const T *x[4];              // variables [L][N]
    T *out[4];                  // [N]
    int n_fields;
    int pressure_field;         // index of the variable that is the pressure itself, or -1 (PF:87)
    int64_t ls;
    int L;
    int64_t n;
    double depth;
};

template <typename T>
__global__ void mixed_layer_kernel(const __grid_constant__ MixedLayerParams<T> prm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.n) return;
    const T *pc = prm.p1d ? prm.p : prm.p + i;
    auto pressure_at = [&](int k) { return (double)pc[(int64_t)k * prm.pls]; };
    auto load = [&](int k, double (&x)[4]) {
#pragma unroll
        for (int f = 0; f < 4; ++f)
            x[f] = f < prm.n_fields ? (double)prm.x[f][(int64_t)k * prm.ls + i] : 0.0;
    };
    double bottom, top, mean[4];
    mixed_layer_means<4>(prm.L, pressure_at, load, prm.depth, bottom, top, mean, prm.pressure_field);
    for (int f = 0; f < prm.n_fields; ++f) prm.out[f][i] = (T)mean[f];
}

template <typename T>
struct MixedParcelParams {
    const T *p, *t, *td;        // [L][N] (p: or shared [L])
    int64_t pls, ls;
    int p1d;
    int L;
    int64_t n;
    double depth;
    T *theta, *mixing_ratio, *temperature, *vapour_pressure, *dewpoint, *pressure;   // [N], any may be null
};

template <typename T>
__global__ void mixed_parcel_kernel(const __grid_constant__ MixedParcelParams<T> prm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.n) return;
    const T *pc = prm.p1d ? prm.p : prm.p + i;
    auto pressure_at = [&](int k) { return (double)pc[(int64_t)k * prm.pls]; };
    auto t_at = [&](int k) { return (double)prm.t[(int64_t)k * prm.ls + i]; };
    auto td_at = [&](int k) { return (double)prm.td[(int64_t)k * prm.ls + i]; };
    double o[6];
    mixed_parcel_full(prm.L, pressure_at, t_at, td_at, prm.depth, o);
    if (prm.theta) prm.theta[i] = (T)o[0];
    if (prm.mixing_ratio) prm.mixing_ratio[i] = (T)o[1];
    if (prm.temperature) prm.temperature[i] = (T)o[2];
    if (prm.vapour_pressure) prm.vapour_pressure[i] = (T)o[3];
    if (prm.dewpoint) prm.dewpoint[i] = (T)o[4];
    if (prm.pressure) prm.pressure[i] = (T)o[5];
}

template <typename T>
struct LayerBoundsParams {
    const T *p;
    int64_t pls;
    int p1d;
    int L;
    int64_t n;
    double depth;
    int interpolate;
    T *bottom, *top;            // [N]
};

template <typename T>
__global__ void layer_bounds_kernel(const __grid_constant__ LayerBoundsParams<T> prm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.n) return;
    const T *pc = prm.p1d ? prm.p : prm.p + i;
    auto pressure_at = [&](int k) { return (double)pc[(int64_t)k * prm.pls]; };
    double bottom, top;
    layer_bounds(prm.L, pressure_at, prm.depth, prm.interpolate != 0, bottom, top);
    if (prm.bottom) prm.bottom[i] = (T)bottom;
    if (prm.top) prm.top[i] = (T)top;
}

}  // namespace

template <typename T>
void launch_mixed_layer(const T *p, int64_t pls, int p1d, const T *const *x, T *const *out, int n_fields,
                        int pressure_field, int64_t ls, int L, int64_t n, double depth, cudaStream_t stream) {
    if (n <= 0 || n_fields <= 0) return;
    MixedLayerParams<T> prm;
    prm.p = p; prm.pls = pls; prm.p1d = p1d; prm.n_fields = n_fields; prm.pressure_field = pressure_field; prm.ls = ls; prm.L = L; prm.n = n;
    prm.depth = depth;
    for (int f = 0; f < 4; ++f) { prm.x[f] = f < n_fields ? x[f] : nullptr; prm.out[f] = f < n_fields ? out[f] : nullptr; }
    mixed_layer_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(prm);
}

template <typename T>
void launch_mixed_parcel(const T *p, int64_t pls, int p1d, const T *t, const T *td, int64_t ls, int L, int64_t n,
                         double depth, T *const *out6, cudaStream_t stream) {
    if (n <= 0) return;
    MixedParcelParams<T> prm;
    prm.p = p; prm.t = t; prm.td = td; prm.pls = pls; prm.ls = ls; prm.p1d = p1d; prm.L = L; prm.n = n;
    prm.depth = depth;
    prm.theta = out6[0]; prm.mixing_ratio = out6[1]; prm.temperature = out6[2]; prm.vapour_pressure = out6[3];
    prm.dewpoint = out6[4]; prm.pressure = out6[5];
    mixed_parcel_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(prm);
}

template <typename T>
void launch_layer_bounds(const T *p, int64_t pls, int p1d, int L, int64_t n, double depth, int interpolate,
                         T *bottom, T *top, cudaStream_t stream) {
    if (n <= 0) return;
    LayerBoundsParams<T> prm;
    prm.p = p; prm.pls = pls; prm.p1d = p1d; prm.L = L; prm.n = n; prm.depth = depth; prm.interpolate = interpolate;
    prm.bottom = bottom; prm.top = top;
    layer_bounds_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(prm);
}

#define XP_INST_LAYERS(T)                                                                                             \
    template void launch_mixed_layer<T>(const T *, int64_t, int, const T *const *, T *const *, int, int, int64_t,    \
                                        int, int64_t, double, cudaStream_t);                                               \
    template void launch_mixed_parcel<T>(const T *, int64_t, int, const T *, const T *, int64_t, int, int64_t,       \
                                         double, T *const *, cudaStream_t);                                           \
    template void launch_layer_bounds<T>(const T *, int64_t, int, int, int64_t, double, int, T *, T *, cudaStream_t);
XP_INST_LAYERS(float)
XP_INST_LAYERS(double)
#undef XP_INST_LAYERS

}  // namespace xp
