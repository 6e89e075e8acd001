// xp_levels.cu -- level primitives of the parcel path as stand-alone column kernels (per-column code in
// xp_levels.cuh): insert_level_kernel (PF:933-990), shift_out_nans_kernel (PF:1699-1720), trapz_kernel (PF:164-206)
// pressure_order_kernel (valid_data PF:2308-2321) and find_intersections_kernel (PF:992-1064).  One thread per column, float64 arithmetic, level-major
// arrays so every level access of a warp is one coalesced line; each input level is read once (insert_level keeps
// the previous level in registers), each output level written once.
#include "xp_kernels.cuh"
#include "xp_levels.cuh"

namespace xp {

namespace {

template <typename T>
struct InsertParams {
    const T *coords;            // [L][N] or shared [L]
    int64_t cls;
    int c1d;
    const T *lev_c;             // [N] coordinate of the new level
    const T *x[4];              // variables [L][N]
    const T *lev_x[4];          // [N] their values at the new level
    T *out[4];                  // [L+1][N]
    T *coords_out;              // [L+1][N] or null
    int n_fields;
    int64_t ls, ols;
    int L;
    int64_t n;
};

template <typename T>
__global__ void insert_level_kernel(const __grid_constant__ InsertParams<T> prm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.n) return;
    const T *c = prm.c1d ? prm.coords : prm.coords + i;
    const double lev_c = (double)prm.lev_c[i];
    double lev_v[4], v_jm[4] = {0, 0, 0, 0};
    for (int f = 0; f < 4; ++f) lev_v[f] = f < prm.n_fields ? (double)prm.lev_x[f][i] : 0.0;
    double c_jm = qnan();
    for (int j = 0; j <= prm.L; ++j) {
        const double c_j = j < prm.L ? (double)c[(int64_t)j * prm.cls] : qnan();
        if (prm.coords_out)
            prm.coords_out[(int64_t)j * prm.ols + i] = (T)insert_level_value(j, prm.L, c_j, c_jm, c_j, c_jm, lev_c, lev_c);
        for (int f = 0; f < prm.n_fields; ++f) {
            const double v_j = j < prm.L ? (double)prm.x[f][(int64_t)j * prm.ls + i] : qnan();
            prm.out[f][(int64_t)j * prm.ols + i] = (T)insert_level_value(j, prm.L, c_j, c_jm, v_j, v_jm[f], lev_c, lev_v[f]);
            v_jm[f] = v_j;
        }
        c_jm = c_j;
    }
}

template <typename T>
struct ShiftParams {
    const T *ref;               // [L][N] the variable whose leading NaNs are removed
    const T *x[4];
    T *out[4];
    int n_fields;
    int64_t ls;
    int L;
    int64_t n;
    int32_t *shift;             // [N] or null
};

template <typename T>
__global__ void shift_out_nans_kernel(const __grid_constant__ ShiftParams<T> prm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.n) return;
    auto ref_at = [&](int k) { return (double)prm.ref[(int64_t)k * prm.ls + i]; };
    const int k0 = leading_nans(prm.L, ref_at);
    if (prm.shift) prm.shift[i] = k0;
    for (int f = 0; f < prm.n_fields; ++f)
        for (int j = 0; j < prm.L; ++j)
            prm.out[f][(int64_t)j * prm.ls + i] = j + k0 < prm.L ? prm.x[f][(int64_t)(j + k0) * prm.ls + i] : (T)qnan();
}

template <typename T>
struct TrapzParams {
    const T *x;                 // integration variable [L][N] or shared [L]
    int64_t xls;
    int x1d;
    const T *v[4];
    T *out[4];                  // [N]
    int n_fields;
    int64_t ls;
    int L;
    int64_t n;
    const uint8_t *mask;        // [>= L-1][N] (non-zero = include the interval above this level) or null
    int64_t mls;
    int sign;
};

template <typename T>
__global__ void trapz_kernel(const __grid_constant__ TrapzParams<T> prm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.n) return;
    const T *xc = prm.x1d ? prm.x : prm.x + i;
    auto x_at = [&](int k) { return (double)xc[(int64_t)k * prm.xls]; };
    auto use = [&](int k) { return !prm.mask || prm.mask[(int64_t)k * prm.mls + i] != 0; };
    for (int f = 0; f < prm.n_fields; ++f) {
        const T *v = prm.v[f];
        auto v_at = [&](int k) { return (double)v[(int64_t)k * prm.ls + i]; };
        prm.out[f][i] = (T)trapz_column(prm.L, x_at, v_at, use, prm.sign);
    }
}

template <typename T>
struct IntersectParams {
    const T *x;                 // [L][N] or shared [L]
    int64_t xls;
    int x1d;
    const T *a, *b;             // [L][N]
    int64_t ls, ols;
    int L;
    int64_t n;
    int log_x;
    T *out[6];                  // all x, all y, increasing x, y, decreasing x, y: [L-1][N], any may be null
};

template <typename T>
__global__ void find_intersections_kernel(const __grid_constant__ IntersectParams<T> prm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.n) return;
    const T *xc = prm.x1d ? prm.x : prm.x + i;
    auto x_at = [&](int k) { const double v = (double)xc[(int64_t)k * prm.xls]; return prm.log_x ? log(v) : v; };
    double x0 = x_at(0), a0 = (double)prm.a[i], b0 = (double)prm.b[i];
    for (int k = 1; k < prm.L; ++k) {
        const double x1 = x_at(k), a1 = (double)prm.a[(int64_t)k * prm.ls + i], b1 = (double)prm.b[(int64_t)k * prm.ls + i];
        double ix, iy, sc;
        interval_crossing(x0, x1, a0, a1, b0, b1, prm.log_x != 0, ix, iy, sc);
        const int64_t o = (int64_t)(k - 1) * prm.ols + i;
        if (prm.out[0]) prm.out[0][o] = (T)ix;
        if (prm.out[1]) prm.out[1][o] = (T)iy;
        if (prm.out[2]) prm.out[2][o] = (T)(sc > 0 ? ix : qnan());
        if (prm.out[3]) prm.out[3][o] = (T)(sc > 0 ? iy : qnan());
        if (prm.out[4]) prm.out[4][o] = (T)(sc < 0 ? ix : qnan());
        if (prm.out[5]) prm.out[5][o] = (T)(sc < 0 ? iy : qnan());
        x0 = x1; a0 = a1; b0 = b1;
    }
}

template <typename T>
struct ZeroAreasParams {
    const T *x;                 // [L][N] or shared [L]
    int64_t xls;
    int x1d;
    const T *y;                 // [L][N]
    int64_t ls, ols;
    int L;
    int64_t n;
    int log_x;
    T *out[5];                  // area, x, dx, x_from, x_to: [2L-1][N] (rows 0..L-1 before, L..2L-2 after), may be null
    uint8_t *mask;              // [L][N]: 1 where the ordinary interval above this level stays in the integral
};

// trap_around_zeros (PF:1200-1289): one thread per column, previous level in registers, every row written once.
template <typename T>
__global__ void trap_around_zeros_kernel(const __grid_constant__ ZeroAreasParams<T> prm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.n) return;
    const T *xc = prm.x1d ? prm.x : prm.x + i;
    const int L = prm.L;
    auto put = [&](int row, const double (&v)[3]) {
        const int64_t o = (int64_t)row * prm.ols + i;
        if (prm.out[0]) prm.out[0][o] = (T)v[0];
        if (prm.out[1]) prm.out[1][o] = (T)v[1];
        if (prm.out[2]) prm.out[2][o] = (T)v[2];
        if (prm.out[3]) prm.out[3][o] = (T)(v[1] - v[2] / 2);            // PF:1277
        if (prm.out[4]) prm.out[4][o] = (T)(v[1] + v[2] / 2);            // PF:1278
    };
    double x0 = (double)xc[0], y0 = (double)prm.y[i];
    for (int k = 0; k + 1 < L; ++k) {
        const double x1 = (double)xc[(int64_t)(k + 1) * prm.xls], y1 = (double)prm.y[(int64_t)(k + 1) * prm.ls + i];
        double before[3], after[3];
        zero_half_areas(x0, x1, y0, y1, prm.log_x != 0, before, after);
        put(k, before);
        put(L + k, after);
        if (prm.mask) prm.mask[(int64_t)k * prm.n + i] = isnan(before[0]) ? 1 : 0;      // PF:1285-1287
        x0 = x1; y0 = y1;
    }
    const double nanv[3] = {qnan(), qnan(), qnan()};
    put(L - 1, nanv);                                                   // the top level has no interval above it
    if (prm.mask) prm.mask[(int64_t)(L - 1) * prm.n + i] = 1;
}

// interp1d_numba (PF:23-37): row-major core dimensions, one thread per output point; the threads of a warp work on
// neighbouring points of one row, so the binary-search probes of xp are shared cache lines.
template <typename T>
__global__ void interp1d_kernel(const T *at, const T *xp, const T *fp, T *out, int64_t rows, int m, int n, int xp1d) {
    const int64_t total = rows * (int64_t)m;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / m;
        const T *xr = xp1d ? xp : xp + r * (int64_t)n;
        const T *fr = fp + r * (int64_t)n;
        auto xp_at = [&](int k) { return (double)xr[k]; };
        auto fp_at = [&](int k) { return (double)fr[k]; };
        out[i] = (T)interp1d_point((double)at[i], n, xp_at, fp_at);
    }
}

template <typename T>
__global__ void pressure_order_kernel(const T *p, int64_t pls, int p1d, int L, int64_t n, uint32_t *flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int r = 0;
    if (i < n) {
        const T *pc = p1d ? p : p + i;
        auto pressure_at = [&](int k) { return (double)pc[(int64_t)k * pls]; };
        r = pressure_order(L, pressure_at);
    }
    const unsigned bad = __ballot_sync(0xffffffffu, r & 1), seen = __ballot_sync(0xffffffffu, r & 2);
    if ((threadIdx.x & 31) == 0) {
        const uint32_t f = (bad ? kFlagPressureNotDecreasing : 0u) | (seen ? kFlagPressureOrderChecked : 0u);
        if (f) atomicOr(flags, f);
    }
}

}  // namespace

template <typename T>
void launch_insert_level(const T *coords, int64_t cls, int c1d, const T *lev_c, const T *const *x,
                         const T *const *lev_x, T *const *out, T *coords_out, int n_fields, int64_t ls, int64_t ols,
                         int L, int64_t n, cudaStream_t stream) {
    if (n <= 0) return;
    InsertParams<T> prm;
    prm.coords = coords; prm.cls = cls; prm.c1d = c1d; prm.lev_c = lev_c; prm.coords_out = coords_out;
    prm.n_fields = n_fields; prm.ls = ls; prm.ols = ols; prm.L = L; prm.n = n;
    for (int f = 0; f < 4; ++f) {
        prm.x[f] = f < n_fields ? x[f] : nullptr;
        prm.lev_x[f] = f < n_fields ? lev_x[f] : nullptr;
        prm.out[f] = f < n_fields ? out[f] : nullptr;
    }
    insert_level_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(prm);
}

template <typename T>
void launch_shift_out_nans(const T *ref, const T *const *x, T *const *out, int n_fields, int64_t ls, int L,
                           int64_t n, int32_t *shift, cudaStream_t stream) {
    if (n <= 0) return;
    ShiftParams<T> prm;
    prm.ref = ref; prm.n_fields = n_fields; prm.ls = ls; prm.L = L; prm.n = n; prm.shift = shift;
    for (int f = 0; f < 4; ++f) { prm.x[f] = f < n_fields ? x[f] : nullptr; prm.out[f] = f < n_fields ? out[f] : nullptr; }
    shift_out_nans_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(prm);
}

template <typename T>
void launch_trapz(const T *x, int64_t xls, int x1d, const T *const *v, T *const *out, int n_fields, int64_t ls,
                  int L, int64_t n, const uint8_t *mask, int64_t mls, int sign, cudaStream_t stream) {
    if (n <= 0 || n_fields <= 0) return;
    TrapzParams<T> prm;
    prm.x = x; prm.xls = xls; prm.x1d = x1d; prm.n_fields = n_fields; prm.ls = ls; prm.L = L; prm.n = n;
    prm.mask = mask; prm.mls = mls; prm.sign = sign;
    for (int f = 0; f < 4; ++f) { prm.v[f] = f < n_fields ? v[f] : nullptr; prm.out[f] = f < n_fields ? out[f] : nullptr; }
    trapz_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(prm);
}

template <typename T>
void launch_find_intersections(const T *x, int64_t xls, int x1d, const T *a, const T *b, int64_t ls, int64_t ols,
                               int L, int64_t n, int log_x, T *const *out6, cudaStream_t stream) {
    if (n <= 0 || L < 2) return;
    IntersectParams<T> prm;
    prm.x = x; prm.xls = xls; prm.x1d = x1d; prm.a = a; prm.b = b; prm.ls = ls; prm.ols = ols; prm.L = L; prm.n = n;
    prm.log_x = log_x;
    for (int f = 0; f < 6; ++f) prm.out[f] = out6[f];
    find_intersections_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(prm);
}

template <typename T>
void launch_interp1d(const T *at, const T *xp, const T *fp, T *out, int64_t rows, int m, int n, int xp1d,
                     cudaStream_t stream) {
    const int64_t total = rows * (int64_t)m;
    if (total <= 0 || n < 1) return;
    const int64_t blocks = (total + 255) / 256;
    interp1d_kernel<T><<<(unsigned)(blocks < 148 * 64 ? blocks : 148 * 64), 256, 0, stream>>>(at, xp, fp, out, rows, m, n, xp1d);
}

template <typename T>
void launch_trap_around_zeros(const T *x, int64_t xls, int x1d, const T *y, int64_t ls, int64_t ols, int L, int64_t n,
                              int log_x, T *const *out5, uint8_t *mask, cudaStream_t stream) {
    if (n <= 0 || L < 1) return;
    ZeroAreasParams<T> prm;
    prm.x = x; prm.xls = xls; prm.x1d = x1d; prm.y = y; prm.ls = ls; prm.ols = ols; prm.L = L; prm.n = n;
    prm.log_x = log_x; prm.mask = mask;
    for (int f = 0; f < 5; ++f) prm.out[f] = out5[f];
    trap_around_zeros_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(prm);
}

template <typename T>
void launch_pressure_order(const T *p, int64_t pls, int p1d, int L, int64_t n, uint32_t *flags, cudaStream_t stream) {
    if (n <= 0) return;
    pressure_order_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(p, pls, p1d, L, n, flags);
}

#define XP_INST_LEVELS(T)                                                                                             \
    template void launch_insert_level<T>(const T *, int64_t, int, const T *, const T *const *, const T *const *,     \
                                         T *const *, T *, int, int64_t, int64_t, int, int64_t, cudaStream_t);         \
    template void launch_shift_out_nans<T>(const T *, const T *const *, T *const *, int, int64_t, int, int64_t,      \
                                           int32_t *, cudaStream_t);                                                  \
    template void launch_trapz<T>(const T *, int64_t, int, const T *const *, T *const *, int, int64_t, int, int64_t, \
                                  const uint8_t *, int64_t, int, cudaStream_t);                                       \
    template void launch_pressure_order<T>(const T *, int64_t, int, int, int64_t, uint32_t *, cudaStream_t);       \
    template void launch_interp1d<T>(const T *, const T *, const T *, T *, int64_t, int, int, int, cudaStream_t);      \
    template void launch_trap_around_zeros<T>(const T *, int64_t, int, const T *, int64_t, int64_t, int, int64_t,    \
                                              int, T *const *, uint8_t *, cudaStream_t);                              \
    template void launch_find_intersections<T>(const T *, int64_t, int, const T *, const T *, int64_t, int64_t, int, \
                                               int64_t, int, T *const *, cudaStream_t);
XP_INST_LEVELS(float)
XP_INST_LEVELS(double)
#undef XP_INST_LEVELS

}  // namespace xp
