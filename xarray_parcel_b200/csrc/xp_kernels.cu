// xp_kernels.cu -- sm_100a kernels of the parcel path: one thread per column, columns
// contiguous so every level read is a coalesced line.
#include <cstdlib>

#include "xp_kernels.cuh"
#include "xp_parcels.cuh"
#include "xp_kernels_common.cuh"

namespace xp {

// ---- the fused kernel: parcel selection + lift, for every requested parcel kind ---------
template <typename T>
struct CapeCinParams {
    ColsArg<T> cols;
    Tables tb;
    Opts o;
    int kind_mask;
    OutArg<T> outs[4];
    ParcelArg<T> ex;
    uint32_t *flags;
};

template <typename T>
__global__ void __launch_bounds__(128) cape_cin_kernel(const __grid_constant__ CapeCinParams<T> prm) {
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= prm.cols.n) return;
    const GlobalReader<T> rd = make_reader(prm.cols, col);
    uint32_t flags = 0;
#pragma unroll 1
    for (int kind = 0; kind < 4; ++kind) {
        if (!(prm.kind_mask & (1 << kind))) continue;
        const OutArg<T> &out = prm.outs[kind];
        double ex_p = qnan(), ex_t = qnan(), ex_td = qnan();
        if (kind == 3) {
            ex_p = (double)prm.ex.p[col]; ex_t = (double)prm.ex.t[col]; ex_td = (double)prm.ex.td[col];
        }
        ParcelResult r;
        double p0, t0, td0;
        int shift;
        ProfWriter<T> w = make_writer(out, col);
        run_column(rd, kind, prm.tb, prm.o, ex_p, ex_t, ex_td, r, p0, t0, td0, shift, w);
        flags |= r.flags;
        store_result(out, col, r, p0, t0, td0, shift);
    }
    if (flags && prm.flags) atomicOr(prm.flags, flags);
}

template <typename T>
void launch_cape_cin(const ColsArg<T> &cols, const Tables &tb, const Opts &o, int kind_mask,
                     const OutArg<T> *outs, const ParcelArg<T> &ex, uint32_t *flags,
                     cudaStream_t stream) {
    if (cols.n <= 0) return;
    CapeCinParams<T> prm;
    prm.cols = cols; prm.tb = tb; prm.o = o; prm.kind_mask = kind_mask;
    for (int i = 0; i < 4; ++i) prm.outs[i] = outs[i];
    prm.ex = ex; prm.flags = flags;
    const int block = 128;
    const int64_t grid = (cols.n + block - 1) / block;
    cape_cin_kernel<T><<<(unsigned)grid, block, 0, stream>>>(prm);
}

// ---- individually exposed steps -------------------------------------------------------------
template <typename T>
__global__ void lcl_kernel(const T *p, const T *t, const T *td, int64_t n, Opts o, T *lcl_p,
                           T *lcl_t, T *lcl_tv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double p0 = (double)p[i], t0 = (double)t[i], td0 = (double)td[i];
    double lp = qnan(), lt = qnan(), ltv = qnan();
    if (!(isnan(p0) || isnan(t0) || isnan(td0))) {              // PF:627-634, 680
        lcl_solve(p0, t0, td0, lp, lt);
        ltv = virtual_temperature(lt, mixing_ratio_t_td(lt, lt, lp, o.compat));
    }
    if (lcl_p) lcl_p[i] = (T)lp;
    if (lcl_t) lcl_t[i] = (T)lt;
    if (lcl_tv) lcl_tv[i] = (T)ltv;
}

template <typename T>
void launch_lcl(const T *p, const T *t, const T *td, int64_t n, const Opts &o, T *lcl_p, T *lcl_t,
                T *lcl_tv, cudaStream_t stream) {
    if (n <= 0) return;
    lcl_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(p, t, td, n, o, lcl_p, lcl_t, lcl_tv);
}

template <typename T>
__global__ void moist_lapse_kernel(const T *pressure, int64_t ls, int L, int64_t n, Tables tb,
                                   const T *parcel_t, const T *parcel_p, T *out, int64_t out_ls) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double t0 = (double)parcel_t[i], p0 = (double)parcel_p[i];
    const int adiabat = adiabat_lookup(tb, p0, t0);
    const float *curve = tb.curves + (size_t)(adiabat > 0 ? adiabat - 1 : 0) * kNP;
    for (int k = 0; k < L; ++k) {
        const double p = (double)pressure[(int64_t)k * ls + i];
        out[(int64_t)k * out_ls + i] = (T)((adiabat > 0) ? adiabat_temperature(curve, p) : qnan());
    }
}

template <typename T>
void launch_moist_lapse(const T *pressure, int64_t ls, int L, int64_t n, const Tables &tb,
                        const T *parcel_t, const T *parcel_p, T *out, int64_t out_ls,
                        cudaStream_t stream) {
    if (n <= 0) return;
    moist_lapse_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(
        pressure, ls, L, n, tb, parcel_t, parcel_p, out, out_ls);
}

// parcel_profile PF:712-780 (no LCL level inserted).
template <typename T>
__global__ void parcel_profile_kernel(const T *pressure, int64_t ls, int L, int64_t n, Tables tb,
                                      Opts o, ParcelArg<T> parcel, T *out_t, T *out_tv,
                                      int64_t out_ls, T *lcl_p_o, T *lcl_t_o, T *lcl_tv_o) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double p0 = (double)parcel.p[i], t0 = (double)parcel.t[i], td0 = (double)parcel.td[i];
    double lp = qnan(), lt = qnan(), ltv = qnan();
    if (!(isnan(p0) || isnan(t0) || isnan(td0))) {
        lcl_solve(p0, t0, td0, lp, lt);
        ltv = virtual_temperature(lt, mixing_ratio_t_td(lt, lt, lp, o.compat));
    }
    if (lcl_p_o) lcl_p_o[i] = (T)lp;
    if (lcl_t_o) lcl_t_o[i] = (T)lt;
    if (lcl_tv_o) lcl_tv_o[i] = (T)ltv;
    const double w_parcel = mixing_ratio_t_td(t0, td0, p0, o.compat);
    const int adiabat = adiabat_lookup(tb, lp, lt);
    const float *curve = tb.curves + (size_t)(adiabat > 0 ? adiabat - 1 : 0) * kNP;
    for (int k = 0; k < L; ++k) {
        const double p = (double)pressure[(int64_t)k * ls + i];
        const double above = (adiabat > 0) ? adiabat_temperature(curve, p) : qnan();
        double tp, wp;
        if (p >= lp) tp = dry_lapse(p, t0, p0); else tp = above;
        if (p <= lp) wp = sat_mixing_ratio(p, above); else wp = w_parcel;
        if (out_t) out_t[(int64_t)k * out_ls + i] = (T)tp;
        if (out_tv) out_tv[(int64_t)k * out_ls + i] = (T)virtual_temperature(tp, wp);
    }
}

template <typename T>
void launch_parcel_profile(const T *pressure, int64_t ls, int L, int64_t n, const Tables &tb,
                           const Opts &o, const ParcelArg<T> &parcel, T *out_t, T *out_tv,
                           int64_t out_ls, T *lcl_p, T *lcl_t, T *lcl_tv, cudaStream_t stream) {
    if (n <= 0) return;
    parcel_profile_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(
        pressure, ls, L, n, tb, o, parcel, out_t, out_tv, out_ls, lcl_p, lcl_t, lcl_tv);
}

// lfc_el PF:1066-1198 on caller-supplied curves.
template <typename T>
__global__ void lfc_el_kernel(const T *pressure, const T *parcel_t, const T *env_t, int64_t ls,
                              int L, int64_t n, const T *lcl_p, const T *lcl_t, T *lfc_p, T *lfc_t,
                              T *el_p, T *el_t, uint32_t *flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Sweep sw;
    sw.init((double)lcl_p[i], (double)lcl_t[i], 1);
    for (int k = 0; k < L; ++k) {
        const int64_t o = (int64_t)k * ls + i;
        sw.emit((double)pressure[o], (double)parcel_t[o], (double)env_t[o], false);
    }
    ParcelResult r;
    r.flags = 0;
    sw.finish(r, 0);
    if (lfc_p) lfc_p[i] = (T)r.lfc_p;
    if (lfc_t) lfc_t[i] = (T)r.lfc_t;
    if (el_p) el_p[i] = (T)r.el_p;
    if (el_t) el_t[i] = (T)r.el_t;
    if (r.flags && flags) atomicOr(flags, r.flags);
}

template <typename T>
void launch_lfc_el(const T *pressure, const T *parcel_t, const T *env_t, int64_t ls, int L,
                   int64_t n, const T *lcl_p, const T *lcl_t, T *lfc_p, T *lfc_t, T *el_p, T *el_t,
                   uint32_t *flags, cudaStream_t stream) {
    if (n <= 0) return;
    lfc_el_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(
        pressure, parcel_t, env_t, ls, L, n, lcl_p, lcl_t, lfc_p, lfc_t, el_p, el_t, flags);
}

// cape_cin_base PF:1291-1392 with caller-supplied LFC/EL: the inclusion tests are evaluated
// literally per trapezoid / triangle (the LFC and EL need not be crossings of these curves).
template <typename T>
__global__ void cape_cin_base_kernel(const T *pressure, const T *env_t, const T *parcel_t,
                                     int64_t ls, int L, int64_t n, const T *lfc_p_in,
                                     const T *el_p_in, Opts o, T *cape_o, T *cin_o) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double lfc = (double)lfc_p_in[i];
    double el = (double)el_p_in[i];
    if (isnan(el)) {                                              // PF:1329
        double m = qnan();
        for (int k = 0; k < L; ++k) {
            double p = (double)pressure[(int64_t)k * ls + i];
            if (!isnan(p) && !(p >= m)) m = p;
        }
        el = m;
    }
    double cape = 0.0, cin = 0.0;
    auto add = [&](double area, bool in_cape, bool in_cin) {
        if (isnan(area)) return;
        if (in_cape && (!o.pos_neg || area > 0.0)) cape += area;
        if (in_cin && (!o.pos_neg || area < 0.0)) cin += area;
    };
    double pp = qnan(), xp_ = qnan(), dp_ = qnan();
    for (int k = 0; k < L; ++k) {
        const int64_t off = (int64_t)k * ls + i;
        const double p = (double)pressure[off];
        const double d1 = (double)parcel_t[off] - (double)env_t[off];
        const double x = log(p);
        if (k > 0) {
            const double d0 = dp_;
            const double s0 = sign_of(d0), s1 = sign_of(d1);
            bool masked = false;
            if ((s0 == s0) && (s1 == s1) && (s0 != s1)) {
                const double ix = (d1 * xp_ - d0 * x) / (d1 - d0);
                const double frac = (ix - xp_) / (x - xp_);
                const double zy = frac * (d1 - d0) + d0;
                const double zx = log(exp(ix));
                if (!isnan(zy)) {
                    const double dx_lo = xp_ - zx, dx_hi = x - zx;
                    const double a_lo = (d0 / 2) * fabs(dx_lo), a_hi = (d1 / 2) * fabs(dx_hi);
                    const double m_lo = exp(xp_ - dx_lo / 2), m_hi = exp(x - dx_hi / 2);   // PF:1262, 1347
                    masked = !isnan(a_lo);
                    add(a_lo, (m_lo <= lfc) && (m_lo >= el), m_lo >= lfc);                // PF:1354-1376
                    add(a_hi, (m_hi <= lfc) && (m_hi >= el), m_hi >= lfc);
                }
            }
            if (!masked) {
                const double area = fabs(x - xp_) * ((d0 + d1) / 2);
                const bool c0 = (pp <= lfc) && (pp >= el), c1 = (p <= lfc) && (p >= el);  // PF:1352-1353
                add(area, c0 && c1, (pp >= lfc) && (p >= lfc));                          // PF:1371
            }
        }
        pp = p; xp_ = x; dp_ = d1;
    }
    cape *= kRd; cin *= kRd;
    if (o.post_zero && !(cin <= 0.0)) cin = 0.0;
    if (cape_o) cape_o[i] = (T)cape;
    if (cin_o) cin_o[i] = (T)cin;
}

template <typename T>
void launch_cape_cin_base(const T *pressure, const T *env_t, const T *parcel_t, int64_t ls, int L,
                          int64_t n, const T *lfc_p, const T *el_p, const Opts &o, T *cape, T *cin,
                          cudaStream_t stream) {
    if (n <= 0) return;
    cape_cin_base_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(
        pressure, env_t, parcel_t, ls, L, n, lfc_p, el_p, o, cape, cin);
}

// ---- explicit instantiations ------------------------------------------------------------------
#define XP_INST(T)                                                                                 \
    template void launch_cape_cin<T>(const ColsArg<T> &, const Tables &, const Opts &, int,        \
                                     const OutArg<T> *, const ParcelArg<T> &, uint32_t *,          \
                                     cudaStream_t);                                                \
    template void launch_lcl<T>(const T *, const T *, const T *, int64_t, const Opts &, T *, T *,  \
                                T *, cudaStream_t);                                                \
    template void launch_moist_lapse<T>(const T *, int64_t, int, int64_t, const Tables &,          \
                                        const T *, const T *, T *, int64_t, cudaStream_t);         \
    template void launch_parcel_profile<T>(const T *, int64_t, int, int64_t, const Tables &,       \
                                           const Opts &, const ParcelArg<T> &, T *, T *, int64_t,  \
                                           T *, T *, T *, cudaStream_t);                           \
    template void launch_lfc_el<T>(const T *, const T *, const T *, int64_t, int, int64_t,         \
                                   const T *, const T *, T *, T *, T *, T *, uint32_t *,           \
                                   cudaStream_t);                                                  \
    template void launch_cape_cin_base<T>(const T *, const T *, const T *, int64_t, int, int64_t,  \
                                          const T *, const T *, const Opts &, T *, T *,            \
                                          cudaStream_t);
XP_INST(float)
XP_INST(double)

}  // namespace xp
