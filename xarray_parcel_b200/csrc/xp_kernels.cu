// xp_kernels.cu -- sm_100a kernels of the parcel path: one thread per column, columns
// contiguous so every level read is a coalesced line.
#include <cstdlib>

#include "xp_kernels.cuh"
#include "xp_parcels.cuh"

namespace xp {

// ---- column readers --------------------------------------------------------------------
template <typename T>
struct GlobalReader {
    const T *p, *t, *td;
    int64_t ls, pls;
    int L;
    int qmode;          // != 0: td holds specific humidity (ColsArg::qmode): converted as the level is loaded, in float64
    __device__ __forceinline__ double P(int k) const { return (double)__ldg(p + (int64_t)k * pls); }
    __device__ __forceinline__ double Tk(int k) const { return (double)__ldg(t + (int64_t)k * ls); }
    __device__ __forceinline__ double Td(int k) const {
        const double raw = (double)__ldg(td + (int64_t)k * ls);
        return qmode ? dewpoint_from_q(P(k), Tk(k), raw, qmode) : raw;          // PF:1889, 1969
    }
};

template <typename T>
__device__ __forceinline__ GlobalReader<T> make_reader(const ColsArg<T> &c, int64_t col) {
    GlobalReader<T> r;
    r.p = c.p1d ? c.p : c.p + col;
    r.t = c.t + col;
    r.td = c.td + col;
    r.ls = c.ls;
    r.pls = c.pls;
    r.L = c.L;
    r.qmode = c.qmode;
    return r;
}

// ---- profile writer ------------------------------------------------------------------------
template <typename T>
struct ProfWriter {
    T *p, *t, *tv, *et, *etv, *etd;
    int64_t ls;
    bool any;
    __device__ __forceinline__ void put(int v, const ProfileRow &r) const {
        if (!any) return;
        const int64_t o = (int64_t)v * ls;
        if (p) p[o] = (T)r.p;
        if (t) t[o] = (T)r.t;
        if (tv) tv[o] = (T)r.tv;
        if (et) et[o] = (T)r.env_t;
        if (etv) etv[o] = (T)r.env_tv;
        if (etd) etd[o] = (T)r.env_td;
    }
};

template <typename T>
__device__ __forceinline__ ProfWriter<T> make_writer(const OutArg<T> &o, int64_t col) {
    ProfWriter<T> w;
    w.p = o.prof_p ? o.prof_p + col : nullptr;
    w.t = o.prof_t ? o.prof_t + col : nullptr;
    w.tv = o.prof_tv ? o.prof_tv + col : nullptr;
    w.et = o.prof_et ? o.prof_et + col : nullptr;
    w.etv = o.prof_etv ? o.prof_etv + col : nullptr;
    w.etd = o.prof_etd ? o.prof_etd + col : nullptr;
    w.ls = o.prof_ls;
    w.any = w.p || w.t || w.tv || w.et || w.etv || w.etd;
    return w;
}

template <typename T>
__device__ __forceinline__ void store_result(const OutArg<T> &o, int64_t col, const ParcelResult &r,
                                             double pp, double pt, double ptd, int shift) {
    if (o.cape) o.cape[col] = (T)r.cape;
    if (o.cin) o.cin[col] = (T)r.cin;
    if (o.lcl_p) o.lcl_p[col] = (T)r.lcl_p;
    if (o.lcl_t) o.lcl_t[col] = (T)r.lcl_t;
    if (o.lcl_tv) o.lcl_tv[col] = (T)r.lcl_tv;
    if (o.lfc_p) o.lfc_p[col] = (T)r.lfc_p;
    if (o.lfc_t) o.lfc_t[col] = (T)r.lfc_t;
    if (o.el_p) o.el_p[col] = (T)r.el_p;
    if (o.el_t) o.el_t[col] = (T)r.el_t;
    if (o.par_p) o.par_p[col] = (T)pp;
    if (o.par_t) o.par_t[col] = (T)pt;
    if (o.par_td) o.par_td[col] = (T)ptd;
    if (o.shift) o.shift[col] = shift;
}

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

// ---- exact fix-up over the list of columns handed over by the float32 fast paths -----------------------------
struct NoProf {
    __device__ __forceinline__ void put(int, const ProfileRow &) const {}
};

// A column staged once in shared memory ([level][thread], conflict-free): the exact path makes ~8 passes over a
// column (layer bounds, theta-e search, LCL bracket, lift, ...), each of which would otherwise fetch the item's
// scattered 32-byte sectors from DRAM again (ncu, 10 M x 90 most-unstable + profile rows: 27.5 GB read for
// 375 k items = 73 KB per item against 1 KB of input).  A shared 1-D pressure axis is staged once per CTA.
struct StagedReader {
    const float *p, *t, *td;
    int ps, s;                  // element strides between levels (p: 1 for the shared axis)
    int L;
    int qmode;                  // as GlobalReader
    __device__ __forceinline__ double P(int k) const { return (double)p[k * ps]; }
    __device__ __forceinline__ double Tk(int k) const { return (double)t[k * s]; }
    __device__ __forceinline__ double Td(int k) const {
        const double raw = (double)td[k * s];
        return qmode ? dewpoint_from_q(P(k), Tk(k), raw, qmode) : raw;
    }
};

// One thread per (column, parcel kind) item.  STAGED: dynamic shared memory holds blockDim.x columns.
template <bool STAGED>
__global__ void __launch_bounds__(128, 3) suite_list_kernel(const __grid_constant__ ListParams prm) {
    extern __shared__ float s_cols[];
    const uint32_t c0 = prm.list_count[0], c1 = prm.list_count[1], c2 = prm.list_count[2];
    const uint64_t total = (uint64_t)c0 + c1 + c2;
    const int L = prm.cols.L, nt = (int)blockDim.x;
    float *s_t = s_cols, *s_td = s_cols + (size_t)L * nt, *s_p = s_cols + (size_t)2 * L * nt;
    if (STAGED && prm.cols.p1d) {
        for (int k = threadIdx.x; k < L; k += nt) s_p[k] = __ldg(prm.cols.p + (int64_t)k * prm.cols.pls);
        __syncthreads();
    }
    for (uint64_t it = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total;
         it += (uint64_t)gridDim.x * blockDim.x) {
        const int kind = it < c0 ? 0 : (it < (uint64_t)c0 + c1 ? 1 : 2);
        const uint64_t idx = it - (kind == 0 ? 0 : (kind == 1 ? c0 : (uint64_t)c0 + c1));
        const uint32_t e = prm.list[(uint64_t)kind * prm.capacity + idx];
        const bool also_mu = kind == 0 && ((e >> 28) & kListMuIsSb);
        const int64_t col = (int64_t)(e & 0x0fffffffu);
        const GlobalReader<float> rd = make_reader(prm.cols, col);
        ParcelResult r;
        double p0, t0, td0;
        int shift;
        ProfWriter<float> np = make_writer(prm.outs[kind], col);       // profile rows too, where requested ...
        if ((e >> 28) & kListRowsOk) np.any = false;                   // ... unless the float32 rows stand
        if (STAGED) {
            // independent loads, all in flight at once; only this thread reads its slots back: no barrier needed
            for (int k = 0; k < L; ++k) {
                s_t[k * nt + threadIdx.x] = __ldg(rd.t + (int64_t)k * rd.ls);
                s_td[k * nt + threadIdx.x] = __ldg(rd.td + (int64_t)k * rd.ls);
                if (!prm.cols.p1d) s_p[k * nt + threadIdx.x] = __ldg(rd.p + (int64_t)k * rd.pls);
            }
            StagedReader sr;
            sr.p = prm.cols.p1d ? s_p : s_p + threadIdx.x; sr.ps = prm.cols.p1d ? 1 : nt;
            sr.t = s_t + threadIdx.x; sr.td = s_td + threadIdx.x; sr.s = nt; sr.L = L; sr.qmode = prm.cols.qmode;
            run_column(sr, kind, prm.tb, prm.o, qnan(), qnan(), qnan(), r, p0, t0, td0, shift, np);
        } else {
            run_column(rd, kind, prm.tb, prm.o, qnan(), qnan(), qnan(), r, p0, t0, td0, shift, np);
        }
        for (int w = 0; w < (also_mu ? 2 : 1); ++w) store_result(prm.outs[w == 0 ? kind : 2], col, r, p0, t0, td0, shift);
        if (r.flags && prm.flags) atomicOr(prm.flags, r.flags);
    }
}

void launch_suite_list(const ListParams &lp, int sm_count, cudaStream_t stream) {
    // staged variant when 3 CTAs per SM fit (the register budget allows no more): 128 threads, else 64
    const size_t per_thread = (size_t)lp.cols.L * (lp.cols.p1d ? 2 : 3) * sizeof(float);
    const size_t axis = lp.cols.p1d ? (size_t)lp.cols.L * sizeof(float) : 0;
    const size_t budget = 72 * 1024;
    static const bool staged = getenv("XP_LIST_STAGED") && atoi(getenv("XP_LIST_STAGED")) == 1;   // A/B knob, off by default until measured
    int threads = 0;
    if (!staged) threads = 0;
    else if (per_thread * 128 + axis <= budget) threads = 128;
    else if (per_thread * 64 + axis <= budget) threads = 64;
    if (threads) {
        const size_t smem = per_thread * threads + axis;
        // per device, so set on every launch (a host-side call of about a microsecond)
        cudaFuncSetAttribute(suite_list_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
        suite_list_kernel<true><<<sm_count * 8 * (128 / threads), threads, smem, stream>>>(lp);
    } else {
        suite_list_kernel<false><<<sm_count * 8, 128, 0, stream>>>(lp);
    }
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
