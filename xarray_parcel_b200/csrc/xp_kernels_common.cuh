// xp_kernels_common.cuh -- column readers, profile writer and result stores shared by the float64 kernels
// (xp_kernels.cu: cape_cin_kernel and the individually exposed steps; xp_list.cu: the exact fix-up of the fast paths).
#pragma once
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

}  // namespace xp
