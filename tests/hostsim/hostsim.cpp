// hostsim.cpp -- TEST-ONLY host build of the per-column kernel code.
//
// Compiles xarray_parcel_b200/csrc/xp_parcels.cuh (the exact functions the sm_100a kernels
// inline) with g++ so that tests/test_hostsim_vs_oracle.py can check the kernel *logic*
// against the oracle on a machine without a GPU.  This library is built into
// tests/hostsim/_build, is never imported by the product package and is not a CPU fallback:
// the product path (libxparcel.so) has no host implementation.
#include <cstdint>

#include "../../xarray_parcel_b200/csrc/xp_parcels.cuh"

namespace {

struct HostReader {
    const double *p, *t, *td;
    int64_t ls, pls;
    int L;
    double P(int k) const { return p[(int64_t)k * pls]; }
    double Tk(int k) const { return t[(int64_t)k * ls]; }
    double Td(int k) const { return td[(int64_t)k * ls]; }
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
    o.compat = iopts[4]; o.ml_depth = ml_depth; o.mu_depth = mu_depth;
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
