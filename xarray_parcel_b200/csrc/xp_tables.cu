// xp_tables.cu -- moist-adiabat lookup tables built on the GPU.
//
// Replaces moist_adiabat_lookup (PF:447-523): the reference integrates 14 300 pseudo-adiabats
// with metpy.calc.moist_lapse (SciPy ODE solver, ~100 s) and marks a (pressure, temperature)
// index grid in two passes per adiabat.  Here one thread owns one adiabat: it integrates
// dT/dln(p) with classical RK4 (sub-stepped so |dln p| <= 0.005; global error < 1e-7 K) from
// 1100 hPa to 2.5 hPa and marks the grid as it goes.  "Later adiabats overwrite earlier ones"
// (PF:488-504, the loop runs in ascending adiabat number) == atomicMax of the adiabat number.
#include "xp_kernels.cuh"

namespace xp {

namespace {

__device__ __forceinline__ double dT_dlnp(double x, double t) {
    const double p = exp(x);
    return moist_lapse_rhs(p, t) * p;
}

constexpr double kMaxDlnp = 0.005;

__global__ void __launch_bounds__(64) build_tables_kernel(uint32_t *__restrict__ grid,
                                                          float *__restrict__ curves) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;     // adiabat number - 1
    if (i >= kNAdiabats) return;
    // PF:478-482: for T in temperatures: for offset in (0, temp_step / 2)
    double t = table_temperature(i >> 1) + ((i & 1) ? 0.02 / 2 : 0.0);
    const uint32_t id = (uint32_t)(i + 1);
    float *curve = curves + (size_t)i * kNP;

    auto mark_node = [&](int k, double tk) {                 // pass 1, PF:484-489
        const double q = rint(tk / 0.02);
        const double kk = q - 8650.0;
        if (kk >= 0.0 && kk < (double)kNT) atomicMax(grid + (size_t)k * kNT + (int)kk, id);
    };

    double x_prev = log(1100.0);
    double t_prev = t;
    curve[kNP - 1] = (float)t;
    mark_node(0, t);
    for (int k = 1; k < kNP; ++k) {
        const double pk = 1100.0 - 0.5 * k;
        const double xk = log(pk);
        const double h_tot = xk - x_prev;
        const int n_sub = max(1, (int)ceil(fabs(h_tot) / kMaxDlnp));
        const double h = h_tot / n_sub;
        double x = x_prev;
        for (int s = 0; s < n_sub; ++s) {
            const double k1 = dT_dlnp(x, t);
            const double k2 = dT_dlnp(x + 0.5 * h, t + 0.5 * h * k1);
            const double k3 = dT_dlnp(x + 0.5 * h, t + 0.5 * h * k2);
            const double k4 = dT_dlnp(x + h, t + h * k3);
            t = t + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4);
            x = x + h;
        }
        curve[kNP - 1 - k] = (float)t;                       // PF:54: ascending pressure
        mark_node(k, t);
        // pass 2, PF:495-504: pressure of this adiabat at every table temperature in
        // [T_k, T_{k-1}) by np.interp on the reversed profile (xp[j] = T_k, xp[j+1] = T_{k-1}).
        const double t_lo = t, t_hi = t_prev;
        if (t_hi > t_lo) {
            const double slope = 0.5 / (t_hi - t_lo);        // (fp[j+1]-fp[j]) / (xp[j+1]-xp[j])
            int m = (int)ceil((t_lo - 173.0) * 50.0) - 1;
            if (m < 0) m = 0;
            while (m < kNT && table_temperature(m) < t_lo) ++m;
            for (; m < kNT; ++m) {
                const double tm = table_temperature(m);
                const bool last_node = (k == 1 && tm == t_hi);   // x == xp[-1] -> fp[-1]
                if (!(tm < t_hi) && !last_node) break;
                double pres;
                if (last_node) pres = 1100.0;
                else if (tm == t_lo) pres = pk;
                else pres = __dadd_rn(__dmul_rn(slope, tm - t_lo), pk);
                const double q = rint(pres / 0.5);           // round_to(pres, 0.5), PF:499
                const double row = 2200.0 - q;               // position of q*0.5 in 1100, 1099.5, ...
                if (row >= 0.0 && row < (double)kNP) atomicMax(grid + (size_t)row * kNT + m, id);
            }
        }
        x_prev = xk;
        t_prev = t;
    }
}

__global__ void pack_grid_kernel(const uint32_t *__restrict__ in, uint16_t *__restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint16_t)in[i];
}

}  // namespace

void launch_build_tables(uint16_t *index_grid, float *curves, uint32_t *scratch_u32,
                         cudaStream_t stream) {
    const size_t cells = (size_t)kNP * kNT;
    cudaMemsetAsync(scratch_u32, 0, cells * sizeof(uint32_t), stream);
    build_tables_kernel<<<(kNAdiabats + 63) / 64, 64, 0, stream>>>(scratch_u32, curves);
    pack_grid_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, stream>>>(scratch_u32, index_grid, cells);
}

}  // namespace xp
