// xp_list.cu -- the exact (float64) fix-up over the list of (column, parcel kind) items that the float32 fast paths
// hand over because a decision fell inside their error margin.
//
// One thread per item runs the SAME column code as cape_cin_kernel (xp_column.cuh / xp_parcels.cuh), with libm
// exp / log / pow, no FMA contraction (like xp_kernels.cu).  Measured and dropped: building this file with
// XP_F64_FAST_MATH (branch-free ~3-ulp exp / log / pow of xp_math.cuh) -- the kernel's duration did not move
// (0.205 vs 0.194 ms for the 31 k items of the ERA5 bench step): its time is not in the libm calls.
#include <cstdlib>

#include "xp_kernels_common.cuh"

namespace xp {

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
template <bool STAGED, typename T = float>
__global__ void __launch_bounds__(128, 3) suite_list_kernel(const __grid_constant__ ListParamsT<T> prm) {
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
        const GlobalReader<T> rd = make_reader(prm.cols, col);
        ParcelResult r;
        double p0, t0, td0;
        int shift;
        ProfWriter<T> np = make_writer(prm.outs[kind], col);       // profile rows too, where requested ...
        if ((e >> 28) & kListRowsOk) np.any = false;                   // ... unless the float32 rows stand
        if constexpr (STAGED) {
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

void launch_suite_list(const ListParamsT<double> &lp, int sm_count, cudaStream_t stream) {
    suite_list_kernel<false, double><<<sm_count * 8, 128, 0, stream>>>(lp);
}

}  // namespace xp
