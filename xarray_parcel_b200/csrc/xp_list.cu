// xp_list.cu -- the exact (float64) fix-up over the list of (column, parcel kind) items that the float32 fast paths
// hand over because a decision fell inside their error margin.
//
// One thread per item runs the SAME column code as cape_cin_kernel (xp_column.cuh / xp_parcels.cuh), with libm
// exp / log / pow, no FMA contraction (like xp_kernels.cu).  The kernel lasts as long as one item's dependent float64
// chain; what was measured against that (branch-free math, rows batched or spread over 2 / 4 / 8 lanes, compact code,
// register budgets) is in profiles/r2_list_kernel.md -- none of it is in the tree.
#include <cstdlib>

#include "xp_kernels_common.cuh"

namespace xp {

// ---- exact fix-up over the list of columns handed over by the float32 fast paths -----------------------------
struct NoProf {
    __device__ __forceinline__ void put(int, const ProfileRow &) const {}
};

// The item's PRESSURE column staged once in shared memory ([level][thread], conflict-free) when pressure is per
// column: the float64 column code reads it in up to five passes (column maximum, layer bound, theta-e search, unique
// count, lift), each of which otherwise fetches the item's scattered 32-byte sectors again -- ncu, 10 M x 90
// most-unstable + profile rows: the pressure loads alone are 24 % of the kernel's stall samples and it reads 73 KB per
// item for 8.6 KB of sectors.  Temperature and dewpoint are read in one to two passes and stay in global memory (staging
// all three -- round 1, 138 KB per 128 items at 90 levels -- halved the resident items and gained nothing).  A shared
// 1-D axis is L1-resident and is not staged.
template <typename T>
struct PStagedReader {
    GlobalReader<T> g;
    const float *sp;            // this thread's slots: sp[k * nt]
    int nt, L, qmode;
    __device__ __forceinline__ double P(int k) const { return (double)sp[k * nt]; }
    __device__ __forceinline__ double Tk(int k) const { return g.Tk(k); }
    __device__ __forceinline__ double Td(int k) const {
        const double raw = (double)__ldg(g.td + (int64_t)k * g.ls);
        return qmode ? dewpoint_from_q(P(k), Tk(k), raw, qmode) : raw;
    }
};

// A shared pressure axis: ln p of its levels, tabulated once per CTA in shared memory by the same log() that the
// column code would call per (item, level) -- 8 % of an item's instructions (37 x 86 of 40.7 k per warp, ncu).
template <typename T>
struct AxisLnReader : GlobalReader<T> {
    const double *ln_axis;
    __device__ __forceinline__ double LnP(int k) const { return ln_axis[k]; }
};

// One thread per (column, parcel kind) item.  MODE 0: columns read as they are (per-column pressure, float64 or
// unstaged); 1: float columns with their own pressure, dynamic shared memory holds blockDim.x pressure columns; 2: a
// shared pressure axis, dynamic shared memory holds ln p of its levels.
template <int MODE, typename T = float>
__global__ void __launch_bounds__(128, 3) suite_list_kernel(const __grid_constant__ ListParamsT<T> prm) {
    extern __shared__ __align__(8) float s_p[];
    double *s_ln = reinterpret_cast<double *>(s_p);
    if (MODE == 2) {
        for (int k = threadIdx.x; k < prm.cols.L; k += blockDim.x)
            s_ln[k] = xp_log((double)__ldg(prm.cols.p + (int64_t)k * prm.cols.pls));
        __syncthreads();
    }
    const uint32_t c0 = prm.list_count[0], c1 = prm.list_count[1], c2 = prm.list_count[2];
    const uint64_t total = (uint64_t)c0 + c1 + c2;
    const int L = prm.cols.L, nt = (int)blockDim.x;
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
        if constexpr (MODE == 1) {
            // independent loads, all in flight at once; only this thread reads its slots back: no barrier needed
            for (int k = 0; k < L; ++k) s_p[k * nt + threadIdx.x] = (float)__ldg(rd.p + (int64_t)k * rd.pls);
            PStagedReader<T> sr;
            sr.g = rd; sr.sp = s_p + threadIdx.x; sr.nt = nt; sr.L = L; sr.qmode = prm.cols.qmode;
            run_column(sr, kind, prm.tb, prm.o, qnan(), qnan(), qnan(), r, p0, t0, td0, shift, np);
        } else if constexpr (MODE == 2) {
            AxisLnReader<T> ar;
            static_cast<GlobalReader<T> &>(ar) = rd;
            ar.ln_axis = s_ln;
            run_column(ar, kind, prm.tb, prm.o, qnan(), qnan(), qnan(), r, p0, t0, td0, shift, np);
        } else {
            run_column(rd, kind, prm.tb, prm.o, qnan(), qnan(), qnan(), r, p0, t0, td0, shift, np);
        }
        for (int w = 0; w < (also_mu ? 2 : 1); ++w) store_result(prm.outs[w == 0 ? kind : 2], col, r, p0, t0, td0, shift);
        if (r.flags && prm.flags) atomicOr(prm.flags, r.flags);
    }
}

thread_local cudaEvent_t g_event_before_list = nullptr;

void launch_suite_list(const ListParams &lp, int sm_count, cudaStream_t stream) {
    if (g_event_before_list) cudaEventRecord(g_event_before_list, stream);
    // per-column pressure: staged when three CTAs of 128 items fit one SM (L <= 147 levels)
    const size_t smem = (size_t)lp.cols.L * 128 * sizeof(float);
    static const bool off = getenv("XP_LIST_STAGED") && atoi(getenv("XP_LIST_STAGED")) == 0;      // A/B knob
    if (!lp.cols.p1d && !off && smem <= 75 * 1024) {
        // per device, so set on every launch (a host-side call of about a microsecond)
        cudaFuncSetAttribute(suite_list_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 75 * 1024);
        suite_list_kernel<1><<<sm_count * 8, 128, smem, stream>>>(lp);
    } else if (lp.cols.p1d) {
        suite_list_kernel<2><<<sm_count * 8, 128, (size_t)lp.cols.L * sizeof(double), stream>>>(lp);
    } else {
        suite_list_kernel<0><<<sm_count * 8, 128, 0, stream>>>(lp);
    }
}

void launch_suite_list(const ListParamsT<double> &lp, int sm_count, cudaStream_t stream) {
    if (g_event_before_list) cudaEventRecord(g_event_before_list, stream);
    if (lp.cols.p1d) suite_list_kernel<2, double><<<sm_count * 8, 128, (size_t)lp.cols.L * sizeof(double), stream>>>(lp);
    else suite_list_kernel<0, double><<<sm_count * 8, 128, 0, stream>>>(lp);
}

}  // namespace xp
