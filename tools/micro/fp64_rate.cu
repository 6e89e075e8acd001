// DFMA issue rate and dependent-issue latency of one SM (sm_100a): how many float64 FMA warp-instructions per cycle an
// SM sustains with W warps of independent chains, and the cycles per DFMA of ONE dependent chain.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_rate fp64_rate.cu && ./fp64_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma(double *out, long long *cycles, int iters, double a, double b) {
    double x[CHAINS];
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 1 << 16);
    const int iters = 4096;
    long long h;
    for (int threads : {32, 128, 256, 512, 1024}) {
        dfma<8><<<1, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("8 independent chains, %4d threads on one SM: %.3f DFMA warp-instructions per cycle per SM (%.1f lanes/clk)\n",
               threads, (double)iters * 8 * (threads / 32) / h, (double)iters * 8 * threads / h);
    }
    dfma<1><<<1, 32>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("one dependent chain, one warp: %.2f cycles per DFMA\n", (double)h / iters);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
