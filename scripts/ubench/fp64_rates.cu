// Microbenchmark: throughput of DADD / DMUL / DFMA and mixes on sm_100a, 8 independent chains per thread.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/ubench/fp64_rates_bin scripts/ubench/fp64_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double* out, int iters, double m, double c) {
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-9 + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) a[j] = fma(a[j], m, c);
            if (MODE == 1) a[j] = a[j] + c;
            if (MODE == 2) a[j] = a[j] * m;
            if (MODE == 3) { if (j & 1) a[j] = a[j] + c; else a[j] = fma(a[j], m, c); }   // half DADD half DFMA
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(int warps_per_sm, double* d) {
    const int threads = 256, blocks = 148 * warps_per_sm / 8, iters = 1 << 14;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, threads>>>(d, iters, 1.0000001, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = ms < best ? ms : best;
    }
    return 8.0 * iters * threads * blocks / (best * 1e-3) / 1e12;  // T instr/s (thread instructions)
}

int main() {
    double* d; cudaMalloc(&d, 148 * 64 * 256 * sizeof(double));
    const char* names[4] = {"DFMA", "DADD", "DMUL", "DADD+DFMA"};
    for (int w : {8, 16, 32, 64}) {
        double r[4] = {run<0>(w, d), run<1>(w, d), run<2>(w, d), run<3>(w, d)};
        for (int mth = 0; mth < 4; ++mth)
            printf("warps/SM %2d %-9s %.2f T thread-instr/s = %.1f lanes/clk/SM\n", w, names[mth], r[mth],
                   r[mth] * 1e12 / 148 / 1.965e9);
    }
    return 0;
}
