// Microbenchmark: tcgen05.ld / tcgen05.st throughput when TMEM is used as per-thread scratch (32x32b shapes).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/ubench/tmem_bw scripts/ubench/tmem_bw.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(addr));
}
__device__ __forceinline__ void tmem_st32(uint32_t addr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

template <int MODE>  // 0: ld only, 1: st only, 2: ld + add + st
__global__ void bench(int iters, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t base_addr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"l"(
            (uint64_t)__cvta_generic_to_shared(&base_addr)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = base_addr;
    // lane quarter = warp % 4 (hardware rule); column block = (warp / 4) * 64
    const uint32_t addr = base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 64u;
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = threadIdx.x * 32 + i;
    tmem_st32(addr, r);
    tmem_st32(addr + 32, r);
    asm volatile("tcgen05.wait::st.sync.aligned;");
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
            tmem_ld32(addr + (it & 1) * 32, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;");
        }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] += 1;
        }
        if (MODE == 1 || MODE == 2) {
            tmem_st32(addr + (it & 1) * 32, r);
            if (MODE == 1) asm volatile("tcgen05.wait::st.sync.aligned;");
        }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += r[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    // read back check
    tmem_ld32(addr, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    if (MODE == 2 && r[3] != threadIdx.x * 32 + 3 + (uint32_t)((iters + 1) / 2)) sink[0] = 0xdeadbeef;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
}

int main() {
    long long* d_cycles;
    uint32_t* d_sink;
    cudaMalloc(&d_cycles, 148 * sizeof(long long));
    cudaMalloc(&d_sink, 148 * 1024 * sizeof(uint32_t));
    const int iters = 4096;
    for (int warps : {1, 4, 8, 16}) {
        for (int mode = 0; mode < 3; ++mode) {
            cudaMemset(d_sink, 0, 148 * 1024 * 4);
            if (mode == 0) bench<0><<<148, warps * 32>>>(iters, d_cycles, d_sink);
            if (mode == 1) bench<1><<<148, warps * 32>>>(iters, d_cycles, d_sink);
            if (mode == 2) bench<2><<<148, warps * 32>>>(iters, d_cycles, d_sink);
            cudaError_t e = cudaDeviceSynchronize();
            long long c = 0;
            uint32_t s0 = 0;
            cudaMemcpy(&c, d_cycles, sizeof c, cudaMemcpyDeviceToHost);
            cudaMemcpy(&s0, d_sink, 4, cudaMemcpyDeviceToHost);
            const double bytes = (double)iters * warps * 32 * 32 * 4 * (mode == 2 ? 2 : 1);
            printf("warps %2d mode %s: %s cycles %lld -> %.1f B/clk/SM%s\n", warps,
                   mode == 0 ? "ld   " : mode == 1 ? "st   " : "ld+st", cudaGetErrorString(e), c, bytes / (double)c,
                   s0 == 0xdeadbeef ? "  READBACK MISMATCH" : "");
        }
    }
    return 0;
}
