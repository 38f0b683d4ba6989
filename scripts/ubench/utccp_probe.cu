// utccp_probe.cu — which smem bytes land in which tensor-memory lane for tcgen05.cp.64x128b.warpx2 (both flavours),
// fed by a cp.async.bulk (TMA 1-D) load.  Prints, per thread, the words it reads back with tcgen05.ld.32x32b.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o utccp_probe_bin utccp_probe.cu && ./utccp_probe_bin
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128, 1) probe(const uint32_t* __restrict__ src, uint32_t* __restrict__ out) {
    __shared__ __align__(1024) uint32_t tile[2][256];  // 2 x (64 rows x 16 B)
    __shared__ __align__(8) uint64_t bar_tma, bar_cp;
    __shared__ uint32_t tm_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"l"(
            (uint64_t)__cvta_generic_to_shared(&tm_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    const uint32_t b_tma = (uint32_t)__cvta_generic_to_shared(&bar_tma), b_cp = (uint32_t)__cvta_generic_to_shared(&bar_cp);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b_tma));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b_cp));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tm_base_s;
    if (threadIdx.x == 0) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tile[0][0]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b_tma), "r"(2048u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(2048u), "r"(b_tma) : "memory");
        uint32_t done;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(b_tma), "r"(0u) : "memory");
        } while (!done);
        // matrix descriptor: start address >> 4, LBO 0, SBO = 128 B (8 rows x 16 B atoms, contiguous), version 1, no swizzle
        const uint64_t desc0 = (uint64_t)((dst & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);
        const uint64_t desc1 = (uint64_t)(((dst + 1024u) & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);
        asm volatile("tcgen05.cp.cta_group::1.64x128b.warpx2::02_13 [%0], %1;" ::"r"(tm + 0u), "l"(desc0) : "memory");
        asm volatile("tcgen05.cp.cta_group::1.64x128b.warpx2::01_23 [%0], %1;" ::"r"(tm + 4u), "l"(desc1) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b_cp) : "memory");
    }
    {
        uint32_t done;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(b_cp), "r"(0u) : "memory");
        } while (!done);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tm + ((uint32_t)(warp * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out[threadIdx.x * 8 + j] = r[j];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tm));
}

int main() {
    uint32_t h[512], *d_src, *d_out, o[1024];
    for (int i = 0; i < 512; ++i) h[i] = (i < 256 ? 0 : 100000) + (i % 256 / 4) * 10 + (i % 4);  // row*10 + word
    cudaMalloc(&d_src, sizeof h);
    cudaMalloc(&d_out, sizeof o);
    cudaMemcpy(d_src, h, sizeof h, cudaMemcpyHostToDevice);
    cudaMemset(d_out, 0xFF, sizeof o);
    probe<<<1, 128>>>(d_src, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    cudaMemcpy(o, d_out, sizeof o, cudaMemcpyDeviceToHost);
    for (int t : {0, 1, 31, 32, 33, 63, 64, 65, 95, 96, 97, 127}) {
        printf("thread %3d (quarter %d lane %2d): 02_13 ->", t, t / 32, t % 32);
        for (int j = 0; j < 4; ++j) printf(" %u", o[t * 8 + j]);
        printf("   01_23 ->");
        for (int j = 4; j < 8; ++j) printf(" %u", o[t * 8 + j]);
        printf("\n");
    }
    return e != cudaSuccess;
}
