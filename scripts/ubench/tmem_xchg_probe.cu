// tmem_xchg_probe.cu — can tensor memory carry a transform exchange?  A warp writes its lanes with tcgen05.st.32x32b
// (thread t -> lane t, register j -> column j) and reads them back with tcgen05.ld.16x256b, whose fragment layout
// hands thread t = c + 4 q the 64-bit units c of lanes q and q + 8 (and, from a second load at lane offset 16, of
// lanes q + 16 and q + 24): a 4 x 4 transposition between two lane bits and two unit-index bits that never touches
// the LSU / L1 data pipe ("a transform with fewer exchange bytes").  The expected pattern below is the
// one CuTe's copy traits state for SM100_TMEM_LOAD_16dp256b2x; this program prints what the hardware does and the
// cycles one st + ld round trip of 16 columns costs a warp.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tmem_xchg_probe_bin tmem_xchg_probe.cu && ./tmem_xchg_probe_bin
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128, 1) probe(uint32_t* __restrict__ out, long long* __restrict__ cycles) {
    __shared__ uint32_t tm_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"l"(
            (uint64_t)__cvta_generic_to_shared(&tm_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tm_base_s + ((uint32_t)(warp * 32) << 16);

    uint32_t w[16], r[16];
    for (int j = 0; j < 16; ++j) w[j] = (uint32_t)(threadIdx.x * 100 + j);  // lane * 100 + column
    long long best = 1 << 30;
    for (int rep = 0; rep < 8; ++rep) {
        const long long t0 = clock64();
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(tm),
            "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]),
            "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15])
            : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        // lanes 0-15 of the warp's quarter, columns 0-15 -> registers 0-7; lanes 16-31 -> registers 8-15
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tm));
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tm + (16u << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const long long t1 = clock64();
        if (t1 - t0 < best) best = t1 - t0;
    }
    for (int j = 0; j < 16; ++j) out[threadIdx.x * 16 + j] = r[j];
    if (lane == 0) cycles[warp] = best;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tm_base_s));
}

int main() {
    uint32_t *d_out, o[128 * 16];
    long long *d_cyc, cyc[4];
    cudaMalloc(&d_out, sizeof o);
    cudaMalloc(&d_cyc, sizeof cyc);
    cudaMemset(d_out, 0xFF, sizeof o);
    probe<<<1, 128>>>(d_out, d_cyc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    cudaMemcpy(o, d_out, sizeof o, cudaMemcpyDeviceToHost);
    cudaMemcpy(cyc, d_cyc, sizeof cyc, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int t = 0; t < 128; ++t) {
        const int base = (t / 32) * 32, c = t % 4, q = (t % 32) / 4;
        // expected (CuTe SM100_TMEM_LOAD_16dp256b2x): registers {0,1} lane q cols 2c,2c+1; {2,3} lane q+8 same cols;
        // {4,5} lane q cols 8+2c, 9+2c; {6,7} lane q+8 cols 8+2c, 9+2c; second load: the same with lanes + 16
        for (int j = 0; j < 16; ++j) {
            const int ld = j / 8, jj = j % 8;
            const int src_lane = base + q + 8 * ((jj / 2) % 2) + 16 * ld;
            const int col = 8 * (jj / 4) + 2 * c + (jj % 2);
            const uint32_t want = (uint32_t)(src_lane * 100 + col);
            if (o[t * 16 + j] != want) {
                if (bad < 16) printf("thread %d reg %d: got %u, expected %u\n", t, j, o[t * 16 + j], want);
                ++bad;
            }
        }
    }
    for (int t : {0, 1, 4, 5, 31, 32, 37}) {
        printf("thread %3d:", t);
        for (int j = 0; j < 16; ++j) printf(" %u", o[t * 16 + j]);
        printf("\n");
    }
    printf("mismatches against the CuTe layout: %d\n", bad);
    printf("st.x16 + 2 x ld.16x256b.x2 round trip, cycles per warp: %lld %lld %lld %lld\n", cyc[0], cyc[1], cyc[2], cyc[3]);
    return e != cudaSuccess || bad != 0;
}
