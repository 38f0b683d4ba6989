// peba1_cuda.cu — libpeba1_cuda: the sm_100a gate-bootstrapping engine behind include/peba1_cuda.h.
//
// Kernels:
//   K0+K1  gate prologue, mod-switch, 630-step blind rotation, sample extraction (replaces tfhe_bootstrap_woKS_FFT and
//          everything below it):
//            pbs_kernel_tm       THE throughput kernel: one team of 64 threads per bootstrap, accumulators / twiddles /
//                                decomposition input / own accumulator words in tensor memory, persistent CTAs on two
//                                alternating exchange buffers (TeamSmemTm2, blind_rotate_tm.cuh); also the narrow shapes
//                                with 1..8 teams per CTA (TeamSmemTm)
//            pbs_kernel_wide     the latency kernel for at most one bootstrap per SM: one CTA of 2l teams per bootstrap,
//                                rows in parallel, key rows by TMA (blind_rotate_wide.cuh)
//            pbs_kernel, pbs_kernel_tmsh, pbs_kernel_pair, pbs_kernel_tmk: measured alternatives (kernel_variant 0, 2, 3, 4)
//   K2     keyswitch_lut_kernel  lweKeySwitch over the extracted samples through a device-built table of key-row sums
//                                (ksk_lut_build_kernel: one fetch per group of 4 digits), writes the ciphertext pool;
//                                narrow levels: ks_init_kernel + the same kernel with a ciphertext split over up to 64 CTAs
//          keyswitch_kernel / keyswitch_split_kernel: one key row per non-zero digit (ks_lut = 0), the form it replaced
//   K3     key_transform_kernel  bootstrapping key -> transform domain, once per key
//   K4     linear_kernel         bootsNOT / bootsCOPY / bootsCONSTANT on pool slots
//          lwe_encrypt_kernel    device-side encryption for synthetic inputs (peba1_cuda_encrypt_bits)
// There is no CPU fallback in this file: every entry point needs a CUDA device.
#include "../../include/peba1_cuda.h"

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "blind_rotate.cuh"
#include "blind_rotate_tm.cuh"
#ifdef PEBA1_ALL_VARIANTS  // measured alternatives that lost: `make VARIANTS=all`
#include "blind_rotate_tm_variants.cuh"
#include "blind_rotate_pair.cuh"
#include "blind_rotate_tmk.cuh"
#endif
#include "blind_rotate_wide.cuh"
#include "twiddles_host.h"

using namespace peba1;

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return fail(PEBA1_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                           __FILE__, __LINE__);                                        \
    } while (0)

extern "C" const char* peba1_cuda_last_error(void) { return g_err; }

// ------------------------------------------------------------------------------------------------ device side
namespace {

constexpr int kPoolPad = 4;        // pool stride is (n+1) rounded up to a multiple of 4 words (16-byte rows)
constexpr int kUStride = kN + 4;   // extracted sample: 1025 words, padded to 1028

struct DevEnv {
    int tid;      // thread index inside the team
    int bar_id;   // named barrier owned by the team (1..15)
    __device__ __forceinline__ void team_sync() const { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kTeam) : "memory"); }
    __device__ __forceinline__ void warp_sync() const { __syncwarp(); }
    __device__ __forceinline__ int32_t ldg(const int32_t* p) const { return __ldg(p); }
    __device__ __forceinline__ double2 ldg2(const double2* p) const { return __ldg(p); }
};

// tcgen05 32x32b load/store wrappers (tensor memory used as per-thread scratch; see blind_rotate_tm.cuh)
#define R8(a, o) "%" #o
__device__ __forceinline__ void tmem_ld_x8(uint32_t addr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(addr));
}
__device__ __forceinline__ void tmem_st_x8(uint32_t addr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void tmem_st_x4(uint32_t addr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                 ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t addr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr));
}
__device__ __forceinline__ void tmem_st_x16(uint32_t addr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t addr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(addr));
}
__device__ __forceinline__ void tmem_st_x32(uint32_t addr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
#undef R8

struct DevEnvTm {
    int tid;
    int bar_id;
    int sums_col;
    uint32_t tm_base;  // allocation base + this warp's lane quarter
    int chunk_id = 0;  // shared-key variant: quarter of each key row this team fetches
    uint32_t group_mbar = 0;  // shared-key variant: smem address of the group's two mbarriers (one per row buffer)
    uint32_t ex1_mbar = 0;    // smem address of the team's "exchange buffer free" mbarrier
    uint32_t ex1_phase = 1;   // parity to wait for: the buffer starts free (the phase before the first counts as done)
    int tv_col = 0;           // 16 tensor-memory columns where the decomposition input waits between digit levels
    int own_col = 0;          // 16 tensor-memory columns with this thread's own words of the body accumulator
    int xpar = 0;             // which exchange buffer the next transform uses (double-buffered layouts)
    // ---- kXTm layouts (TeamSmemTmX): the cross-team exchange through tensor memory
    // (all derived from threadIdx on the spot: as fields they would sit in six registers for the whole kernel)
    __device__ __forceinline__ int xgrp() const { return (int)(threadIdx.x >> 7); }         // warp / 4: the column block
    __device__ __forceinline__ int xwb() const { return (int)(threadIdx.x >> 7) & 1; }      // tid = 32 wb + lane
    __device__ __forceinline__ int x_col() const { return TeamSmemTmX::kX + 32 * xgrp(); }  // this warp's 32 exchange columns
    __device__ __forceinline__ int x_peer() const { return TeamSmemTmX::kX + 32 * (xgrp() ^ 1); }  // the team partner's
    __device__ __forceinline__ int tw1_col() const { return TeamSmemTmX::kTw1 + 32 * xwb(); }
    // logical position after the exchange: hi = 4 wb + (lane & 3), lo = lane >> 2
    __device__ __forceinline__ int xhi() const { return 4 * xwb() + (int)(threadIdx.x & 3); }
    __device__ __forceinline__ int xlo() const { return (int)(threadIdx.x & 31) >> 2; }
    // (Write-after-read protection of the blocks by a split-phase mbarrier instead of the first team barrier of a trip was
    // measured: 155.7 ms instead of 147.6 — the polling costs more issue slots than the rendezvous, as with the "buffer
    // free" mbarrier of TeamSmemTm.)
    // forward: 32x32b store of the own block, 16x256b loads of columns [16 wb, +16) of both blocks (S = 0, 1: the block
    // of the team's warp S) and both lane halves
    __device__ __forceinline__ void xtm_exchange_fwd(uint32_t (&v)[32]) const {
        team_sync();  // the partner has read what the previous trip left in the blocks
        tmem_st_x32(tm_base + (uint32_t)x_col(), v);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        team_sync();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int S = q >> 1, H = q & 1;
            const uint32_t a = tm_base + (uint32_t)((S == xwb() ? x_col() : x_peer()) + 16 * xwb()) + ((uint32_t)(16 * H) << 16);
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[8 * q]), "=r"(v[8 * q + 1]), "=r"(v[8 * q + 2]), "=r"(v[8 * q + 3]), "=r"(v[8 * q + 4]),
                           "=r"(v[8 * q + 5]), "=r"(v[8 * q + 6]), "=r"(v[8 * q + 7])
                         : "r"(a));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    // inverse: the same trip backwards — 16x256b stores into both blocks, 32x32b load of the own block
    __device__ __forceinline__ void xtm_exchange_inv(uint32_t (&v)[32], uint32_t (&tw)[32]) const {
        team_sync();  // the partner is done with the blocks' previous contents
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int S = q >> 1, H = q & 1;
            const uint32_t a = tm_base + (uint32_t)((S == xwb() ? x_col() : x_peer()) + 16 * xwb()) + ((uint32_t)(16 * H) << 16);
            asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         ::"r"(a), "r"(v[8 * q]), "r"(v[8 * q + 1]), "r"(v[8 * q + 2]), "r"(v[8 * q + 3]), "r"(v[8 * q + 4]),
                           "r"(v[8 * q + 5]), "r"(v[8 * q + 6]), "r"(v[8 * q + 7])
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        team_sync();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_ld_x32(tm_base + (uint32_t)x_col(), v);
        tmem_ld_x32(tm_base + (uint32_t)tw1_col(), tw);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    __device__ __forceinline__ int flip() { const int p = xpar; xpar ^= 1; return p; }
    // Split-phase write-after-read protection of the cross-team exchange buffer: a warp signals as soon as its reads
    // are done, the next writer waits just before writing, one whole transform pass later, so nobody actually blocks.
    __device__ __forceinline__ void ex1_release() const {
        __syncwarp();
        if ((threadIdx.x & 31) == 0)
            asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(ex1_mbar) : "memory");
    }
    __device__ __forceinline__ void ex1_acquire() {
        uint32_t done;
        do {
            asm volatile(
                "{\n.reg .pred p;\n"
                "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n"
                "selp.u32 %0, 1, 0, p;\n}"
                : "=r"(done) : "r"(ex1_mbar), "r"(ex1_phase) : "memory");
        } while (!done);
        ex1_phase ^= 1u;
    }
    // split-phase group barrier: one arrival per warp (8 warps per group)
    __device__ __forceinline__ void group_arrive(int buf) const {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if ((threadIdx.x & 31) == 0)
            asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(group_mbar + 8u * (uint32_t)buf) : "memory");
    }
    __device__ __forceinline__ void group_wait(int buf, int parity) const {
        const uint32_t addr = group_mbar + 8u * (uint32_t)buf;
        uint32_t done;
        do {
            asm volatile(
                "{\n.reg .pred p;\n"
                "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n"
                "selp.u32 %0, 1, 0, p;\n}"
                : "=r"(done) : "r"(addr), "r"((uint32_t)parity) : "memory");
        } while (!done);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    __device__ __forceinline__ void team_sync() const { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kTeam) : "memory"); }
    __device__ __forceinline__ void warp_sync() const { __syncwarp(); }
    // a complex double to / from the lane whose index differs in `mask` (lane8_transpose, blind_rotate_tm.cuh)
    __device__ __forceinline__ double2 shfl_xor_c(double2 v, int mask) const {
        return make_double2(__shfl_xor_sync(0xffffffffu, v.x, mask), __shfl_xor_sync(0xffffffffu, v.y, mask));
    }
    __device__ __forceinline__ int32_t ldg(const int32_t* p) const { return __ldg(p); }
    __device__ __forceinline__ double2 ldg2(const double2* p) const { return __ldg(p); }
    __device__ __forceinline__ void tmem_ld4(int col, uint32_t (&w)[4]) const { tmem_ld_x4(tm_base + (uint32_t)col, w); }
    __device__ __forceinline__ void tmem_st4(int col, const uint32_t (&w)[4]) const { tmem_st_x4(tm_base + (uint32_t)col, w); }
    __device__ __forceinline__ void tmem_ld8(int col, uint32_t (&w)[8]) const { tmem_ld_x8(tm_base + (uint32_t)col, w); }
    __device__ __forceinline__ void tmem_st8(int col, const uint32_t (&w)[8]) const { tmem_st_x8(tm_base + (uint32_t)col, w); }
    __device__ __forceinline__ void tmem_ld16(int col, uint32_t (&w)[16]) const { tmem_ld_x16(tm_base + (uint32_t)col, w); }
    __device__ __forceinline__ void tmem_st16(int col, const uint32_t (&w)[16]) const { tmem_st_x16(tm_base + (uint32_t)col, w); }
    __device__ __forceinline__ void tmem_ld32(int col, uint32_t (&w)[32]) const { tmem_ld_x32(tm_base + (uint32_t)col, w); }
    __device__ __forceinline__ void tmem_st32(int col, const uint32_t (&w)[32]) const { tmem_st_x32(tm_base + (uint32_t)col, w); }
    __device__ __forceinline__ void tmem_wait_ld() const { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
    __device__ __forceinline__ void tmem_wait_st() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
};

#ifdef PEBA1_ALL_VARIANTS
// Key pipeline producer (see blind_rotate_tmk.cuh): one thread of a dedicated 17th warp.
// Barriers (8 bytes each, from `bars`): tma_full[3] @0, smem_free[3] @24, tmem_full[4] @48, tmem_free[4] @80.
__device__ __forceinline__ void mbar_wait_parity(uint32_t addr, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
// The producer thread shares its scheduler with four compute warps: it must not spin at full rate (measured: 214 M
// failed polls per launch = a fifth of that scheduler's issue slots, and its four warps become the CTA's laggards).
__device__ __forceinline__ void mbar_wait_parity_sleep(uint32_t addr, uint32_t parity) {
    uint32_t done;
    for (;;) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(64);
    }
}
__device__ __forceinline__ void key_issue_tma(uint32_t bars, uint32_t stage, const unsigned char* key, int g) {
    const int slot = g % 3, use = g / 3;
    if (use >= 1) mbar_wait_parity_sleep(bars + 24u + 8u * (uint32_t)slot, (uint32_t)(use - 1) & 1u);  // smem_free
    const uint32_t dst = stage + (uint32_t)slot * (uint32_t)kKeyHalfBytes;
    const uint32_t full = bars + 8u * (uint32_t)slot;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"((uint32_t)kKeyHalfBytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(key + (size_t)g * kKeyHalfBytes), "r"((uint32_t)kKeyHalfBytes), "r"(full) : "memory");
}
__device__ __noinline__ void key_producer_loop(uint32_t tm_alloc, uint32_t bars, uint32_t stage, const unsigned char* key,
                                               int total_half) {
    for (int g = 0; g < 2 && g < total_half; ++g) key_issue_tma(bars, stage, key, g);
    for (int h = 0; h < total_half; ++h) {
        const int slot = h % 3, use = h / 3;
        mbar_wait_parity_sleep(bars + 8u * (uint32_t)slot, (uint32_t)use & 1u);                                     // tma_full
        if (h >= 4) mbar_wait_parity_sleep(bars + 80u + 8u * (uint32_t)(h & 3), (uint32_t)((h >> 2) - 1) & 1u);      // tmem_free
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t src = stage + (uint32_t)slot * (uint32_t)kKeyHalfBytes;
        const uint32_t dst = tm_alloc + (uint32_t)(kTmKey + 32 * (h & 3));
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // block j = 2 * (k2 & 3) + c: 64 thread indices x 16 bytes
            // matrix descriptor: start >> 4, LBO 0, SBO 128 B (8-row x 16-byte atoms, contiguous), version 1, no swizzle
            const uint64_t desc = (uint64_t)(((src + 1024u * (uint32_t)j) & 0x3FFFFu) >> 4) |
                                  ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);
            asm volatile("tcgen05.cp.cta_group::1.64x128b.warpx2::02_13 [%0], %1;"
                         ::"r"(dst + (uint32_t)(16 * (j & 1) + 4 * (j >> 1))), "l"(desc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                     ::"r"(bars + 48u + 8u * (uint32_t)(h & 3)) : "memory");   // tmem_full
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                     ::"r"(bars + 24u + 8u * (uint32_t)slot) : "memory");      // smem_free
        // refill the slot half-row h-1 was copied from (its copies were issued one iteration ago: done by now)
        if (h + 2 < total_half) key_issue_tma(bars, stage, key, h + 2);
    }
}

// Environment of the key-through-tensor-memory kernel (blind_rotate_tmk.cuh): DevEnvTm + the consumer side of the
// key pipeline.
struct DevEnvTmk : DevEnvTm {
    uint32_t bars = 0;  // smem address of the 14 mbarriers
    __device__ __forceinline__ void key_pump(int) const {}
    __device__ __forceinline__ void key_wait(int h) const {
        mbar_wait_parity(bars + 48u + 8u * (uint32_t)(h & 3), (uint32_t)(h >> 2) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    __device__ __forceinline__ void key_ld16(int h, int c, uint32_t (&k)[16]) const {
        tmem_ld_x16(tm_base + (uint32_t)(kTmKey + 32 * (h & 3) + 16 * c), k);
    }
    __device__ __forceinline__ void key_release(int h) const {  // after the tcgen05.wait::ld that covers the loads
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if ((threadIdx.x & 31) == 0)
            asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bars + 80u + 8u * (uint32_t)(h & 3)) : "memory");
    }
};

#endif  // PEBA1_ALL_VARIANTS
__constant__ TwiddleTables c_tables;

// The per-thread twiddles are read ONCE into registers.  They must not be rematerialised from constant memory
// inside the iteration loop: a thread-indexed LDC serialises over the warp (ncu r1b: 60% of all stall samples sat on
// the first DMUL after such loads), so every value is laundered through an empty asm that hides its origin.
__device__ __forceinline__ void load_twiddles_dev(int tid, TeamTwiddles& tw) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        tw.tw1[j] = c_tables.tw1[tid][j];
        tw.tw2[j] = c_tables.tw2[tid & 7][j];
        asm volatile("" : "+d"(tw.tw1[j].x), "+d"(tw.tw1[j].y), "+d"(tw.tw2[j].x), "+d"(tw.tw2[j].y));
    }
}

// One keyswitch job: out_slot <- KS( (0,c0) + u[u0] + u[u1] ), u1 < 0: absent (non-MUX gates).
struct __align__(16) KsJob {
    int32_t u0, u1;
    int32_t c0;
    int32_t pad;
    int64_t out_slot;
};

#ifdef PEBA1_ALL_VARIANTS
// K0+K1.  One team of 64 threads per bootstrap, TEAMS teams per CTA; job index = blockIdx.x*TEAMS + team.
template <int L, int BGBIT, int TEAMS>
__global__ void __launch_bounds__(TEAMS* kTeam, 1)
    pbs_kernel(const double2* __restrict__ bkfft, int n, const int32_t* __restrict__ pool, int64_t pool_stride,
               const PbsJob* __restrict__ jobs, int64_t count, int32_t* __restrict__ u_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TeamSmem* teams = reinterpret_cast<TeamSmem*>(smem_raw);
    const int team = threadIdx.x / kTeam;
    const int64_t g = (int64_t)blockIdx.x * TEAMS + team;
    if (g >= count) return;  // whole team leaves together; named barriers are per team
    DevEnv env{(int)(threadIdx.x % kTeam), team + 1};
    TeamTwiddles tw;
    load_twiddles_dev(env.tid, tw);
    blind_rotate_team<L, BGBIT>(env, teams[team], tw, bkfft, n, pool, pool_stride, jobs[g],
                                u_out + g * (int64_t)kUStride);
}

#endif  // PEBA1_ALL_VARIANTS
// K0+K1, tensor-memory version: 8 teams (512 threads, 16 warps) per CTA, one CTA per SM, <= 128 registers.
constexpr int kTeamsTm = 8;
constexpr int kTeamsTm10 = 10;  // (VARIANTS=all) with the single-exchange-buffer layout (TeamSmemTm1): 20 warps, 96 registers
template <int L, int BGBIT, int TEAMS, class SM = TeamSmemTm, bool REGACC = false>
__global__ void __launch_bounds__(TEAMS* kTeam, 1)
    pbs_kernel_tm(const double2* __restrict__ bkfft, int n, const int32_t* __restrict__ pool, int64_t pool_stride,
                  const PbsJob* __restrict__ jobs, int64_t count, int32_t* __restrict__ u_out,
                  unsigned long long* __restrict__ job_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM* teams = reinterpret_cast<SM*>(smem_raw);
    __shared__ long long next_job[TEAMS];
    __shared__ uint32_t tm_alloc_base;
    const int warp = threadIdx.x >> 5;
    // A team is two warps.  Ordinarily warps (2t, 2t + 1); with the tensor-memory exchange the two must share a lane quarter
    // (warp % 4), so team t = (warp & 3) + 4 (warp >> 3) is warps (w, w + 4) and wb = (warp >> 2) & 1 tells them apart.
    static_assert(!SM::kXTm || TEAMS == 8, "the tensor-memory exchange pairs warps (w, w + 4) of a 16-warp CTA");
#if defined(PEBA1_ABLATE) && PEBA1_ABLATE == 5  /* what-if: the tensor-memory exchange code with teams of ADJACENT warps (wrong data) */
    const int wb = warp & 1;
    const int team = (int)(threadIdx.x / kTeam);
#else
    const int wb = SM::kXTm ? (warp >> 2) & 1 : warp & 1;
    const int team = SM::kXTm ? (warp & 3) + 4 * (warp >> 3) : (int)(threadIdx.x / kTeam);
#endif
    const int team_tid = 32 * wb + (int)(threadIdx.x & 31);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"l"(
            (uint64_t)__cvta_generic_to_shared(&tm_alloc_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (team_tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" ::"r"((uint32_t)__cvta_generic_to_shared(&teams[team].ex1_free)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    DevEnvTm env{team_tid, team + 1, kTmSums + 64 * (warp >> 2),
                 tm_alloc_base + ((uint32_t)(warp & 3) * 32u << 16)};
    env.ex1_mbar = (uint32_t)__cvta_generic_to_shared(&teams[team].ex1_free);
    env.tv_col = SM::kTv + 16 * (warp >> 2);
    env.own_col = SM::kOwn + 32 * (warp >> 2);
    if constexpr (SM::kXTm) {
        const int lane = (int)(threadIdx.x & 31);
        // twiddles: tw1 belongs to tid = 32 wb + lane (one set per wb: warps 0..7 publish), tw2 to lo = lane >> 2 (warps 0..3)
        if (warp < 8) {
            double2 t1[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t1[j] = c_tables.tw1[team_tid][j];
            tm_store8(env, env.tw1_col(), t1);
        }
        if (warp < 4) {
            double2 t2[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t2[j] = c_tables.tw2[lane >> 2][j];
            tm_store8(env, SM::kTw2, t2);
        }
        env.tmem_wait_st();
    } else if ((warp >> 2) == 0) {  // one warp per lane quarter publishes the per-thread twiddles
        TeamTwiddles tw;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            tw.tw1[j] = c_tables.tw1[env.tid][j];
            tw.tw2[j] = c_tables.tw2[env.tid & 7][j];
        }
        tm_store_twiddles<SM>(env, tw);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (job_counter == nullptr) {  // one job per team, fixed by the grid
        // (the narrow shapes launch this same instantiation with fewer teams per CTA: blockDim.x / 64 of them)
        const int64_t g = (int64_t)blockIdx.x * (int64_t)(blockDim.x / kTeam) + team;
        if (g < count)
            blind_rotate_team_tm<L, BGBIT>(env, teams[team], bkfft, n, pool, pool_stride, jobs[g],
                                           u_out + g * (int64_t)kUStride);
    } else {
        // Persistent CTAs (one per SM): every team draws its next job from a global counter.  The bootstraps all take
        // the same time, so this only matters at the end of a batch: instead of a last wave of whole CTAs on a few
        // SMs, the remaining jobs spread over all SMs, whose teams run faster the fewer of them are left.
        for (;;) {
            if (env.tid == 0) next_job[team] = (long long)atomicAdd(job_counter, 1ULL);
            env.team_sync();
            const int64_t g = next_job[team];
            env.team_sync();
            if (g >= count) break;
#ifdef PEBA1_ALL_VARIANTS
            if constexpr (REGACC)
                blind_rotate_team_tmr<L, BGBIT>(env, teams[team], bkfft, n, pool, pool_stride, jobs[g],
                                                u_out + g * (int64_t)kUStride);
            else
#endif
                blind_rotate_team_tm<L, BGBIT>(env, teams[team], bkfft, n, pool, pool_stride, jobs[g],
                                               u_out + g * (int64_t)kUStride);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_alloc_base));
}

#ifdef PEBA1_ALL_VARIANTS
// K0+K1, shared-key tensor-memory version (blind_rotate_team_tmsh): as pbs_kernel_tm, plus the even teams and the odd
// teams each form a group that fetches every key row once and shares it through tensor memory.
template <int L, int BGBIT>
__global__ void __launch_bounds__(kTeamsTm* kTeam, 1)
    pbs_kernel_tmsh(const double2* __restrict__ bkfft, int n, const int32_t* __restrict__ pool, int64_t pool_stride,
                    const PbsJob* __restrict__ jobs, int64_t count, int32_t* __restrict__ u_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TeamSmemTm* teams = reinterpret_cast<TeamSmemTm*>(smem_raw);
    __shared__ uint32_t tm_alloc_base;
    __shared__ __align__(8) uint64_t group_mbars[2][2];  // [group][row buffer]
    const int warp = threadIdx.x >> 5;
    const int team = threadIdx.x / kTeam;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"l"(
            (uint64_t)__cvta_generic_to_shared(&tm_alloc_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if ((threadIdx.x % kTeam) == 0)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" ::"r"((uint32_t)__cvta_generic_to_shared(&teams[team].ex1_free)));
    if (threadIdx.x == 32) {
        for (int q = 0; q < 4; ++q)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(
                (uint32_t)__cvta_generic_to_shared(&group_mbars[q >> 1][q & 1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    DevEnvTm env{(int)(threadIdx.x % kTeam), team + 1, kTmSums + 64 * (warp >> 2),
                 tm_alloc_base + ((uint32_t)(warp & 3) * 32u << 16), team >> 1,
                 (uint32_t)__cvta_generic_to_shared(&group_mbars[team & 1][0])};
    env.ex1_mbar = (uint32_t)__cvta_generic_to_shared(&teams[team].ex1_free);
    if ((warp >> 2) == 0) {
        TeamTwiddles tw;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            tw.tw1[j] = c_tables.tw1[env.tid][j];
            tw.tw2[j] = c_tables.tw2[env.tid & 7][j];
        }
        tm_store_twiddles(env, tw);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    int64_t g = (int64_t)blockIdx.x * kTeamsTm + team;
    const bool live = g < count;
    if (!live) g = count - 1;  // a team without a job shadows the last one: it must still feed its group
    blind_rotate_team_tmsh<L, BGBIT>(env, teams[team], bkfft, n, pool, pool_stride, jobs[g], live,
                                     u_out + g * (int64_t)kUStride);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_alloc_base));
}

// K0+K1, pair version (blind_rotate_pair.cuh): TEAMS teams of 64 threads, TWO bootstraps per team; 4 teams = 8 warps
// per SM with up to 255 registers.  Job index = 2 * (blockIdx.x * TEAMS + team) (+1).
constexpr int kTeamsPair = 4;
template <int L, int BGBIT, int TEAMS>
__global__ void __launch_bounds__(TEAMS* kTeam, 1)
    pbs_kernel_pair(const double2* __restrict__ bkfft, int n, const int32_t* __restrict__ pool, int64_t pool_stride,
                    const PbsJob* __restrict__ jobs, int64_t count, int32_t* __restrict__ u_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TeamSmemPair* teams = reinterpret_cast<TeamSmemPair*>(smem_raw);
    __shared__ uint32_t tm_alloc_base;
    const int warp = threadIdx.x >> 5;
    const int team = threadIdx.x / kTeam;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"l"(
            (uint64_t)__cvta_generic_to_shared(&tm_alloc_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if ((threadIdx.x % kTeam) == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" ::"r"((uint32_t)__cvta_generic_to_shared(&teams[team].ex1_free)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    DevEnvTm env{(int)(threadIdx.x % kTeam), team + 1, kTmSums + 128 * (warp >> 2),
                 tm_alloc_base + ((uint32_t)(warp & 3) * 32u << 16)};
    env.ex1_mbar = (uint32_t)__cvta_generic_to_shared(&teams[team].ex1_free);
    // one warp per lane quarter publishes the per-thread twiddles: with fewer than 4 warps (TEAMS == 1) quarters 2 and 3
    // are never touched
    if ((warp >> 2) == 0) {
        TeamTwiddles tw;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            tw.tw1[j] = c_tables.tw1[env.tid][j];
            tw.tw2[j] = c_tables.tw2[env.tid & 7][j];
        }
        tm_store_twiddles(env, tw);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int64_t g = 2 * ((int64_t)blockIdx.x * TEAMS + team);
    if (g < count) {
        const bool two = g + 1 < count;
        blind_rotate_pair_tm<L, BGBIT>(env, teams[team], bkfft, n, pool, pool_stride, jobs[g], jobs[two ? g + 1 : g],
                                       u_out + g * (int64_t)kUStride,
                                       two ? u_out + (g + 1) * (int64_t)kUStride : nullptr);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_alloc_base));
}

// K0+K1, key-through-tensor-memory version (blind_rotate_tmk.cuh): 8 teams (16 warps) + a 17th warp whose lane 0 streams
// every key row once per CTA with TMA + tcgen05.cp (544 threads => 120 registers).  Dynamic smem: [3 x 8 KB key
// staging][8 x TeamSmemTm].
constexpr int kTmkThreads = kTeamsTm * kTeam + 32;
template <int L, int BGBIT>
__global__ void __launch_bounds__(kTmkThreads, 1)
    pbs_kernel_tmk(const double2* __restrict__ bkfft, int n, const int32_t* __restrict__ pool, int64_t pool_stride,
                   const PbsJob* __restrict__ jobs, int64_t count, int32_t* __restrict__ u_out, int stagger_ns) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TeamSmemTm* teams = reinterpret_cast<TeamSmemTm*>(smem_raw + 3 * kKeyHalfBytes);
    __shared__ uint32_t tm_alloc_base;
    __shared__ __align__(8) uint64_t key_bars[14];
    const int warp = threadIdx.x >> 5;
    const int team = threadIdx.x / kTeam;  // 8 = the producer warp
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"l"(
            (uint64_t)__cvta_generic_to_shared(&tm_alloc_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (team < kTeamsTm && (threadIdx.x % kTeam) == 0)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" ::"r"((uint32_t)__cvta_generic_to_shared(&teams[team].ex1_free)));
    if (threadIdx.x == 32) {
        for (int b = 0; b < 14; ++b)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&key_bars[b])),
                         "r"(b >= 10 ? 2 * kTeamsTm : 1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int total_half = n * 2 * L * 2;
    if (team < kTeamsTm) {
        DevEnvTmk env;
        env.tid = (int)(threadIdx.x % kTeam);
        env.bar_id = team + 1;
        env.sums_col = kTmSums + 64 * (warp >> 2);
        env.tm_base = tm_alloc_base + ((uint32_t)(warp & 3) * 32u << 16);
        env.ex1_mbar = (uint32_t)__cvta_generic_to_shared(&teams[team].ex1_free);
        env.bars = (uint32_t)__cvta_generic_to_shared(&key_bars[0]);
        if ((warp >> 2) == 0) {
            TeamTwiddles tw;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                tw.tw1[j] = c_tables.tw1[env.tid][j];
                tw.tw2[j] = c_tables.tw2[env.tid & 7][j];
            }
            tm_store_twiddles(env, tw);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 15, %0;" ::"n"(kTeamsTm * kTeam) : "memory");  // the 16 compute warps: twiddles are published
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        int64_t g = (int64_t)blockIdx.x * kTeamsTm + team;
        const bool live = g < count;
        if (!live) g = count - 1;  // a team without a job shadows the last one: it must still consume the key rows
        // The shared key ring keeps the teams within two rows of each other; started together they stay in phase and
        // hit the FP64 pipe and the exchange buffers in bursts.  A start-up stagger spreads them over a row.
        if (stagger_ns > 0) __nanosleep((unsigned)(stagger_ns * team));
        blind_rotate_team_tmk<L, BGBIT>(env, teams[team], n, pool, pool_stride, jobs[g], live,
                                        u_out + g * (int64_t)kUStride);
    } else if ((threadIdx.x & 31) == 0) {
        key_producer_loop(tm_alloc_base, (uint32_t)__cvta_generic_to_shared(&key_bars[0]),
                          (uint32_t)__cvta_generic_to_shared(smem_raw), reinterpret_cast<const unsigned char*>(bkfft),
                          total_half);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_alloc_base));
}

#endif  // PEBA1_ALL_VARIANTS
// Device-side encryption: synthetic gallery / benchmark inputs generated where they are used
// instead of 2.5 KB per bit over PCIe.  One warp per ciphertext; counter-based generator (Philox4x32-10, keyed by the
// seed, counter = ciphertext index and word block), so the output depends only on (seed, index), not on the launch
// shape.  c = (a, <a,s> + e + mu), a uniform, e = round(N(0, sigma) * 2^32), mu = +-1/8: TFHE's bootsSymEncrypt
// (src/main.cpp:64).  TEST / BENCHMARK use: it needs the secret LWE key on the device.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void __launch_bounds__(256)
    lwe_encrypt_kernel(int32_t* __restrict__ pool, int64_t stride, const int64_t* __restrict__ slots, int64_t first_slot,
                       const int32_t* __restrict__ bits, int64_t count, int64_t index0, int n,
                       const uint32_t* __restrict__ keymask, uint64_t seed, double sigma) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= count) return;
    const int64_t slot = slots ? slots[g] : first_slot + g;
    int32_t* row = pool + slot * stride;
    const uint64_t idx = (uint64_t)(index0 + g);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t sum = 0;
    for (int q = lane; 4 * q < stride; q += 32) {
        uint32_t r[4];
        philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)q, 0u, k0, k1, r);
        const uint32_t km = keymask[q >> 3] >> (4 * (q & 7));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int w = 4 * q + j;
            if (w >= n) r[j] = 0u;                 // body (written below) and padding
            else if ((km >> j) & 1u) sum += r[j];
        }
        *reinterpret_cast<int4*>(row + 4 * q) = make_int4((int)r[0], (int)r[1], (int)r[2], (int)r[3]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    if (lane == 0) {
        uint32_t r[4];
        philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), 0xFFFFFFFFu, 1u, k0, k1, r);
        // Box-Muller on two 32-bit uniforms in (0,1]
        const double u1 = ((double)r[0] + 1.0) * (1.0 / 4294967296.0), u2 = (double)r[1] * (1.0 / 4294967296.0);
        const double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
        const double d = z * sigma;
        const uint32_t e = (uint32_t)(int32_t)(int64_t)((d - (double)(int64_t)d) * 4294967296.0);
        const uint32_t mu = bits[g] ? (1u << 29) : 0u - (1u << 29);
        row[n] = (int32_t)(sum + e + mu);
    }
}

// K0+K1, low-latency version (blind_rotate_wide.cuh): one CTA of 2l teams per bootstrap, for narrow circuit levels.
struct DevEnvWide : DevEnvTm {
    __device__ __forceinline__ void cta_sync() const { __syncthreads(); }
    __device__ __forceinline__ void cta_sync_tm() const {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // a team's 16 partial products <-> the 64 tensor-memory columns of its warp (block = warp / 4; thread = lane), moved in
    // pieces of 4 values: a 32- or 64-register staging vector next to the 64 twiddle registers makes ptxas rematerialise the
    // twiddles from constant memory inside the step loop (thread-indexed LDC: 6.9 ms instead of 2.5, ncu r2h)
    __device__ __forceinline__ void prod_store(const double2 (&p0)[8], const double2 (&p1)[8]) const {
        const uint32_t col = 64u * (uint32_t)(threadIdx.x >> 7);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            uint32_t w[16];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double2 v = h < 2 ? p0[4 * h + k] : p1[4 * (h - 2) + k];
                c_to_words(v, w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
            }
            tmem_st_x16(tm_base + col + 16u * (uint32_t)h, w);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    // p0, p1 += the products of team r (same parity as the caller's team: its warps sit in the caller's lane quarters)
    __device__ __forceinline__ void prod_accumulate(int r, double2 (&p0)[8], double2 (&p1)[8]) const {
        const uint32_t col = 64u * (uint32_t)((2 * r + (int)((threadIdx.x >> 5) & 1)) >> 2);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            uint32_t w[16];
            tmem_ld_x16(tm_base + col + 16u * (uint32_t)h, w);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double2 v = words_to_c(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
                if (h < 2) p0[4 * h + k] = cadd(p0[4 * h + k], v);
                else p1[4 * (h - 2) + k] = cadd(p1[4 * (h - 2) + k], v);
            }
        }
    }
    __device__ __forceinline__ void pair_sync() const { asm volatile("bar.sync 15, 128;" ::: "memory"); }
    __device__ __forceinline__ void key_fetch(void* dst, const void* src, unsigned bytes, unsigned long long* mbar) const {
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst), b = (uint32_t)__cvta_generic_to_shared(mbar);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
    }
    __device__ __forceinline__ void key_wait(unsigned long long* mbar, unsigned parity) const {
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(mbar);
        uint32_t done;
        do {
            asm volatile(
                "{\n.reg .pred p;\n"
                "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n"
                "selp.u32 %0, 1, 0, p;\n}"
                : "=r"(done) : "r"(b), "r"(parity) : "memory");
        } while (!done);
    }
};
template <int L, int BGBIT>
__global__ void __launch_bounds__(2 * L * kTeam, 1)
    pbs_kernel_wide(const double2* __restrict__ bkfft, int n, const int32_t* __restrict__ pool, int64_t pool_stride,
                    const PbsJob* __restrict__ jobs, int64_t count, int32_t* __restrict__ u_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WideSmem<2 * L>& sm = *reinterpret_cast<WideSmem<2 * L>*>(smem_raw);
    __shared__ uint32_t tm_alloc_base;
    const int team = threadIdx.x / kTeam;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {  // 64 columns per warp of a lane quarter, up to 3 warps per quarter: 256 allocated
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"l"(
            (uint64_t)__cvta_generic_to_shared(&tm_alloc_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    DevEnvWide env;
    env.tid = (int)(threadIdx.x % kTeam);
    env.bar_id = team + 1;
    env.tm_base = tm_alloc_base + ((uint32_t)(warp & 3) * 32u << 16);
    if (warp < 4) {  // one warp per lane quarter publishes the twiddles of tid = 32 (warp & 1) + lane
        double2 t1[8], t2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            t1[j] = c_tables.tw1[env.tid][j];
            t2[j] = c_tables.tw2[env.tid & 7][j];
        }
        tm_store8(env, kWideTw1, t1);
        tm_store8(env, kWideTw2, t2);
        env.tmem_wait_st();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int64_t g = blockIdx.x; g < count; g += gridDim.x) {
        if (env.tid == 0) {  // a fresh barrier per job: the parity of step i is i & 1
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&sm.key_full[team])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        blind_rotate_wide<L, BGBIT>(env, sm, team, bkfft, n, pool, pool_stride, jobs[g], u_out + g * (int64_t)kUStride);
        __syncthreads();  // the shared accumulator is reused by the CTA's next job
        if (env.tid == 0)  // the next job initialises the barrier again: a live mbarrier must be invalidated first
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&sm.key_full[team])) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tm_alloc_base));
}

// K3.  One team per key polynomial q = (i*rows + r)*2 + c.
template <int TEAMS>
__global__ void __launch_bounds__(TEAMS* kTeam, 1)
    key_transform_kernel(const int32_t* __restrict__ bk, int rows, int64_t polys, double2* __restrict__ bkfft) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TeamSmem* teams = reinterpret_cast<TeamSmem*>(smem_raw);
    const int team = threadIdx.x / kTeam;
    const int64_t q = (int64_t)blockIdx.x * TEAMS + team;
    if (q >= polys) return;
    DevEnv env{(int)(threadIdx.x % kTeam), team + 1};
    TeamTwiddles tw;
    load_twiddles_dev(env.tid, tw);
    const int c = (int)(q & 1);
    const int64_t ir = q >> 1;
    key_transform_team(env, teams[team], tw, bk + q * kN,
                       bkfft + bkfft_index(rows, (int)(ir / rows), (int)(ir % rows), 0, c, 0));
}

// K2.  Keyswitch.  KSK device layout: [i][j][d-1][stride] (the all-zero digit-0 rows are dropped).
// One CTA of 160 threads handles G ciphertexts; thread t owns output words 4t..4t+3 of ALL G of them (4G accumulator
// registers; the row stride is a multiple of 4 words).  For every digit position (i, j) each ciphertext with a
// non-zero digit streams one 2.5 KB key row (ld.global.nc.v4).  The G requests of one position are issued back to
// back by the same thread and can only name 3 different rows.  Measured (ncu r1e): this does NOT help — requests to
// a line whose miss is still in flight are not merged (L2 sectors go up, not down) and the kernel is bound by the L1
// data pipe (82 % busy at G = 1: 15.5 MB of key rows per ciphertext at 128 B/clk), so G = 1 is the default and the
// other values remain as tuning options.
constexpr int kKsThreads = 160;

template <int T, int BASEBIT, int G>
__global__ void __launch_bounds__(kKsThreads)
    keyswitch_kernel(const int32_t* __restrict__ ksk, int n, int64_t kstride, int64_t stride,
                     const int32_t* __restrict__ u, const KsJob* __restrict__ jobs, int64_t count,
                     int32_t* __restrict__ pool) {
    constexpr int BASE = 1 << BASEBIT;
    constexpr uint32_t PREC_OFFSET = 1u << (32 - (1 + BASEBIT * T));
    __shared__ uint32_t abar[kN][G];  // [i][g]: the G words of one coefficient are one broadcast load
    const int t = threadIdx.x;
    const int64_t g0 = (int64_t)blockIdx.x * G;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (g0 + g < count) {
            const KsJob job = jobs[g0 + g];
            const int32_t* u0 = u + (int64_t)job.u0 * kUStride;
            const int32_t* u1 = job.u1 >= 0 ? u + (int64_t)job.u1 * kUStride : nullptr;
            for (int i = t; i < kN; i += kKsThreads) {
                uint32_t a = (uint32_t)u0[i];
                if (u1) a += (uint32_t)u1[i];
                abar[i][g] = a + PREC_OFFSET;
            }
        } else {
            for (int i = t; i < kN; i += kKsThreads) abar[i][g] = 0;  // all digits 0: nothing is fetched
        }
    }
    __syncthreads();
    const int w0 = 4 * t;
    if (w0 > n) return;
    uint4 acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = make_uint4(0, 0, 0, 0);
    // address of a row = p0 + (i*T*(BASE-1) + j*(BASE-1) + d-1) * row_bytes: one IADD3 + one IMAD.WIDE.U32 per digit
    const uint32_t row_bytes = (uint32_t)kstride * 4u;  // key rows are padded to 128-byte lines (see ksk_stride)
    uint64_t p0 = reinterpret_cast<uint64_t>(ksk) + 16u * (uint32_t)t;
    asm volatile("" : "+l"(p0));  // keep the per-thread base in a register pair (ptxas otherwise re-adds it per digit)
#pragma unroll 1
    for (int i = 0; i < kN; ++i) {
        uint32_t a[G];
#pragma unroll
        for (int g = 0; g < G; ++g) a[g] = abar[i][g];
        const uint32_t ki = (uint32_t)i * (T * (BASE - 1));
#pragma unroll
        for (int j = 0; j < T; ++j) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const uint32_t d = (a[g] >> (32 - (j + 1) * BASEBIT)) & (BASE - 1);
                if (d != 0) {
                    const uint32_t idx = d + ki + (uint32_t)(j * (BASE - 1) - 1);
                    uint64_t addr;
                    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(addr) : "r"(idx), "r"(row_bytes), "l"(p0));
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(addr));
                    acc[g].x -= v.x; acc[g].y -= v.y; acc[g].z -= v.z; acc[g].w -= v.w;
                }
            }
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (g0 + g >= count) break;
        const KsJob job = jobs[g0 + g];
        // b' (+ b'' + c0) goes to word n
        uint32_t bsum = (uint32_t)u[(int64_t)job.u0 * kUStride + kN] + (uint32_t)job.c0;
        if (job.u1 >= 0) bsum += (uint32_t)u[(int64_t)job.u1 * kUStride + kN];
        uint32_t r[4] = {acc[g].x, acc[g].y, acc[g].z, acc[g].w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (w0 + q == n) r[q] += bsum;
        int32_t* out = pool + job.out_slot * stride;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (w0 + q <= n) out[w0 + q] = (int32_t)r[q];
    }
}

// K2 for narrow levels: the keyswitch of ONE ciphertext split over S CTAs (coefficient ranges), partial sums added with
// red.global.add (integer addition mod 2^32 commutes: same words as the one-CTA kernel).  A narrow circuit level is
// latency-bound: one CTA walks 8,192 digit positions in 0.8 ms whatever the batch; split 16 ways that is ~0.1 ms.
// ks_init_kernel first sets every output to (0, ..., 0, b' [+ b''] + c0).
__global__ void ks_init_kernel(int n, int64_t stride, const int32_t* __restrict__ u, const KsJob* __restrict__ jobs,
                               int64_t count, int32_t* __restrict__ pool) {
    const int64_t g = blockIdx.x;
    if (g >= count) return;
    const KsJob job = jobs[g];
    int32_t* out = pool + job.out_slot * stride;
    for (int w = threadIdx.x; w < stride; w += blockDim.x) {
        uint32_t v = 0;
        if (w == n) {
            v = (uint32_t)u[(int64_t)job.u0 * kUStride + kN] + (uint32_t)job.c0;
            if (job.u1 >= 0) v += (uint32_t)u[(int64_t)job.u1 * kUStride + kN];
        }
        out[w] = (int32_t)v;
    }
}
template <int T, int BASEBIT>
__global__ void __launch_bounds__(kKsThreads)
    keyswitch_split_kernel(const int32_t* __restrict__ ksk, int n, int64_t kstride, int64_t stride,
                           const int32_t* __restrict__ u, const KsJob* __restrict__ jobs, int64_t count, int parts,
                           int32_t* __restrict__ pool) {
    constexpr int BASE = 1 << BASEBIT;
    constexpr uint32_t PREC_OFFSET = 1u << (32 - (1 + BASEBIT * T));
    __shared__ uint32_t abar[kN];
    const int t = threadIdx.x;
    const int64_t g = blockIdx.x / parts;
    const int part = blockIdx.x % parts, span = kN / parts, i0 = part * span;
    if (g >= count) return;
    const KsJob job = jobs[g];
    const int32_t* u0 = u + (int64_t)job.u0 * kUStride;
    const int32_t* u1 = job.u1 >= 0 ? u + (int64_t)job.u1 * kUStride : nullptr;
    for (int i = t; i < span; i += kKsThreads) {
        uint32_t a = (uint32_t)u0[i0 + i];
        if (u1) a += (uint32_t)u1[i0 + i];
        abar[i] = a + PREC_OFFSET;
    }
    __syncthreads();
    const int w0 = 4 * t;
    if (w0 > n) return;
    uint4 acc = make_uint4(0, 0, 0, 0);
    const uint32_t row_bytes = (uint32_t)kstride * 4u;
    uint64_t p0 = reinterpret_cast<uint64_t>(ksk) + 16u * (uint32_t)t;
    asm volatile("" : "+l"(p0));
#pragma unroll 1
    for (int i = 0; i < span; ++i) {
        const uint32_t a = abar[i];
        const uint32_t ki = (uint32_t)(i0 + i) * (T * (BASE - 1));
#pragma unroll
        for (int j = 0; j < T; ++j) {
            const uint32_t d = (a >> (32 - (j + 1) * BASEBIT)) & (BASE - 1);
            if (d != 0) {
                const uint32_t idx = d + ki + (uint32_t)(j * (BASE - 1) - 1);
                uint64_t addr;
                asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(addr) : "r"(idx), "r"(row_bytes), "l"(p0));
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(addr));
                acc.x -= v.x; acc.y -= v.y; acc.z -= v.z; acc.w -= v.w;
            }
        }
    }
    unsigned int* out = reinterpret_cast<unsigned int*>(pool + job.out_slot * stride);
    const uint32_t r[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (w0 + q <= n) atomicAdd(out + w0 + q, r[q]);
}

// K2 through a table of ROW SUMS (the default keyswitch, every level width).  The M digits of a group (M = 2, 4) select M
// key rows whose sum is subtracted; the table holds that sum for every combined digit value D = 1 .. 4^M - 1, so one row
// fetch replaces up to M (on average 3M/4): 2,040 instead of 6,144 rows of 2.5 KB per ciphertext at M = 4, a third of the
// bytes through the L2 -> SM path that binds the per-digit kernel (ncu r1s: lts 83 %).  The same rows of the same key are
// added in another order, and addition mod 2^32 is associative: the output words are identical.  Cost: (4^M - 1)/(3M)
// times the key in HBM (1.34 GB at M = 4, built on the device when a keyswitch first needs it).  The CTAs of a launch
// start together and take the same time, so they walk the coefficients in step and the part of the table in use
// (1.3 MB per coefficient) stays in L2 (ncu r2p: hit rate 71 %, DRAM 11.8 GB per 16,384).
// Layout: [i][g][D-1][stride], g = group of digits gM .. gM+M-1, the first digit in the top bits of D.
template <int T, int BASEBIT, int M>
__global__ void ksk_lut_build_kernel(const int32_t* __restrict__ ksk, int64_t kstride, int32_t* __restrict__ lut) {
    constexpr int BASE = 1 << BASEBIT, NG = T / M, ROWS = (1 << (BASEBIT * M)) - 1;
    const int64_t r = blockIdx.x;  // output row
    const int D = (int)(r % ROWS) + 1;
    const int64_t ig = r / ROWS;
    const int g = (int)(ig % NG);
    const int64_t i = ig / NG;
    const int32_t* src[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const int d = (D >> (BASEBIT * (M - 1 - m))) & (BASE - 1);
        src[m] = d ? ksk + ((i * T + g * M + m) * (BASE - 1) + d - 1) * kstride : nullptr;
    }
    for (int w = threadIdx.x; w < kstride; w += blockDim.x) {
        uint32_t s = 0;
#pragma unroll
        for (int m = 0; m < M; ++m)
            if (src[m]) s += (uint32_t)src[m][w];
        lut[r * kstride + w] = (int32_t)s;
    }
}

// G ciphertexts per CTA walk the coefficients together (thread t keeps words 4t..4t+3 of all G): a wide level then has
// G times fewer waves of CTAs, and each wave streams the table from HBM once.  U = coefficients per unrolled iteration:
// the loads a thread has in flight (U x T/M x G) decide the speed (profiles/r2p_keyswitch_lut_kbench.log).
template <int T, int BASEBIT, int M, bool SPLIT, int G, int U = 4>
__global__ void __launch_bounds__(kKsThreads)
    keyswitch_lut_kernel(const int32_t* __restrict__ lut, int n, int64_t kstride, int64_t stride,
                         const int32_t* __restrict__ u, const KsJob* __restrict__ jobs, int64_t count, int parts,
                         int32_t* __restrict__ pool) {
    static_assert(!SPLIT || G == 1, "a split ciphertext has a CTA range to itself");
    constexpr int NG = T / M, FIELD = BASEBIT * M;
    constexpr uint32_t ROWS = (1u << FIELD) - 1;
    constexpr uint32_t PREC_OFFSET = 1u << (32 - (1 + BASEBIT * T));
    __shared__ __align__(16) uint32_t abar[kN][G];  // [i][g]: the G words of a coefficient are one broadcast load
    const int t = threadIdx.x;
    // SPLIT (narrow levels): the coefficients of a ciphertext are split over `parts` CTAs, partial sums by
    // red.global.add onto what ks_init_kernel wrote — as in keyswitch_split_kernel
    const int64_t g0 = SPLIT ? blockIdx.x / parts : (int64_t)blockIdx.x * G;
    const int part = SPLIT ? blockIdx.x % parts : 0, span = SPLIT ? kN / parts : kN, i0 = part * span;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (g0 + g < count) {
            const KsJob job = jobs[g0 + g];
            const int32_t* u0 = u + (int64_t)job.u0 * kUStride;
            const int32_t* u1 = job.u1 >= 0 ? u + (int64_t)job.u1 * kUStride : nullptr;
            for (int i = t; i < span; i += kKsThreads) {
                uint32_t a = (uint32_t)u0[i0 + i];
                if (u1) a += (uint32_t)u1[i0 + i];
                abar[i][g] = a + PREC_OFFSET;
            }
        } else {
            for (int i = t; i < span; i += kKsThreads) abar[i][g] = 0;  // beyond the batch: every group reads D = 0, nothing is fetched
        }
    }
    __syncthreads();
    const int w0 = 4 * t;
    if (w0 > n) return;
    uint4 acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = make_uint4(0, 0, 0, 0);
    const uint32_t row_bytes = (uint32_t)kstride * 4u;
    uint64_t p0 = reinterpret_cast<uint64_t>(lut) + 16u * (uint32_t)t;
    asm volatile("" : "+l"(p0));
#pragma unroll(U)
    for (int i = 0; i < span; ++i) {
        uint32_t a[G];
#pragma unroll
        for (int g = 0; g < G; ++g) a[g] = abar[i][g];
#pragma unroll
        for (int q = 0; q < NG; ++q) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const uint32_t D = (a[g] >> (32 - (q + 1) * FIELD)) & ROWS;
                if (D != 0) {
                    const uint32_t idx = ((uint32_t)(i0 + i) * NG + q) * ROWS + D - 1;
                    uint64_t addr;
                    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(addr) : "r"(idx), "r"(row_bytes), "l"(p0));
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(addr));
                    acc[g].x -= v.x; acc[g].y -= v.y; acc[g].z -= v.z; acc[g].w -= v.w;
                }
            }
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (g0 + g >= count) break;
        const KsJob job = jobs[g0 + g];
        uint32_t r[4] = {acc[g].x, acc[g].y, acc[g].z, acc[g].w};
        int32_t* out = pool + job.out_slot * stride;
        if (!SPLIT) {
            uint32_t bsum = (uint32_t)u[(int64_t)job.u0 * kUStride + kN] + (uint32_t)job.c0;  // b' (+ b'' + c0) goes to word n
            if (job.u1 >= 0) bsum += (uint32_t)u[(int64_t)job.u1 * kUStride + kN];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (w0 + q == n) r[q] += bsum;
                if (w0 + q <= n) out[w0 + q] = (int32_t)r[q];
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (w0 + q <= n) atomicAdd(reinterpret_cast<unsigned int*>(out) + w0 + q, r[q]);
        }
    }
}

// K2 for wide levels, TILED: the key is streamed once per TILE of ciphertexts instead of once per ciphertext.
// The one-CTA-per-ciphertext kernel above pulls ~6,144 rows of 2.5 KB per ciphertext through the L1 data pipe: 214 GB of
// L2->SM traffic per 16,384 ciphertexts (ncu r1s: lts 83 %, L1 70 % busy) for 62 MB of key.  Here a thread keeps 8 output
// words of G ciphertexts in registers (8 G accumulators), loads the 3 rows of a digit position ONCE (its 8 words of each:
// 6 LDG.128) and subtracts, per ciphertext, the row its digit selects.  The digit of a (position, ciphertext) is the
// same for every thread of the group, so the selection is a warp-uniform branch around 8 integer subtractions; the key
// traffic through L1/L2 drops G-fold.  A CTA = 2 groups of 80 threads (words 8tt .. 8tt+7 of a 640-word padded row),
// i.e. 2 G ciphertexts; the digits of a chunk of coefficients wait in shared memory.  PARTS > 1 splits the 1,024
// coefficients of a tile over several CTAs (finer load balance; partial sums by red.global.add after ks_init_kernel).
// Integer additions mod 2^32 commute: the same words as the other two kernels (tests/test_gpu_parity.py).
constexpr int kKsTileGroup = 80;            // threads per ciphertext group: 80 x 8 words = one padded key row
constexpr int kKsTileThreads = 2 * kKsTileGroup;
constexpr int kKsTileChunk = 64;            // coefficients per shared-memory chunk of digits

template <int T, int BASEBIT, int G>
__global__ void __launch_bounds__(kKsTileThreads)
    keyswitch_tile_kernel(const int32_t* __restrict__ ksk, int n, int64_t kstride, int64_t stride,
                          const int32_t* __restrict__ u, const KsJob* __restrict__ jobs, int64_t count, int parts,
                          int32_t* __restrict__ pool) {
    constexpr int BASE = 1 << BASEBIT;
    constexpr uint32_t PREC_OFFSET = 1u << (32 - (1 + BASEBIT * T));
    __shared__ __align__(16) uint32_t abar[kKsTileChunk][2 * G];
    const int t = threadIdx.x, h = t / kKsTileGroup, tt = t % kKsTileGroup;
    const int64_t tile = blockIdx.x / parts;
    const int part = blockIdx.x % parts, span = kN / parts, i0 = part * span;
    const int64_t g0 = tile * (2 * G);
    uint4 acc[G][2];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g][0] = acc[g][1] = make_uint4(0, 0, 0, 0);
    const uint32_t row_bytes = (uint32_t)kstride * 4u;
    // words 4tt..4tt+3 and 320+4tt..320+4tt+3 of a 640-word row: both 16-byte loads of a warp are contiguous
    const unsigned char* kp = reinterpret_cast<const unsigned char*>(ksk) + 16u * (uint32_t)tt;
    for (int c0 = 0; c0 < span; c0 += kKsTileChunk) {
        __syncthreads();  // the previous chunk's digits are no longer read
        for (int g = 0; g < 2 * G; ++g) {
            if (t < kKsTileChunk) {
                uint32_t v = 0;  // a ciphertext beyond the batch: all digits 0
                if (g0 + g < count) {
                    const KsJob job = jobs[g0 + g];
                    v = (uint32_t)u[(int64_t)job.u0 * kUStride + i0 + c0 + t] + PREC_OFFSET;
                    if (job.u1 >= 0) v += (uint32_t)u[(int64_t)job.u1 * kUStride + i0 + c0 + t];
                }
                abar[t][g] = v;
            }
        }
        __syncthreads();
        // positions (i, j) of the chunk in one rolled loop (an unrolled body of 8 x G branches overflows the instruction
        // cache: 37 ms instead of 11); the rows of position p + 1 are requested before position p is consumed
        const unsigned char* kr = kp + (size_t)(i0 + c0) * (T * (BASE - 1)) * row_bytes;
        uint4 r[3][2], nx[3][2];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            nx[d][0] = __ldg(reinterpret_cast<const uint4*>(kr + (size_t)d * row_bytes));
            nx[d][1] = __ldg(reinterpret_cast<const uint4*>(kr + (size_t)d * row_bytes + 1280));
        }
        uint32_t a[G];
#pragma unroll 1
        for (int p = 0; p < kKsTileChunk * T; ++p) {
            const int j = p & (T - 1);
#pragma unroll
            for (int d = 0; d < 3; ++d) { r[d][0] = nx[d][0]; r[d][1] = nx[d][1]; }
            kr += (size_t)(BASE - 1) * row_bytes;
            if (p + 1 < kKsTileChunk * T) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    nx[d][0] = __ldg(reinterpret_cast<const uint4*>(kr + (size_t)d * row_bytes));
                    nx[d][1] = __ldg(reinterpret_cast<const uint4*>(kr + (size_t)d * row_bytes + 1280));
                }
            }
            if (j == 0) {
#pragma unroll
                for (int q = 0; q < G / 4; ++q) {
                    const uint4 v = *reinterpret_cast<const uint4*>(&abar[p / T][h * G + 4 * q]);
                    a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
                }
            }
            const int sh = 32 - (j + 1) * BASEBIT;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const uint32_t d = (a[g] >> sh) & (BASE - 1);
                if (d == 1) {
                    acc[g][0].x -= r[0][0].x; acc[g][0].y -= r[0][0].y; acc[g][0].z -= r[0][0].z; acc[g][0].w -= r[0][0].w;
                    acc[g][1].x -= r[0][1].x; acc[g][1].y -= r[0][1].y; acc[g][1].z -= r[0][1].z; acc[g][1].w -= r[0][1].w;
                } else if (d == 2) {
                    acc[g][0].x -= r[1][0].x; acc[g][0].y -= r[1][0].y; acc[g][0].z -= r[1][0].z; acc[g][0].w -= r[1][0].w;
                    acc[g][1].x -= r[1][1].x; acc[g][1].y -= r[1][1].y; acc[g][1].z -= r[1][1].z; acc[g][1].w -= r[1][1].w;
                } else if (d == 3) {
                    acc[g][0].x -= r[2][0].x; acc[g][0].y -= r[2][0].y; acc[g][0].z -= r[2][0].z; acc[g][0].w -= r[2][0].w;
                    acc[g][1].x -= r[2][1].x; acc[g][1].y -= r[2][1].y; acc[g][1].z -= r[2][1].z; acc[g][1].w -= r[2][1].w;
                }
            }
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const int64_t gi = g0 + h * G + g;
        if (gi >= count) break;
        const KsJob job = jobs[gi];
        uint32_t r[8] = {acc[g][0].x, acc[g][0].y, acc[g][0].z, acc[g][0].w, acc[g][1].x, acc[g][1].y, acc[g][1].z, acc[g][1].w};
        int32_t* out = pool + job.out_slot * stride;
        if (parts == 1) {
            uint32_t bsum = (uint32_t)u[(int64_t)job.u0 * kUStride + kN] + (uint32_t)job.c0;  // b' (+ b'' + c0) goes to word n
            if (job.u1 >= 0) bsum += (uint32_t)u[(int64_t)job.u1 * kUStride + kN];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int w = (q < 4 ? 4 * tt : 320 + 4 * tt) + (q & 3);
                if (w == n) r[q] += bsum;
                if (w <= n) out[w] = (int32_t)r[q];
            }
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int w = (q < 4 ? 4 * tt : 320 + 4 * tt) + (q & 3);
                if (w <= n) atomicAdd(reinterpret_cast<unsigned int*>(out) + w, r[q]);
            }
        }
    }
}

// K4.  out = sign*in + (0, cst); in < 0: absent.
__global__ void linear_kernel(int32_t* __restrict__ pool, int64_t stride, int n, const int64_t* __restrict__ in,
                              const int32_t* __restrict__ sign, const int32_t* __restrict__ cst,
                              const int64_t* __restrict__ out) {
    const int64_t g = blockIdx.x;
    const int64_t src = in[g];
    const uint32_t s = (uint32_t)sign[g];
    for (int w = threadIdx.x; w <= n; w += blockDim.x) {
        uint32_t v = src >= 0 ? s * (uint32_t)pool[src * stride + w] : 0u;
        if (w == n) v += (uint32_t)cst[g];
        pool[out[g] * stride + w] = (int32_t)v;
    }
}

// pool initialisation: trivial false (a = 0, b = -1/8) — SURVEY.md §8a quirk Q2.
__global__ void pool_init_kernel(int32_t* __restrict__ pool, int64_t stride, int n, int64_t first, int64_t slots) {
    const int64_t total = slots * stride;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(idx % stride);
        pool[first * stride + idx] = (w == n) ? -(1 << 29) : 0;
    }
}

// repack of the keyswitch key: [i][j][d][n+1] -> [i][j][d-1][stride]
__global__ void ksk_pack_kernel(const int32_t* __restrict__ src, int n, int base, int64_t stride, int64_t rows_out,
                                int32_t* __restrict__ dst) {
    const int64_t r = blockIdx.x;  // output row = (i*t + j)*(base-1) + (d-1)
    if (r >= rows_out) return;
    const int64_t ij = r / (base - 1);
    const int d = (int)(r % (base - 1)) + 1;
    const int32_t* s = src + (ij * base + d) * (int64_t)(n + 1);
    for (int w = threadIdx.x; w < stride; w += blockDim.x) dst[r * stride + w] = w <= n ? s[w] : 0;
}

// scatter / gather between a dense staging buffer [count][n+1] and pool slots
__global__ void scatter_kernel(int32_t* __restrict__ pool, int64_t stride, int n, const int64_t* __restrict__ slots,
                               int64_t first_slot, const int32_t* __restrict__ dense) {
    const int64_t g = blockIdx.x;
    const int64_t s = slots ? slots[g] : first_slot + g;
    for (int w = threadIdx.x; w <= n; w += blockDim.x) pool[s * stride + w] = dense[g * (n + 1) + w];
}
__global__ void gather_kernel(const int32_t* __restrict__ pool, int64_t stride, int n,
                              const int64_t* __restrict__ slots, int64_t first_slot, int32_t* __restrict__ dense) {
    const int64_t g = blockIdx.x;
    const int64_t s = slots ? slots[g] : first_slot + g;
    for (int w = threadIdx.x; w <= n; w += blockDim.x) dense[g * (n + 1) + w] = pool[s * stride + w];
}
// u (padded stride) <-> dense [count][N+1]
__global__ void u_unpad_kernel(int32_t* __restrict__ u, int32_t* __restrict__ dense, int to_dense) {
    const int64_t g = blockIdx.x;
    for (int w = threadIdx.x; w <= kN; w += blockDim.x) {
        if (to_dense) dense[g * (kN + 1) + w] = u[g * kUStride + w];
        else u[g * kUStride + w] = dense[g * (kN + 1) + w];
    }
}

// FP64 FMA peak probe: 8 independent DFMA chains per thread.
__global__ void fp64_peak_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host side
struct peba1_cuda_ctx {
    int device = 0;
    int sm_count = 0;
    peba1_cuda_params p{};
    int rows = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // keys
    double2* bkfft = nullptr;
    size_t bkfft_bytes = 0;
    int32_t* ksk = nullptr;
    size_t ksk_bytes = 0;
    bool bk_loaded = false, ksk_loaded = false;
    // pool
    int32_t* pool = nullptr;
    int64_t pool_slots = 0, pool_stride = 0;
    int64_t ksk_stride = 0;  // words per packed keyswitch-key row
    // table of row sums for the keyswitch of wide levels (keyswitch_lut_kernel): built on first use, dropped when the key changes
    int32_t* ksk_lut = nullptr;
    int ksk_lut_digits = 0;  // digits per group the table was built for (0 = not built)
    int ksk_lut_alloc = 0;   // digits per group the buffer is sized for
    // workspaces (grown on demand)
    int32_t* u = nullptr;
    int64_t u_cap = 0;  // in samples
    void* d_jobs = nullptr;
    size_t d_jobs_cap = 0;
    unsigned long long* d_job_counter = nullptr;  // persistent blind-rotation kernels draw job indices from it
    // pinned job lists, double buffered: peba1_cuda_gates returns once its launches are queued, and only waits (on the
    // copy event) before it overwrites the list it filled two chunks ago
    void* h_jobs[2] = {nullptr, nullptr};
    size_t h_jobs_cap[2] = {0, 0};
    cudaEvent_t jobs_copied[2] = {nullptr, nullptr};
    bool jobs_pending[2] = {false, false};
    int jobs_next = 0;
    int32_t* d_stage = nullptr;  // dense staging for uploads/downloads
    size_t d_stage_cap = 0;
    int32_t* h_stage = nullptr;  // pinned
    size_t h_stage_cap = 0;
    // private rows of the host-buffer entry points (gate_batch_host, keyswitch_host): same stride as the pool, never
    // visible to the caller, so those calls leave every pool slot untouched
    int32_t* hb_pool = nullptr;
    int64_t hb_slots = 0;
    cudaEvent_t stream_switch = nullptr;  // orders work across peba1_cuda_set_stream
    // level log (peba1_cuda_level_log_begin/_end): three events per launch pair, resolved at _end without syncing between
    bool log_on = false;
    std::vector<cudaEvent_t> log_ev;
    std::vector<int64_t> log_gates, log_pbs;
    size_t log_n = 0;
    // options
    int teams_per_cta = 6;
    int tm_teams = 0;        // teams per CTA of the tensor-memory kernel; 0 = by batch size
    int narrow_two_buffers = 1;  // the narrow shapes (fixed grids of 1..8 teams) run on TeamSmemTm2 as well; 0 needs VARIANTS=all
    int ks_group = 1;        // ciphertexts per keyswitch CTA (measured: 1 is fastest, the L1 pipe is the bound)
    int ks_tile = 0;         // ciphertexts per group of the tiled keyswitch (8, 12, 16; 0 = off), 2 groups per CTA
    int ks_tile_parts = 1;   // CTAs per tile (coefficient ranges, partial sums by red.global.add)
    int64_t ks_tile_min = 4096;  // levels narrower than this keep the per-ciphertext kernels
    int ks_lut = 4;          // digits per table row of the wide-level keyswitch (4: 255 row sums per group, 1.34 GB; 2: 15, 157 MB; 0 = off)
    int ks_lut_group = 4;    // ciphertexts per CTA of that kernel on very wide levels (1 or 4)
    int64_t ks_lut_group_min = -1;  // ... from this level width on; -1 = eight waves of single-ciphertext CTAs (14,208 on 148 SMs)
    int64_t ks_lut_min = 1;  // levels narrower than this keep the per-digit kernels (a context that stays below it never builds the table)
    int64_t ks_split = -1;   // CTAs per ciphertext in the keyswitch of a narrow level: -1 by batch size, 0 never, else fixed (power of 2)
    int kernel_variant = 1;  // see include/peba1_cuda.h (peba1_cuda_set_option): 1 is the product, 0 and 2..10 are the
                             // measured alternatives
    int64_t wide_max_batch = -1;  // >= 0: batches up to this size use the low-latency kernel; -1: the measured rule
    int tmk_stagger_ns = 0;       // variant 4: start-up delay per team index
    int64_t pair_min_batch = -1;  // variants 3..10 below this batch size fall back to variant 1's shapes; -1 = one full wave
    int64_t max_pbs_per_launch = 1 << 18;
    // stats
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    peba1_cuda_stats stats{};
};

namespace {

int ensure_dev(void** ptr, size_t* cap, size_t need) {
    if (*cap >= need) return 0;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
    size_t want = std::max(need, (size_t)1 << 20);
    CU(cudaMalloc(ptr, want));
    *cap = want;
    return 0;
}
int ensure_pinned(void** ptr, size_t* cap, size_t need) {
    if (*cap >= need) return 0;
    if (*ptr) cudaFreeHost(*ptr);
    *ptr = nullptr;
    *cap = 0;
    size_t want = std::max(need, (size_t)1 << 20);
    CU(cudaMallocHost(ptr, want));
    *cap = want;
    return 0;
}

// temporary device buffer of one call: freed on every return path
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

int use_device(peba1_cuda_ctx* c) {
    if (!c) return fail(PEBA1_EINVAL, "null context");
    CU(cudaSetDevice(c->device));
    return 0;
}

bool params_supported(const peba1_cuda_params& p) {
    if (p.N != kN || p.k != 1) return false;
    if (!((p.l == 3 && p.Bgbit == 7) || (p.l == 2 && p.Bgbit == 10))) return false;
    if (p.ks_t != 8 || p.ks_basebit != 2) return false;
    if (p.n < 1 || p.n > 639) return false;
    return true;
}

size_t team_smem_bytes(int teams) { return sizeof(TeamSmem) * (size_t)teams; }

#ifdef PEBA1_ALL_VARIANTS

template <int TEAMS>
int launch_pbs_t(peba1_cuda_ctx* c, const int32_t* pool, int64_t stride, const PbsJob* jobs, int64_t count,
                 int32_t* u) {
    const size_t smem = team_smem_bytes(TEAMS);
    const unsigned grid = (unsigned)((count + TEAMS - 1) / TEAMS);
    if (c->p.l == 3) {
        auto k = pbs_kernel<3, 7, TEAMS>;
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, TEAMS * kTeam, smem, c->stream>>>(c->bkfft, c->p.n, pool, stride, jobs, count, u);
    } else {
        auto k = pbs_kernel<2, 10, TEAMS>;
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, TEAMS * kTeam, smem, c->stream>>>(c->bkfft, c->p.n, pool, stride, jobs, count, u);
    }
    CU(cudaGetLastError());
    c->stats.launches++;
    return 0;
}

#endif  // PEBA1_ALL_VARIANTS
// narrow shapes: fixed grid, TEAMS (1..8) teams per CTA.  Three instantiations serve them — compiled for 2, 4 and 8 teams
// (launch bounds of 128 / 256 / 512 threads: with one or two warps per scheduler ptxas needs the registers beyond 128 to
// schedule for latency; the 8-team build measured 7.1 ms instead of 4.6 for 296 bootstraps) — each launched with
// TEAMS * 64 threads and TEAMS team areas of shared memory.
template <int L, int BGBIT, int TEAMS, class SM>
int launch_pbs_tm_ts(peba1_cuda_ctx* c, const int32_t* pool, int64_t stride, const PbsJob* jobs, int64_t count,
                     int32_t* u) {
    constexpr int BUILT = TEAMS <= 2 ? 2 : TEAMS <= 4 ? 4 : kTeamsTm;
    const size_t smem = sizeof(SM) * (size_t)TEAMS;
    const unsigned grid = (unsigned)((count + TEAMS - 1) / TEAMS);
    auto k = pbs_kernel_tm<L, BGBIT, BUILT, SM>;
    CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(SM) * (size_t)BUILT)));
    k<<<grid, TEAMS * kTeam, smem, c->stream>>>(c->bkfft, c->p.n, pool, stride, jobs, count, u, nullptr);
    return 0;
}
template <int L, int BGBIT, int TEAMS>
int launch_pbs_tm_t(peba1_cuda_ctx* c, const int32_t* pool, int64_t stride, const PbsJob* jobs, int64_t count,
                    int32_t* u) {
    // the narrow shapes run on the two-buffer layout too (measured r2g, kbench: 296 bootstraps 4.72 -> 4.55 ms, 1,036:
    // 10.63 -> 10.33 ms, 648: 8.33 = 8.33 ms); the one-buffer layout with its "buffer free" mbarrier is a VARIANTS=all build
#ifdef PEBA1_ALL_VARIANTS
    if (!c->narrow_two_buffers) return launch_pbs_tm_ts<L, BGBIT, TEAMS, TeamSmemTm>(c, pool, stride, jobs, count, u);
#endif
    return launch_pbs_tm_ts<L, BGBIT, TEAMS, TeamSmemTm2>(c, pool, stride, jobs, count, u);
}

template <int L, int BGBIT>
int launch_pbs_tm_l(peba1_cuda_ctx* c, const int32_t* pool, int64_t stride, const PbsJob* jobs, int64_t count,
                    int32_t* u) {
#ifdef PEBA1_ALL_VARIANTS
    if (c->kernel_variant == 2) {
        const size_t smem = sizeof(TeamSmemTm) * (size_t)kTeamsTm;
        const unsigned grid = (unsigned)((count + kTeamsTm - 1) / kTeamsTm);
        auto k = pbs_kernel_tmsh<L, BGBIT>;
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, kTeamsTm * kTeam, smem, c->stream>>>(c->bkfft, c->p.n, pool, stride, jobs, count, u);
        return 0;
    }
    if ((c->kernel_variant == 5 || c->kernel_variant == 6) &&
        count >= (c->pair_min_batch >= 0 ? c->pair_min_batch : (int64_t)c->sm_count * kTeamsTm10)) {
        // 10 teams per CTA on the single-exchange-buffer layout; variant 6: persistent CTAs drawing jobs from a counter
        const size_t smem = sizeof(TeamSmemTm1) * (size_t)kTeamsTm10;
        unsigned grid = (unsigned)((count + kTeamsTm10 - 1) / kTeamsTm10);
        unsigned long long* counter = nullptr;
        if (c->kernel_variant == 6) {
            if (!c->d_job_counter) CU(cudaMalloc(&c->d_job_counter, sizeof(unsigned long long)));
            CU(cudaMemsetAsync(c->d_job_counter, 0, sizeof(unsigned long long), c->stream));
            counter = c->d_job_counter;
            grid = std::min(grid, (unsigned)c->sm_count);
        }
        auto k = pbs_kernel_tm<L, BGBIT, kTeamsTm10, TeamSmemTm1>;
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, kTeamsTm10 * kTeam, smem, c->stream>>>(c->bkfft, c->p.n, pool, stride, jobs, count, u, counter);
        return 0;
    }
    if (c->kernel_variant == 9 &&
        count >= (c->pair_min_batch >= 0 ? c->pair_min_batch : (int64_t)c->sm_count * 6)) {
        // 6 teams per SM with the MAC accumulators in registers (blind_rotate_team_tmr), persistent CTAs
        constexpr int T6 = 6;
        const size_t smem = sizeof(TeamSmemTm) * (size_t)T6;
        if (!c->d_job_counter) CU(cudaMalloc(&c->d_job_counter, sizeof(unsigned long long)));
        CU(cudaMemsetAsync(c->d_job_counter, 0, sizeof(unsigned long long), c->stream));
        const unsigned grid = std::min((unsigned)((count + T6 - 1) / T6), (unsigned)c->sm_count);
        auto k = pbs_kernel_tm<L, BGBIT, T6, TeamSmemTm, true>;
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, T6 * kTeam, smem, c->stream>>>(c->bkfft, c->p.n, pool, stride, jobs, count, u, c->d_job_counter);
        return 0;
    }
#endif  // PEBA1_ALL_VARIANTS
    if (c->kernel_variant == 8 &&
        count >= (c->pair_min_batch >= 0 ? c->pair_min_batch : (int64_t)c->sm_count * kTeamsTm)) {
        // two alternating exchange buffers without the "buffer free" mbarrier (TeamSmemTm2): what the default does for
        // wide batches; as a variant it also applies below one full wave (pair_min_batch)
        const size_t smem = sizeof(TeamSmemTm2) * (size_t)kTeamsTm;
        if (!c->d_job_counter) CU(cudaMalloc(&c->d_job_counter, sizeof(unsigned long long)));
        CU(cudaMemsetAsync(c->d_job_counter, 0, sizeof(unsigned long long), c->stream));
        const unsigned grid = std::min((unsigned)((count + kTeamsTm - 1) / kTeamsTm), (unsigned)c->sm_count);
        auto k = pbs_kernel_tm<L, BGBIT, kTeamsTm, TeamSmemTm2>;
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, kTeamsTm * kTeam, smem, c->stream>>>(c->bkfft, c->p.n, pool, stride, jobs, count, u, c->d_job_counter);
        return 0;
    }
    if (c->kernel_variant == 11 &&
        count >= (c->pair_min_batch >= 0 ? c->pair_min_batch : (int64_t)c->sm_count * kTeamsTm)) {
        // cross-team exchange of every transform through tensor memory (TeamSmemTmX)
        const size_t smem = sizeof(TeamSmemTmX) * (size_t)kTeamsTm;
        if (!c->d_job_counter) CU(cudaMalloc(&c->d_job_counter, sizeof(unsigned long long)));
        CU(cudaMemsetAsync(c->d_job_counter, 0, sizeof(unsigned long long), c->stream));
        const unsigned grid = std::min((unsigned)((count + kTeamsTm - 1) / kTeamsTm), (unsigned)c->sm_count);
        auto k = pbs_kernel_tm<L, BGBIT, kTeamsTm, TeamSmemTmX>;
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, kTeamsTm * kTeam, smem, c->stream>>>(c->bkfft, c->p.n, pool, stride, jobs, count, u, c->d_job_counter);
        return 0;
    }
#ifdef PEBA1_ALL_VARIANTS
    if (c->kernel_variant == 10 &&
        count >= (c->pair_min_batch >= 0 ? c->pair_min_batch : (int64_t)c->sm_count * kTeamsTm)) {
        // variant 8 with the inner exchange of every transform done by warp shuffles (TeamSmemTm2S)
        const size_t smem = sizeof(TeamSmemTm2S) * (size_t)kTeamsTm;
        if (!c->d_job_counter) CU(cudaMalloc(&c->d_job_counter, sizeof(unsigned long long)));
        CU(cudaMemsetAsync(c->d_job_counter, 0, sizeof(unsigned long long), c->stream));
        const unsigned grid = std::min((unsigned)((count + kTeamsTm - 1) / kTeamsTm), (unsigned)c->sm_count);
        auto k = pbs_kernel_tm<L, BGBIT, kTeamsTm, TeamSmemTm2S>;
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, kTeamsTm * kTeam, smem, c->stream>>>(c->bkfft, c->p.n, pool, stride, jobs, count, u, c->d_job_counter);
        return 0;
    }
#endif  // PEBA1_ALL_VARIANTS
#ifdef PEBA1_ALL_VARIANTS
    if (c->kernel_variant == 7 &&
        count >= (c->pair_min_batch >= 0 ? c->pair_min_batch : (int64_t)c->sm_count * kTeamsTm)) {
        // persistent 8-team CTAs on the single cross-team exchange buffer with its "buffer free" mbarrier (TeamSmemTm): the
        // default until r1zz, 1.4 % slower than the two-buffer layout now
        const size_t smem = sizeof(TeamSmemTm) * (size_t)kTeamsTm;
        if (!c->d_job_counter) CU(cudaMalloc(&c->d_job_counter, sizeof(unsigned long long)));
        CU(cudaMemsetAsync(c->d_job_counter, 0, sizeof(unsigned long long), c->stream));
        const unsigned grid = std::min((unsigned)((count + kTeamsTm - 1) / kTeamsTm), (unsigned)c->sm_count);
        auto k = pbs_kernel_tm<L, BGBIT, kTeamsTm, TeamSmemTm>;
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, kTeamsTm * kTeam, smem, c->stream>>>(c->bkfft, c->p.n, pool, stride, jobs, count, u, c->d_job_counter);
        return 0;
    }
#endif
#ifdef PEBA1_ALL_VARIANTS
    if (c->kernel_variant == 4 &&
        count >= (c->pair_min_batch >= 0 ? c->pair_min_batch : (int64_t)c->sm_count * kTeamsTm)) {
        const size_t smem = sizeof(TeamSmemTm) * (size_t)kTeamsTm + 3 * kKeyHalfBytes;
        const unsigned grid = (unsigned)((count + kTeamsTm - 1) / kTeamsTm);
        auto k = pbs_kernel_tmk<L, BGBIT>;
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, kTmkThreads, smem, c->stream>>>(c->bkfft, c->p.n, pool, stride, jobs, count, u, c->tmk_stagger_ns);
        return 0;
    }
    if (c->kernel_variant == 3 &&
        count >= (c->pair_min_batch >= 0 ? c->pair_min_batch : (int64_t)c->sm_count * 2 * kTeamsPair)) {
        const size_t smem = sizeof(TeamSmemPair) * (size_t)kTeamsPair;
        const unsigned grid = (unsigned)((count + 2 * kTeamsPair - 1) / (2 * kTeamsPair));
        auto k = pbs_kernel_pair<L, BGBIT, kTeamsPair>;
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, kTeamsPair * kTeam, smem, c->stream>>>(c->bkfft, c->p.n, pool, stride, jobs, count, u);
        return 0;
    }
#endif  // PEBA1_ALL_VARIANTS
    if (c->kernel_variant == 1 && c->tm_teams == 0 && count >= (int64_t)c->sm_count * kTeamsTm) {
        // wide batch (the default path): one persistent 8-team CTA per SM, jobs drawn from a counter, on two alternating
        // exchange buffers (= variant 8).  With the accumulator copies in tensor memory this instantiation has no spills
        // left and needs no "buffer free" mbarrier: 149.3 ms per 16,384 against 151.35 ms for variant 7 (the default
        // before; scripts/ubench/kbench.c, A-B-A on one box, identical output checksums).
        const size_t smem = sizeof(TeamSmemTm2) * (size_t)kTeamsTm;
        if (!c->d_job_counter) CU(cudaMalloc(&c->d_job_counter, sizeof(unsigned long long)));
        CU(cudaMemsetAsync(c->d_job_counter, 0, sizeof(unsigned long long), c->stream));
        auto k = pbs_kernel_tm<L, BGBIT, kTeamsTm, TeamSmemTm2>;
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<(unsigned)c->sm_count, kTeamsTm * kTeam, smem, c->stream>>>(c->bkfft, c->p.n, pool, stride, jobs, count, u,
                                                                     c->d_job_counter);
        return 0;
    }
    // Very narrow level (at most two bootstraps per SM): latency matters, not throughput — one CTA of 2l teams per
    // bootstrap, 2.17 ms per wave of up to 148 against 4.5 ms for a lone team.  Two waves (4.34 ms) still beat the
    // throughput kernel's T = 2 shape (4.57 ms for 296); from three waves on the narrow shapes below win (444: 5.4 ms).
    const bool wide = c->wide_max_batch < 0 ? count <= 2 * (int64_t)c->sm_count : count <= c->wide_max_batch;
    if (c->kernel_variant == 1 && c->tm_teams == 0 && wide) {
        const size_t smem = sizeof(WideSmem<2 * L>);
        const unsigned grid = (unsigned)std::min<int64_t>(count, (int64_t)c->sm_count * 4);
        auto k = pbs_kernel_wide<L, BGBIT>;
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, 2 * L * kTeam, smem, c->stream>>>(c->bkfft, c->p.n, pool, stride, jobs, count, u);
        return 0;
    }
    // Narrow levels (the long tail of a circuit: tens of gates per launch) cannot fill 148 SMs with 8 bootstraps per
    // CTA; fewer teams per CTA spread them over more SMs, where a team runs faster alone.
    int teams = c->tm_teams;
    // the smallest T (1..8) that still fits the level into one CTA per SM: a bootstrap is the faster the fewer teams share
    // its SM (one wave of T = 1 / 2 / 4 / 8 teams per SM: 4.6 / 4.8 / 5.9 / 10.9 ms)
    if (teams == 0) teams = (int)std::min<int64_t>(8, (count + c->sm_count - 1) / c->sm_count);
    switch (teams) {
        case 1: return launch_pbs_tm_t<L, BGBIT, 1>(c, pool, stride, jobs, count, u);
        case 2: return launch_pbs_tm_t<L, BGBIT, 2>(c, pool, stride, jobs, count, u);
        case 3: return launch_pbs_tm_t<L, BGBIT, 3>(c, pool, stride, jobs, count, u);
        case 4: return launch_pbs_tm_t<L, BGBIT, 4>(c, pool, stride, jobs, count, u);
        case 5: return launch_pbs_tm_t<L, BGBIT, 5>(c, pool, stride, jobs, count, u);
        case 6: return launch_pbs_tm_t<L, BGBIT, 6>(c, pool, stride, jobs, count, u);
        case 7: return launch_pbs_tm_t<L, BGBIT, 7>(c, pool, stride, jobs, count, u);
        default: return launch_pbs_tm_t<L, BGBIT, 8>(c, pool, stride, jobs, count, u);
    }
}

int launch_pbs_tm(peba1_cuda_ctx* c, const int32_t* pool, int64_t stride, const PbsJob* jobs, int64_t count,
                  int32_t* u) {
    int rc = c->p.l == 3 ? launch_pbs_tm_l<3, 7>(c, pool, stride, jobs, count, u)
                         : launch_pbs_tm_l<2, 10>(c, pool, stride, jobs, count, u);
    if (rc) return rc;
    CU(cudaGetLastError());
    c->stats.launches++;
    return 0;
}

int launch_pbs(peba1_cuda_ctx* c, const int32_t* pool, int64_t stride, const PbsJob* jobs, int64_t count,
               int32_t* u) {
    if (count <= 0) return 0;
    if (c->timing) CU(cudaEventRecord(c->ev0, c->stream));
    int rc;
    if (c->kernel_variant >= 1) {
        rc = launch_pbs_tm(c, pool, stride, jobs, count, u);
        if (rc) return rc;
        if (c->timing) CU(cudaEventRecord(c->ev1, c->stream));
        c->stats.pbs += (uint64_t)count;
        return 0;
    }
#ifdef PEBA1_ALL_VARIANTS
    switch (c->teams_per_cta) {
        case 1: rc = launch_pbs_t<1>(c, pool, stride, jobs, count, u); break;
        case 2: rc = launch_pbs_t<2>(c, pool, stride, jobs, count, u); break;
        case 3: rc = launch_pbs_t<3>(c, pool, stride, jobs, count, u); break;
        case 5: rc = launch_pbs_t<5>(c, pool, stride, jobs, count, u); break;
        case 4: rc = launch_pbs_t<4>(c, pool, stride, jobs, count, u); break;
        default: rc = launch_pbs_t<6>(c, pool, stride, jobs, count, u); break;
    }
    if (rc) return rc;
    if (c->timing) CU(cudaEventRecord(c->ev1, c->stream));
    c->stats.pbs += (uint64_t)count;
    return 0;
#else
    return fail(PEBA1_EINVAL, "kernel_variant 0 is not compiled into this build (make VARIANTS=all)");
#endif
}

// the table of row sums for c->ks_lut digits per group, built from the packed key on the context's stream
int ensure_ksk_lut(peba1_cuda_ctx* c) {
    if (c->ksk_lut_digits == c->ks_lut) return 0;
    const int M = c->ks_lut;
    const int64_t rows = (int64_t)kN * (c->p.ks_t / M) * ((1 << (c->p.ks_basebit * M)) - 1);
    c->ksk_lut_digits = 0;
    if (c->ksk_lut && c->ksk_lut_alloc != M) {  // another table shape: launches that read the old one are on this stream
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaFree(c->ksk_lut));
        c->ksk_lut = nullptr;
    }
    if (!c->ksk_lut) {
        const cudaError_t e = cudaMalloc(&c->ksk_lut, (size_t)rows * c->ksk_stride * sizeof(int32_t));
        if (e == cudaErrorMemoryAllocation) {
            // no room for the table on this device: this context keeps the per-digit kernels (same words, 3 x the rows)
            cudaGetLastError();
            c->ksk_lut = nullptr;
            fprintf(stderr, "peba1_cuda: %.2f GB for the keyswitch table of row sums not available; using the per-digit keyswitch\n",
                    (double)rows * c->ksk_stride * sizeof(int32_t) / 1e9);
            c->ks_lut = 0;
            return 0;
        }
        CU(e);
        c->ksk_lut_alloc = M;
    }
    if (M == 2) ksk_lut_build_kernel<8, 2, 2><<<(unsigned)rows, 160, 0, c->stream>>>(c->ksk, c->ksk_stride, c->ksk_lut);
    else ksk_lut_build_kernel<8, 2, 4><<<(unsigned)rows, 160, 0, c->stream>>>(c->ksk, c->ksk_stride, c->ksk_lut);
    CU(cudaGetLastError());
    c->stats.launches++;
    c->ksk_lut_digits = M;
    return 0;
}

int launch_ks(peba1_cuda_ctx* c, const int32_t* u, const KsJob* jobs, int64_t count, int32_t* pool) {
    if (count <= 0) return 0;
    const int G = c->ks_group;
    // narrow level: one wave of CTAs, each ciphertext split over up to 64 of them
    int parts = 1;
    if (G == 1 && c->ks_split != 0) {
        const int64_t room = (int64_t)c->sm_count * 12 / std::max<int64_t>(count, 1);
        while (parts * 2 <= room && parts < 64) parts *= 2;
        if (c->ks_split > 0) parts = (int)c->ks_split;
    }
    // one fetch per group of digits from the table of row sums (any width; narrow levels split as above)
    if (G == 1 && c->ks_lut > 0 && c->ks_tile == 0 && count >= c->ks_lut_min) {
        if (int rc = ensure_ksk_lut(c)) return rc;
    }
    if (G == 1 && c->ks_lut > 0 && c->ks_tile == 0 && count >= c->ks_lut_min) {  // (the table may just have been refused)
        if (parts > 1) ks_init_kernel<<<(unsigned)count, 160, 0, c->stream>>>(c->p.n, c->pool_stride, u, jobs, count, pool);
        // several ciphertexts per CTA once the level is wider than one wave of single-ciphertext CTAs
        const int64_t group_from = c->ks_lut_group_min >= 0 ? c->ks_lut_group_min : (int64_t)c->sm_count * 12 * 8;
        const int LG = (parts == 1 && c->ks_lut == 4 && count >= group_from) ? c->ks_lut_group : 1;
        const unsigned lgrid = parts > 1 ? (unsigned)(count * parts) : (unsigned)((count + LG - 1) / LG);
#define PEBA1_KS_LUT(M, SPLIT, GRP, UNR) keyswitch_lut_kernel<8, 2, M, SPLIT, GRP, UNR><<<lgrid, kKsThreads, 0, c->stream>>>( \
    c->ksk_lut, c->p.n, c->ksk_stride, c->pool_stride, u, jobs, count, parts, pool)
        // unroll depths by measurement (profiles/r2p_keyswitch_lut_kbench.log): loads in flight per thread are what counts
        if (parts > 1) { if (c->ks_lut == 2) PEBA1_KS_LUT(2, true, 1, 4); else PEBA1_KS_LUT(4, true, 1, 8); }
        else if (c->ks_lut == 2) PEBA1_KS_LUT(2, false, 1, 4);
        else if (LG == 4) PEBA1_KS_LUT(4, false, 4, 2);
        else PEBA1_KS_LUT(4, false, 1, 8);
#undef PEBA1_KS_LUT
        CU(cudaGetLastError());
        if (c->timing) CU(cudaEventRecord(c->ev2, c->stream));
        c->stats.launches += parts > 1 ? 2 : 1;
        c->stats.keyswitches += (uint64_t)count;
        return 0;
    }
    if (parts > 1) {
        ks_init_kernel<<<(unsigned)count, 160, 0, c->stream>>>(c->p.n, c->pool_stride, u, jobs, count, pool);
        keyswitch_split_kernel<8, 2><<<(unsigned)(count * parts), kKsThreads, 0, c->stream>>>(
            c->ksk, c->p.n, c->ksk_stride, c->pool_stride, u, jobs, count, parts, pool);
        CU(cudaGetLastError());
        if (c->timing) CU(cudaEventRecord(c->ev2, c->stream));
        c->stats.launches += 2;
        c->stats.keyswitches += (uint64_t)count;
        return 0;
    }
#ifdef PEBA1_ALL_VARIANTS
    // wide level: the tiled kernel (key streamed once per tile of ciphertexts) — measured slower (profiles/r2c), VARIANTS=all
    if (c->ks_tile > 0 && count >= (int64_t)c->ks_tile_min) {
        const int parts = c->ks_tile_parts;
        const int64_t per_cta = 2 * c->ks_tile;
        const unsigned tgrid = (unsigned)(((count + per_cta - 1) / per_cta) * parts);
        if (parts > 1) ks_init_kernel<<<(unsigned)count, 160, 0, c->stream>>>(c->p.n, c->pool_stride, u, jobs, count, pool);
        switch (c->ks_tile) {
            case 8: keyswitch_tile_kernel<8, 2, 8><<<tgrid, kKsTileThreads, 0, c->stream>>>(c->ksk, c->p.n, c->ksk_stride, c->pool_stride, u, jobs, count, parts, pool); break;
            case 12: keyswitch_tile_kernel<8, 2, 12><<<tgrid, kKsTileThreads, 0, c->stream>>>(c->ksk, c->p.n, c->ksk_stride, c->pool_stride, u, jobs, count, parts, pool); break;
            default: keyswitch_tile_kernel<8, 2, 16><<<tgrid, kKsTileThreads, 0, c->stream>>>(c->ksk, c->p.n, c->ksk_stride, c->pool_stride, u, jobs, count, parts, pool); break;
        }
        CU(cudaGetLastError());
        if (c->timing) CU(cudaEventRecord(c->ev2, c->stream));
        c->stats.launches += parts > 1 ? 2 : 1;
        c->stats.keyswitches += (uint64_t)count;
        return 0;
    }
#endif
    const unsigned grid = (unsigned)((count + G - 1) / G);
    switch (G) {
#ifdef PEBA1_ALL_VARIANTS  // several ciphertexts per CTA: measured no gain
        case 2: keyswitch_kernel<8, 2, 2><<<grid, kKsThreads, 0, c->stream>>>(c->ksk, c->p.n, c->ksk_stride, c->pool_stride, u, jobs, count, pool); break;
        case 4: keyswitch_kernel<8, 2, 4><<<grid, kKsThreads, 0, c->stream>>>(c->ksk, c->p.n, c->ksk_stride, c->pool_stride, u, jobs, count, pool); break;
        case 8: keyswitch_kernel<8, 2, 8><<<grid, kKsThreads, 0, c->stream>>>(c->ksk, c->p.n, c->ksk_stride, c->pool_stride, u, jobs, count, pool); break;
        case 6: keyswitch_kernel<8, 2, 6><<<grid, kKsThreads, 0, c->stream>>>(c->ksk, c->p.n, c->ksk_stride, c->pool_stride, u, jobs, count, pool); break;
#endif
        default: keyswitch_kernel<8, 2, 1><<<grid, kKsThreads, 0, c->stream>>>(c->ksk, c->p.n, c->ksk_stride, c->pool_stride, u, jobs, count, pool); break;
    }
    CU(cudaGetLastError());
    if (c->timing) CU(cudaEventRecord(c->ev2, c->stream));
    c->stats.launches++;
    c->stats.keyswitches += (uint64_t)count;
    return 0;
}

int collect_timing(peba1_cuda_ctx* c, bool pbs, bool ks) {
    if (!c->timing) return 0;
    CU(cudaStreamSynchronize(c->stream));
    if (pbs) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        c->stats.last_blind_rotate_ms = ms;
        c->stats.sum_blind_rotate_ms += ms;
    }
    if (ks) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, c->ev1, c->ev2));
        c->stats.last_keyswitch_ms = ms;
        c->stats.sum_keyswitch_ms += ms;
    }
    return 0;
}

// gate constants: c0 in eighths of the torus, sa, sb
bool gate_consts(int op, int* c0, int* sa, int* sb) {
    switch (op) {
        case PEBA1_OP_NAND: *c0 = 1; *sa = -1; *sb = -1; return true;
        case PEBA1_OP_AND: *c0 = -1; *sa = 1; *sb = 1; return true;
        case PEBA1_OP_OR: *c0 = 1; *sa = 1; *sb = 1; return true;
        case PEBA1_OP_XOR: *c0 = 2; *sa = 2; *sb = 2; return true;
        case PEBA1_OP_XNOR: *c0 = -2; *sa = -2; *sb = -2; return true;
        case PEBA1_OP_NOR: *c0 = -1; *sa = -1; *sb = -1; return true;
        case PEBA1_OP_ANDNY: *c0 = -1; *sa = -1; *sb = 1; return true;
        case PEBA1_OP_ANDYN: *c0 = -1; *sa = 1; *sb = -1; return true;
        case PEBA1_OP_ORNY: *c0 = 1; *sa = -1; *sb = 1; return true;
        case PEBA1_OP_ORYN: *c0 = 1; *sa = 1; *sb = -1; return true;
        default: return false;
    }
}
constexpr uint32_t kEighth = 1u << 29;

// operand q of gate g with multiplier s: fills (slot, sign) and returns the constant contribution to b.
inline uint32_t operand(const peba1_gate& g, int q, int s, int32_t* slot, int16_t* sign) {
    if (g.flags & PEBA1_IN_CONST(q)) {
        *slot = -1;
        *sign = 0;
        uint32_t mu = g.in[q] ? kEighth : 0u - kEighth;
        if (g.flags & PEBA1_IN_NEG(q)) mu = 0u - mu;
        return (uint32_t)s * mu;
    }
    *slot = (int32_t)g.in[q];
    *sign = (int16_t)((g.flags & PEBA1_IN_NEG(q)) ? -s : s);
    return 0;
}

}  // namespace

extern "C" {

int peba1_cuda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int peba1_cuda_create(int device, const peba1_cuda_params* params, peba1_cuda_ctx** out) {
    if (!params || !out) return fail(PEBA1_EINVAL, "null argument");
    *out = nullptr;
    if (!params_supported(*params))
        return fail(PEBA1_EINVAL, "unsupported parameter set (need N=1024, k=1, (l,Bgbit) in {(3,7),(2,10)}, ks t=8 base 4)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(PEBA1_ENODEVICE, "no CUDA device (libpeba1_cuda has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(PEBA1_EINVAL, "device %d out of range (%d devices)", device, ndev);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(PEBA1_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    CU(cudaSetDevice(device));
    peba1_cuda_ctx* c = new (std::nothrow) peba1_cuda_ctx();
    if (!c) return fail(PEBA1_ENOMEM, "out of host memory");
    struct Guard {  // a failing step below must not leak the context, its stream, events and buffers
        peba1_cuda_ctx* c;
        ~Guard() { if (c) { char keep[sizeof g_err]; memcpy(keep, g_err, sizeof keep); peba1_cuda_destroy(c); memcpy(g_err, keep, sizeof keep); } }
    } guard{c};
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->p = *params;
    c->rows = 2 * params->l;
    c->pool_stride = ((int64_t)params->n + 1 + kPoolPad - 1) / kPoolPad * kPoolPad;
    if (const char* v = getenv("PEBA1_KERNEL_VARIANT")) peba1_cuda_set_option(c, "kernel_variant", atoi(v));  // tuning aid
    CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CU(cudaEventCreateWithFlags(&c->stream_switch, cudaEventDisableTiming));
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    CU(cudaEventCreate(&c->ev2));
    static TwiddleTables tables;  // identical for every context
    build_twiddle_tables(tables);
    CU(cudaMemcpyToSymbol(c_tables, &tables, sizeof tables));
    c->bkfft_bytes = (size_t)params->n * c->rows * 2 * kM * sizeof(double2);
    // key rows start on 128-byte lines: a warp's 512-byte segment then covers 4 lines instead of straddling 5
    c->ksk_stride = ((int64_t)params->n + 1 + 31) / 32 * 32;
    c->ksk_bytes = (size_t)kN * params->ks_t * ((1 << params->ks_basebit) - 1) * c->ksk_stride * sizeof(int32_t);
    CU(cudaMalloc(&c->bkfft, c->bkfft_bytes));
    CU(cudaMalloc(&c->ksk, c->ksk_bytes));
    guard.c = nullptr;
    *out = c;
    return 0;
}

int peba1_cuda_destroy(peba1_cuda_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->bkfft); cudaFree(c->ksk); cudaFree(c->ksk_lut); cudaFree(c->pool); cudaFree(c->u); cudaFree(c->d_jobs); cudaFree(c->d_job_counter); cudaFree(c->d_stage);
    cudaFree(c->hb_pool);
    for (cudaEvent_t e : c->log_ev) cudaEventDestroy(e);
    if (c->stream_switch) cudaEventDestroy(c->stream_switch);
    for (int b = 0; b < 2; ++b) {
        cudaFreeHost(c->h_jobs[b]);
        if (c->jobs_copied[b]) cudaEventDestroy(c->jobs_copied[b]);
    }
    cudaFreeHost(c->h_stage);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev2) cudaEventDestroy(c->ev2);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return 0;
}

int peba1_cuda_set_stream(peba1_cuda_ctx* c, void* s) {
    if (int rc = use_device(c)) return rc;
    cudaStream_t next = s ? (cudaStream_t)s : c->own_stream;
    if (next == c->stream) return 0;
    // work queued on the old stream (launches that still read the job lists, the workspaces, the pool) stays ordered
    // before whatever the new stream is given
    CU(cudaEventRecord(c->stream_switch, c->stream));
    CU(cudaStreamWaitEvent(next, c->stream_switch, 0));
    c->stream = next;
    return 0;
}

int peba1_cuda_sync(peba1_cuda_ctx* c) {
    if (int rc = use_device(c)) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int peba1_cuda_load_bootstrap_key(peba1_cuda_ctx* c, const int32_t* bk_host) {
    if (int rc = use_device(c)) return rc;
    if (!bk_host) return fail(PEBA1_EINVAL, "null key");
    const int64_t polys = (int64_t)c->p.n * c->rows * 2;
    const size_t bytes = (size_t)polys * kN * sizeof(int32_t);
    DevBuf buf;
    CU(cudaMalloc(&buf.p, bytes));
    int32_t* d_bk = buf.as<int32_t>();
    CU(cudaMemcpyAsync(d_bk, bk_host, bytes, cudaMemcpyHostToDevice, c->stream));
    constexpr int TEAMS = 4;
    auto k = key_transform_kernel<TEAMS>;
    CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)team_smem_bytes(TEAMS)));
    k<<<(unsigned)((polys + TEAMS - 1) / TEAMS), TEAMS * kTeam, team_smem_bytes(TEAMS), c->stream>>>(d_bk, c->rows, polys, c->bkfft);
    CU(cudaGetLastError());
    c->stats.launches++;
    CU(cudaStreamSynchronize(c->stream));
    c->bk_loaded = true;
    return 0;
}

int peba1_cuda_load_keyswitch_key(peba1_cuda_ctx* c, const int32_t* ksk_host) {
    if (int rc = use_device(c)) return rc;
    if (!ksk_host) return fail(PEBA1_EINVAL, "null key");
    const int base = 1 << c->p.ks_basebit;
    const size_t words = (size_t)kN * c->p.ks_t * base * (c->p.n + 1);
    DevBuf buf;
    CU(cudaMalloc(&buf.p, words * sizeof(int32_t)));
    int32_t* d_src = buf.as<int32_t>();
    CU(cudaMemcpyAsync(d_src, ksk_host, words * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    const int64_t rows_out = (int64_t)kN * c->p.ks_t * (base - 1);
    ksk_pack_kernel<<<(unsigned)rows_out, 128, 0, c->stream>>>(d_src, c->p.n, base, c->ksk_stride, rows_out, c->ksk);
    CU(cudaGetLastError());
    c->stats.launches++;
    CU(cudaStreamSynchronize(c->stream));
    c->ksk_loaded = true;
    c->ksk_lut_digits = 0;  // the row sums belong to the previous key
    return 0;
}

int peba1_cuda_key_device_buffers(peba1_cuda_ctx* c, void** bkfft, uint64_t* bkfft_bytes, void** ksk, uint64_t* ksk_bytes) {
    if (!c) return fail(PEBA1_EINVAL, "null context");
    if (bkfft) *bkfft = c->bkfft;
    if (bkfft_bytes) *bkfft_bytes = c->bkfft_bytes;
    if (ksk) *ksk = c->ksk;
    if (ksk_bytes) *ksk_bytes = c->ksk_bytes;
    return 0;
}

int peba1_cuda_keys_mark_loaded(peba1_cuda_ctx* c) {
    if (!c) return fail(PEBA1_EINVAL, "null context");
    c->bk_loaded = c->ksk_loaded = true;
    c->ksk_lut_digits = 0;  // the caller wrote a key into the device buffers: row sums are rebuilt on next use
    return 0;
}

int peba1_cuda_pool_reserve(peba1_cuda_ctx* c, int64_t slots) {
    if (int rc = use_device(c)) return rc;
    if (slots <= c->pool_slots) return 0;
    if (slots > ((int64_t)1 << 31) - 1) return fail(PEBA1_EINVAL, "pool limited to 2^31-1 slots");
    int32_t* np = nullptr;
    const size_t bytes = (size_t)slots * c->pool_stride * sizeof(int32_t);
    if (cudaMalloc(&np, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(PEBA1_ENOMEM, "cannot allocate a pool of %lld slots (%zu bytes)", (long long)slots, bytes);
    }
    if (c->pool_slots > 0)
        CU(cudaMemcpyAsync(np, c->pool, (size_t)c->pool_slots * c->pool_stride * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
    pool_init_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(np, c->pool_stride, c->p.n, c->pool_slots, slots - c->pool_slots);
    CU(cudaGetLastError());
    c->stats.launches++;
    CU(cudaStreamSynchronize(c->stream));
    if (c->pool) CU(cudaFree(c->pool));
    c->pool = np;
    c->pool_slots = slots;
    return 0;
}

int peba1_cuda_pool_info(peba1_cuda_ctx* c, void** device_ptr, int64_t* stride_words, int64_t* capacity) {
    if (!c) return fail(PEBA1_EINVAL, "null context");
    if (device_ptr) *device_ptr = c->pool;
    if (stride_words) *stride_words = c->pool_stride;
    if (capacity) *capacity = c->pool_slots;
    return 0;
}

// grows the private rows of the host-buffer entry points (contents need not survive: every call rewrites what it reads)
static int ensure_hb_rows(peba1_cuda_ctx* c, int64_t slots) {
    if (slots <= c->hb_slots) return 0;
    CU(cudaStreamSynchronize(c->stream));
    if (c->hb_pool) CU(cudaFree(c->hb_pool));
    c->hb_pool = nullptr;
    c->hb_slots = 0;
    if (cudaMalloc(&c->hb_pool, (size_t)slots * c->pool_stride * sizeof(int32_t)) != cudaSuccess) {
        cudaGetLastError();
        return fail(PEBA1_ENOMEM, "cannot allocate %lld staging rows for a host-buffer call", (long long)slots);
    }
    c->hb_slots = slots;
    return 0;
}

static int check_slots(int64_t pool_slots, const int64_t* slots, int64_t first, int64_t count) {
    if (count < 0) return fail(PEBA1_EINVAL, "negative count");
    if (slots) {
        for (int64_t i = 0; i < count; ++i)
            if (slots[i] < 0 || slots[i] >= pool_slots) return fail(PEBA1_EINVAL, "slot %lld outside the pool (%lld slots)", (long long)slots[i], (long long)pool_slots);
    } else if (first < 0 || first + count > pool_slots) {
        return fail(PEBA1_EINVAL, "slot range [%lld, %lld) outside the pool (%lld slots)", (long long)first, (long long)(first + count), (long long)pool_slots);
    }
    return 0;
}

// host <-> rows of `pool` (the ciphertext pool, or the private rows of the host-buffer entry points)
static int transfer(peba1_cuda_ctx* c, int32_t* pool, int64_t pool_slots, const int64_t* slots, int64_t first, int64_t count,
                    const int32_t* h_in, int32_t* h_out) {
    if (int rc = use_device(c)) return rc;
    if (int rc = check_slots(pool_slots, slots, first, count)) return rc;
    if (count == 0) return 0;
    if ((h_in == nullptr) == (h_out == nullptr)) return fail(PEBA1_EINVAL, "null host buffer");
    const int w = c->p.n + 1;
    const int64_t chunk = 1 << 16;  // samples per staging round (165 MB)
    for (int64_t off = 0; off < count; off += chunk) {
        const int64_t m = std::min(chunk, count - off);
        const size_t bytes = (size_t)m * w * sizeof(int32_t);
        if (int rc = ensure_dev((void**)&c->d_stage, &c->d_stage_cap, bytes)) return rc;
        int64_t* d_slots = nullptr;
        if (slots) {
            if (int rc = ensure_dev(&c->d_jobs, &c->d_jobs_cap, (size_t)m * sizeof(int64_t))) return rc;
            d_slots = (int64_t*)c->d_jobs;
            CU(cudaMemcpyAsync(d_slots, slots + off, (size_t)m * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        }
        if (h_in) {
            CU(cudaMemcpyAsync(c->d_stage, h_in + (size_t)off * w, bytes, cudaMemcpyHostToDevice, c->stream));
            scatter_kernel<<<(unsigned)m, 128, 0, c->stream>>>(pool, c->pool_stride, c->p.n, d_slots, first + off, c->d_stage);
        } else {
            gather_kernel<<<(unsigned)m, 128, 0, c->stream>>>(pool, c->pool_stride, c->p.n, d_slots, first + off, c->d_stage);
            CU(cudaMemcpyAsync(h_out + (size_t)off * w, c->d_stage, bytes, cudaMemcpyDeviceToHost, c->stream));
        }
        CU(cudaGetLastError());
        c->stats.launches++;
        CU(cudaStreamSynchronize(c->stream));  // staging buffers are reused; host buffers may be pageable
    }
    return 0;
}

int peba1_cuda_upload(peba1_cuda_ctx* c, int64_t first, int64_t count, const int32_t* host) {
    if (!c) return fail(PEBA1_EINVAL, "null context");
    return transfer(c, c->pool, c->pool_slots, nullptr, first, count, host, nullptr);
}
int peba1_cuda_download(peba1_cuda_ctx* c, int64_t first, int64_t count, int32_t* host) {
    if (!c) return fail(PEBA1_EINVAL, "null context");
    return transfer(c, c->pool, c->pool_slots, nullptr, first, count, nullptr, host);
}
int peba1_cuda_upload_slots(peba1_cuda_ctx* c, const int64_t* slots, int64_t count, const int32_t* host) {
    if (!c) return fail(PEBA1_EINVAL, "null context");
    if (!slots) return fail(PEBA1_EINVAL, "null slots");
    return transfer(c, c->pool, c->pool_slots, slots, 0, count, host, nullptr);
}
int peba1_cuda_download_slots(peba1_cuda_ctx* c, const int64_t* slots, int64_t count, int32_t* host) {
    if (!c) return fail(PEBA1_EINVAL, "null context");
    if (!slots) return fail(PEBA1_EINVAL, "null slots");
    return transfer(c, c->pool, c->pool_slots, slots, 0, count, nullptr, host);
}
// slots[i] <- the i-th dense sample (n+1 words) of a DEVICE buffer, e.g. what an NCCL broadcast / scatter delivered:
// no bounce through host memory.  Asynchronous on the context stream; `d_words` must stay valid until it has run.
int peba1_cuda_upload_slots_device(peba1_cuda_ctx* c, const int64_t* slots, int64_t first_slot, int64_t count, const int32_t* d_words) {
    if (int rc = use_device(c)) return rc;
    if (int rc = check_slots(c->pool_slots, slots, first_slot, count)) return rc;
    if (count == 0) return 0;
    if (!d_words) return fail(PEBA1_EINVAL, "null device buffer");
    int64_t* d_slots = nullptr;
    if (slots) {
        CU(cudaStreamSynchronize(c->stream));  // d_jobs may still be read by queued launches
        if (int rc = ensure_dev(&c->d_jobs, &c->d_jobs_cap, (size_t)count * sizeof(int64_t))) return rc;
        d_slots = (int64_t*)c->d_jobs;
        CU(cudaMemcpyAsync(d_slots, slots, (size_t)count * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    }
    scatter_kernel<<<(unsigned)count, 128, 0, c->stream>>>(c->pool, c->pool_stride, c->p.n, d_slots, first_slot, d_words);
    CU(cudaGetLastError());
    c->stats.launches++;
    if (slots) CU(cudaStreamSynchronize(c->stream));  // the pageable slot list has been consumed
    return 0;
}

// one circuit level on the rows of `pool` (the ciphertext pool, or the private rows of the host-buffer entry points)
static int run_gates(peba1_cuda_ctx* c, int32_t* pool, int64_t pool_slots, const peba1_gate* gates, int64_t count) {
    if (int rc = use_device(c)) return rc;
    if (count < 0 || (count > 0 && !gates)) return fail(PEBA1_EINVAL, "bad gate list");
    if (!c->bk_loaded || !c->ksk_loaded) return fail(PEBA1_ENOKEY, "keys are not loaded");
    // validate everything before launching anything
    for (int64_t i = 0; i < count; ++i) {
        const peba1_gate& g = gates[i];
        if (g.op < 0 || g.op > PEBA1_OP_MAJ3) return fail(PEBA1_EINVAL, "gate %lld: unknown op %d", (long long)i, g.op);
        const int nin = g.op >= PEBA1_OP_MUX ? 3 : 2;
        for (int q = 0; q < nin; ++q) {
            if (g.flags & PEBA1_IN_CONST(q)) {
                if (g.in[q] != 0 && g.in[q] != 1) return fail(PEBA1_EINVAL, "gate %lld: constant operand must be 0 or 1", (long long)i);
            } else if (g.in[q] < 0 || g.in[q] >= pool_slots) {
                return fail(PEBA1_EINVAL, "gate %lld: input slot %lld outside the pool (%lld slots)", (long long)i, (long long)g.in[q], (long long)pool_slots);
            }
        }
        if (g.out < 0 || g.out >= pool_slots) return fail(PEBA1_EINVAL, "gate %lld: output slot %lld outside the pool", (long long)i, (long long)g.out);
    }
    // Level semantics (every bootstrap reads before any keyswitch writes) hold inside ONE launch pair.  A level larger than
    // max_pbs_per_launch runs as several pairs in sequence: an output slot that a LATER chunk reads would then hand
    // that chunk the new value, and the result would depend on the chunk size.  Refuse instead of computing it.
    {
        int64_t total_pbs = 0;
        for (int64_t i = 0; i < count; ++i) total_pbs += gates[i].op == PEBA1_OP_MUX ? 2 : 1;
        if (total_pbs > c->max_pbs_per_launch) {
            std::vector<int64_t> outs;
            outs.reserve((size_t)count);
            int64_t npbs = 0, chunk_start = 0;
            std::vector<int64_t> earlier;  // sorted outputs of the chunks before the current one
            for (int64_t i = 0; i < count; ++i) {
                const peba1_gate& g = gates[i];
                const int need = g.op == PEBA1_OP_MUX ? 2 : 1;
                if (i > chunk_start && npbs + need > c->max_pbs_per_launch) {
                    earlier.insert(earlier.end(), outs.begin(), outs.end());
                    std::sort(earlier.begin(), earlier.end());
                    outs.clear();
                    npbs = 0;
                    chunk_start = i;
                }
                npbs += need;
                outs.push_back(g.out);
                if (!earlier.empty()) {
                    const int nin = g.op >= PEBA1_OP_MUX ? 3 : 2;
                    for (int q = 0; q < nin; ++q)
                        if (!(g.flags & PEBA1_IN_CONST(q)) && std::binary_search(earlier.begin(), earlier.end(), g.in[q]))
                            return fail(PEBA1_EINVAL, "gate %lld reads slot %lld, which an earlier chunk of this level (more than "
                                        "max_pbs_per_launch = %lld bootstraps) writes: in-place levels must fit one launch",
                                        (long long)i, (long long)g.in[q], (long long)c->max_pbs_per_launch);
                }
            }
        }
    }
    int64_t done = 0;
    while (done < count) {
        // take gates until the PBS budget of one launch is reached
        int64_t npbs = 0, m = 0;
        while (done + m < count) {
            const int need = gates[done + m].op == PEBA1_OP_MUX ? 2 : 1;
            if (m > 0 && npbs + need > c->max_pbs_per_launch) break;
            npbs += need;
            ++m;
        }
        const size_t jb = (size_t)npbs * sizeof(PbsJob), kb = (size_t)m * sizeof(KsJob);
        const size_t kb_off = (jb + 255) / 256 * 256;
        const int hb = c->jobs_next;
        c->jobs_next ^= 1;
        if (c->jobs_pending[hb]) {  // the copy that read this pinned list two chunks ago
            CU(cudaEventSynchronize(c->jobs_copied[hb]));
            c->jobs_pending[hb] = false;
        }
        if (!c->jobs_copied[hb]) CU(cudaEventCreateWithFlags(&c->jobs_copied[hb], cudaEventDisableTiming));
        if (int rc = ensure_pinned(&c->h_jobs[hb], &c->h_jobs_cap[hb], kb_off + kb)) return rc;
        if (int rc = ensure_dev(&c->d_jobs, &c->d_jobs_cap, kb_off + kb)) return rc;
        if (c->u_cap < npbs) {
            if (c->u) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->u)); c->u = nullptr; c->u_cap = 0; }
            const int64_t want = std::max<int64_t>(npbs, 4096);
            if (cudaMalloc(&c->u, (size_t)want * kUStride * sizeof(int32_t)) != cudaSuccess) {
                cudaGetLastError();
                return fail(PEBA1_ENOMEM, "cannot allocate the extracted-sample workspace for %lld bootstraps", (long long)want);
            }
            c->u_cap = want;
        }
        PbsJob* pj = (PbsJob*)c->h_jobs[hb];
        KsJob* kj = (KsJob*)((char*)c->h_jobs[hb] + kb_off);
        int64_t ip = 0;
        for (int64_t i = 0; i < m; ++i) {
            const peba1_gate& g = gates[done + i];
            KsJob& k = kj[i];
            k.pad = 0;
            k.out_slot = g.out;
            if (g.op == PEBA1_OP_MUX) {
                // u1 = PBS(-1/8 + a + b), u2 = PBS(-1/8 - a + c), out = KS((0,1/8) + u1 + u2)
                PbsJob& j1 = pj[ip];
                PbsJob& j2 = pj[ip + 1];
                j1 = PbsJob{};
                j2 = PbsJob{};
                uint32_t c1 = 0u - kEighth, c2 = 0u - kEighth;
                c1 += operand(g, 0, 1, &j1.in0, &j1.s0);
                c1 += operand(g, 1, 1, &j1.in1, &j1.s1);
                c2 += operand(g, 0, -1, &j2.in0, &j2.s0);
                c2 += operand(g, 2, 1, &j2.in1, &j2.s1);
                j1.c0 = (int32_t)c1;
                j2.c0 = (int32_t)c2;
                k.u0 = (int32_t)ip;
                k.u1 = (int32_t)ip + 1;
                k.c0 = (int32_t)kEighth;
                ip += 2;
            } else if (g.op == PEBA1_OP_XOR3 || g.op == PEBA1_OP_MAJ3) {
                // XOR3: 2(a+b+c) + 1/2 (phases -1/4, +1/4, -1/4, +1/4 for 0..3 ones); MAJ3: a+b+c (-3/8, -1/8, +1/8, +3/8)
                const int s = g.op == PEBA1_OP_XOR3 ? 2 : 1;
                PbsJob& j = pj[ip];
                j = PbsJob{};
                uint32_t cc = g.op == PEBA1_OP_XOR3 ? 4u * kEighth : 0u;
                cc += operand(g, 0, s, &j.in0, &j.s0);
                cc += operand(g, 1, s, &j.in1, &j.s1);
                cc += operand(g, 2, s, &j.in2, &j.s2);
                j.c0 = (int32_t)cc;
                k.u0 = (int32_t)ip;
                k.u1 = -1;
                k.c0 = 0;
                ip += 1;
            } else {
                int c0, sa, sb;
                gate_consts(g.op, &c0, &sa, &sb);
                PbsJob& j = pj[ip];
                j = PbsJob{};
                uint32_t cc = (uint32_t)c0 * kEighth;
                cc += operand(g, 0, sa, &j.in0, &j.s0);
                cc += operand(g, 1, sb, &j.in1, &j.s1);
                j.c0 = (int32_t)cc;
                k.u0 = (int32_t)ip;
                k.u1 = -1;
                k.c0 = 0;
                ip += 1;
            }
        }
        CU(cudaMemcpyAsync(c->d_jobs, c->h_jobs[hb], kb_off + kb, cudaMemcpyHostToDevice, c->stream));
        CU(cudaEventRecord(c->jobs_copied[hb], c->stream));
        c->jobs_pending[hb] = true;
        cudaEvent_t* lev = nullptr;
        if (c->log_on) {
            while (c->log_ev.size() < 3 * (c->log_n + 1)) {
                cudaEvent_t e;
                CU(cudaEventCreate(&e));
                c->log_ev.push_back(e);
            }
            lev = &c->log_ev[3 * c->log_n];
            c->log_gates.push_back(m);
            c->log_pbs.push_back(npbs);
            c->log_n++;
            CU(cudaEventRecord(lev[0], c->stream));
        }
        if (int rc = launch_pbs(c, pool, c->pool_stride, (const PbsJob*)c->d_jobs, npbs, c->u)) return rc;
        if (lev) CU(cudaEventRecord(lev[1], c->stream));
        if (int rc = launch_ks(c, c->u, (const KsJob*)((char*)c->d_jobs + kb_off), m, pool)) return rc;
        if (lev) CU(cudaEventRecord(lev[2], c->stream));
        if (int rc = collect_timing(c, true, true)) return rc;
        done += m;
    }
    return 0;  // asynchronous on the context stream (peba1_cuda_sync, or any download, waits)
}

int peba1_cuda_gates(peba1_cuda_ctx* c, const peba1_gate* gates, int64_t count) {
    if (!c) return fail(PEBA1_EINVAL, "null context");
    return run_gates(c, c->pool, c->pool_slots, gates, count);
}

// Per-level device times without a sync per level: between _begin and _end every launch pair is bracketed by events on
// the context stream; _end waits for the stream and resolves them.
int peba1_cuda_level_log_begin(peba1_cuda_ctx* c) {
    if (!c) return fail(PEBA1_EINVAL, "null context");
    c->log_on = true;
    c->log_n = 0;
    c->log_gates.clear();
    c->log_pbs.clear();
    return 0;
}
int peba1_cuda_level_log_end(peba1_cuda_ctx* c, int64_t* gates, int64_t* pbs, float* blind_rotate_ms, float* keyswitch_ms,
                             int64_t capacity, int64_t* entries) {
    if (int rc = use_device(c)) return rc;
    c->log_on = false;
    CU(cudaStreamSynchronize(c->stream));
    if (entries) *entries = (int64_t)c->log_n;
    for (size_t i = 0; i < c->log_n && (int64_t)i < capacity; ++i) {
        if (gates) gates[i] = c->log_gates[i];
        if (pbs) pbs[i] = c->log_pbs[i];
        float a = 0, b = 0;
        CU(cudaEventElapsedTime(&a, c->log_ev[3 * i], c->log_ev[3 * i + 1]));
        CU(cudaEventElapsedTime(&b, c->log_ev[3 * i + 1], c->log_ev[3 * i + 2]));
        if (blind_rotate_ms) blind_rotate_ms[i] = a;
        if (keyswitch_ms) keyswitch_ms[i] = b;
    }
    return 0;
}

int peba1_cuda_gate_batch(peba1_cuda_ctx* c, int32_t op, const int64_t* in0, const int64_t* in1, const int64_t* in2, const int64_t* out, int64_t count) {
    if (count < 0 || (count > 0 && (!in0 || !in1 || !out))) return fail(PEBA1_EINVAL, "null slot array");
    if (op >= PEBA1_OP_MUX && count > 0 && !in2) return fail(PEBA1_EINVAL, "MUX, XOR3 and MAJ3 need three operands");
    std::vector<peba1_gate> g((size_t)count);
    for (int64_t i = 0; i < count; ++i) {
        g[i].op = op;
        g[i].flags = 0;
        g[i].in[0] = in0[i];
        g[i].in[1] = in1[i];
        g[i].in[2] = in2 ? in2[i] : 0;
        g[i].out = out[i];
    }
    return peba1_cuda_gates(c, g.data(), count);
}

int peba1_cuda_gate_batch_host(peba1_cuda_ctx* c, int32_t op, const int32_t* ca, const int32_t* cb, const int32_t* cc, int32_t* out, int64_t count) {
    if (count < 0 || (count > 0 && (!ca || !cb || !out))) return fail(PEBA1_EINVAL, "null host buffer");
    const bool mux = op >= PEBA1_OP_MUX;  // the three 3-input gates
    if (mux && count > 0 && !cc) return fail(PEBA1_EINVAL, "MUX, XOR3 and MAJ3 need three operands");
    if (count == 0) return 0;
    const int nin = mux ? 3 : 2;
    if (int rc = use_device(c)) return rc;
    if (int rc = ensure_hb_rows(c, (nin + 1) * count)) return rc;  // private rows: the caller's pool slots stay untouched
    if (int rc = transfer(c, c->hb_pool, c->hb_slots, nullptr, 0, count, ca, nullptr)) return rc;
    if (int rc = transfer(c, c->hb_pool, c->hb_slots, nullptr, count, count, cb, nullptr)) return rc;
    if (mux) if (int rc = transfer(c, c->hb_pool, c->hb_slots, nullptr, 2 * count, count, cc, nullptr)) return rc;
    std::vector<peba1_gate> g((size_t)count);
    for (int64_t i = 0; i < count; ++i) {
        g[i].op = op;
        g[i].flags = 0;
        g[i].in[0] = i;
        g[i].in[1] = count + i;
        g[i].in[2] = mux ? 2 * count + i : 0;
        g[i].out = nin * count + i;
    }
    if (int rc = run_gates(c, c->hb_pool, c->hb_slots, g.data(), count)) return rc;
    return transfer(c, c->hb_pool, c->hb_slots, nullptr, nin * count, count, nullptr, out);
}

int peba1_cuda_linear(peba1_cuda_ctx* c, const int64_t* in, const int32_t* sign, const int32_t* cst, const int64_t* out, int64_t count) {
    if (int rc = use_device(c)) return rc;
    if (count < 0 || (count > 0 && (!in || !sign || !cst || !out))) return fail(PEBA1_EINVAL, "null argument");
    if (count == 0) return 0;
    for (int64_t i = 0; i < count; ++i) {
        if (in[i] >= c->pool_slots) return fail(PEBA1_EINVAL, "input slot outside the pool");
        if (out[i] < 0 || out[i] >= c->pool_slots) return fail(PEBA1_EINVAL, "output slot outside the pool");
    }
    const size_t a = (size_t)count * sizeof(int64_t), b = (size_t)count * sizeof(int32_t);
    if (int rc = ensure_dev(&c->d_jobs, &c->d_jobs_cap, 2 * a + 2 * b)) return rc;
    char* d = (char*)c->d_jobs;
    CU(cudaMemcpyAsync(d, in, a, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d + a, out, a, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d + 2 * a, sign, b, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d + 2 * a + b, cst, b, cudaMemcpyHostToDevice, c->stream));
    linear_kernel<<<(unsigned)count, 128, 0, c->stream>>>(c->pool, c->pool_stride, c->p.n, (const int64_t*)d, (const int32_t*)(d + 2 * a), (const int32_t*)(d + 2 * a + b), (const int64_t*)(d + a));
    CU(cudaGetLastError());
    c->stats.launches++;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int peba1_cuda_bootstrap_woks_host(peba1_cuda_ctx* c, const int32_t* in, int32_t* out_extracted, int64_t count) {
    if (int rc = use_device(c)) return rc;
    if (count < 0 || (count > 0 && (!in || !out_extracted))) return fail(PEBA1_EINVAL, "null host buffer");
    if (!c->bk_loaded) return fail(PEBA1_ENOKEY, "bootstrapping key is not loaded");
    if (count == 0) return 0;
    const int w = c->p.n + 1;
    DevBuf b_in, b_u, b_dense, b_j;
    CU(cudaMalloc(&b_in.p, (size_t)count * w * sizeof(int32_t)));
    CU(cudaMalloc(&b_u.p, (size_t)count * kUStride * sizeof(int32_t)));
    CU(cudaMalloc(&b_dense.p, (size_t)count * (kN + 1) * sizeof(int32_t)));
    CU(cudaMalloc(&b_j.p, (size_t)count * sizeof(PbsJob)));
    int32_t *d_in = b_in.as<int32_t>(), *d_u = b_u.as<int32_t>(), *d_dense = b_dense.as<int32_t>();
    PbsJob* d_j = b_j.as<PbsJob>();
    std::vector<PbsJob> jobs((size_t)count);
    for (int64_t i = 0; i < count; ++i) jobs[i] = PbsJob{(int32_t)i, -1, 0, 1, 0};
    CU(cudaMemcpyAsync(d_in, in, (size_t)count * w * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_j, jobs.data(), (size_t)count * sizeof(PbsJob), cudaMemcpyHostToDevice, c->stream));
    if (int rc = launch_pbs(c, d_in, w, d_j, count, d_u)) return rc;
    u_unpad_kernel<<<(unsigned)count, 128, 0, c->stream>>>(d_u, d_dense, 1);
    CU(cudaGetLastError());
    c->stats.launches++;
    CU(cudaMemcpyAsync(out_extracted, d_dense, (size_t)count * (kN + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int peba1_cuda_keyswitch_host(peba1_cuda_ctx* c, const int32_t* in_extracted, int32_t* out, int64_t count) {
    if (int rc = use_device(c)) return rc;
    if (count < 0 || (count > 0 && (!in_extracted || !out))) return fail(PEBA1_EINVAL, "null host buffer");
    if (!c->ksk_loaded) return fail(PEBA1_ENOKEY, "keyswitch key is not loaded");
    if (count == 0) return 0;
    if (int rc = ensure_hb_rows(c, count)) return rc;  // private rows, not the caller's pool
    DevBuf b_u, b_dense, b_j;
    CU(cudaMalloc(&b_u.p, (size_t)count * kUStride * sizeof(int32_t)));
    CU(cudaMalloc(&b_dense.p, (size_t)count * (kN + 1) * sizeof(int32_t)));
    CU(cudaMalloc(&b_j.p, (size_t)count * sizeof(KsJob)));
    int32_t *d_u = b_u.as<int32_t>(), *d_dense = b_dense.as<int32_t>();
    KsJob* d_j = b_j.as<KsJob>();
    std::vector<KsJob> jobs((size_t)count);
    for (int64_t i = 0; i < count; ++i) jobs[i] = KsJob{(int32_t)i, -1, 0, 0, i};
    CU(cudaMemcpyAsync(d_dense, in_extracted, (size_t)count * (kN + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_j, jobs.data(), (size_t)count * sizeof(KsJob), cudaMemcpyHostToDevice, c->stream));
    u_unpad_kernel<<<(unsigned)count, 128, 0, c->stream>>>(d_u, d_dense, 0);
    CU(cudaGetLastError());
    c->stats.launches++;
    if (c->timing) CU(cudaEventRecord(c->ev1, c->stream));
    if (int rc = launch_ks(c, d_u, d_j, count, c->hb_pool)) return rc;
    if (int rc = collect_timing(c, false, true)) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return transfer(c, c->hb_pool, c->hb_slots, nullptr, 0, count, nullptr, out);
}

int peba1_cuda_encrypt_bits(peba1_cuda_ctx* c, const int32_t* lwe_key, const int32_t* bits, const int64_t* slots,
                            int64_t first_slot, int64_t count, int64_t index0, uint64_t seed, double sigma) {
    if (int rc = use_device(c)) return rc;
    if (count < 0 || !lwe_key || (count > 0 && !bits)) return fail(PEBA1_EINVAL, "bad encryption request");
    if (count == 0) return 0;
    if (!slots && (first_slot < 0 || first_slot + count > c->pool_slots))
        return fail(PEBA1_EINVAL, "slot range [%lld, +%lld) outside the pool", (long long)first_slot, (long long)count);
    if (slots)
        for (int64_t i = 0; i < count; ++i)
            if (slots[i] < 0 || slots[i] >= c->pool_slots) return fail(PEBA1_EINVAL, "slot %lld outside the pool", (long long)slots[i]);
    uint32_t mask[32] = {0};  // key bit w at bit (w & 31) of word w >> 5: the kernel reads 4 bits per word quad
    for (int w = 0; w < c->p.n; ++w)
        if (lwe_key[w]) mask[w >> 5] |= 1u << (w & 31);
    const size_t bytes = sizeof mask + (size_t)count * sizeof(int32_t) + 8 + (slots ? (size_t)count * sizeof(int64_t) : 0);
    DevBuf buf;
    CU(cudaMalloc(&buf.p, bytes));
    unsigned char* d = buf.as<unsigned char>();
    CU(cudaMemcpyAsync(d, mask, sizeof mask, cudaMemcpyHostToDevice, c->stream));
    int32_t* d_bits = reinterpret_cast<int32_t*>(d + sizeof mask);
    CU(cudaMemcpyAsync(d_bits, bits, (size_t)count * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    int64_t* d_slots = nullptr;
    if (slots) {
        d_slots = reinterpret_cast<int64_t*>(d + sizeof mask + (((size_t)count * sizeof(int32_t) + 7) & ~(size_t)7));
        // the int64 array needs 8-byte alignment: the allocation is 256-byte aligned and sizeof mask is 128
        CU(cudaMemcpyAsync(d_slots, slots, (size_t)count * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    }
    const unsigned grid = (unsigned)((count + 7) / 8);
    lwe_encrypt_kernel<<<grid, 256, 0, c->stream>>>(c->pool, c->pool_stride, d_slots, first_slot, d_bits, count, index0,
                                                    c->p.n, reinterpret_cast<const uint32_t*>(d), seed, sigma);
    CU(cudaGetLastError());
    c->stats.launches++;
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int peba1_cuda_enable_timing(peba1_cuda_ctx* c, int enable) {
    if (!c) return fail(PEBA1_EINVAL, "null context");
    c->timing = enable != 0;
    return 0;
}
int peba1_cuda_get_stats(peba1_cuda_ctx* c, peba1_cuda_stats* out) {
    if (!c || !out) return fail(PEBA1_EINVAL, "null argument");
    *out = c->stats;
    return 0;
}
int peba1_cuda_reset_stats(peba1_cuda_ctx* c) {
    if (!c) return fail(PEBA1_EINVAL, "null context");
    c->stats = peba1_cuda_stats{};
    return 0;
}

int peba1_cuda_measure_fp64_peak(peba1_cuda_ctx* c, double* tflops) {
    if (int rc = use_device(c)) return rc;
    if (!tflops) return fail(PEBA1_EINVAL, "null argument");
    const int threads = 256, blocks = c->sm_count * 8, iters = 1 << 14;
    DevBuf buf;
    CU(cudaMalloc(&buf.p, (size_t)threads * blocks * sizeof(double)));
    double* d = buf.as<double>();
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(c->ev0, c->stream));
        fp64_peak_kernel<<<blocks, threads, 0, c->stream>>>(d, iters);
        CU(cudaEventRecord(c->ev1, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (rep > 0) best = std::min(best, ms);
    }
    c->stats.launches += 4;
    *tflops = 2.0 * 8 * (double)iters * threads * blocks / (best * 1e-3) / 1e12;
    return 0;
}

int peba1_cuda_set_option(peba1_cuda_ctx* c, const char* name, int64_t value) {
    if (!c || !name) return fail(PEBA1_EINVAL, "null argument");
    if (!strcmp(name, "teams_per_cta")) {
        if (value == 0) value = 6;
        if (value < 1 || value > 6) return fail(PEBA1_EINVAL, "teams_per_cta must be 1..6");
        c->teams_per_cta = (int)value;
        return 0;
    }
    if (!strcmp(name, "narrow_two_buffers")) {
        c->narrow_two_buffers = value != 0;
        return 0;
    }
    if (!strcmp(name, "tm_teams")) {
        if (value < 0 || value > 8) return fail(PEBA1_EINVAL, "tm_teams must be 0..8");
        c->tm_teams = (int)value;
        return 0;
    }
    if (!strcmp(name, "ks_group")) {
        if (value == 0) value = 1;
#ifndef PEBA1_ALL_VARIANTS
        if (value != 1) return fail(PEBA1_EINVAL, "ks_group > 1 measured no gain and is not compiled into this build (make VARIANTS=all)");
#endif
        if (value != 1 && value != 2 && value != 4 && value != 6 && value != 8) return fail(PEBA1_EINVAL, "ks_group must be 1, 2, 4, 6 or 8");
        c->ks_group = (int)value;
        return 0;
    }
    if (!strcmp(name, "ks_tile")) {
        if (value != 0 && value != 8 && value != 12 && value != 16) return fail(PEBA1_EINVAL, "ks_tile must be 0, 8, 12 or 16");
#ifndef PEBA1_ALL_VARIANTS
        if (value != 0) return fail(PEBA1_EINVAL, "the tiled keyswitch measured slower and is not compiled into this build (make VARIANTS=all)");
#endif
        c->ks_tile = (int)value;
        return 0;
    }
    if (!strcmp(name, "ks_tile_parts")) {
        if (value < 1 || value > 16 || (value & (value - 1))) return fail(PEBA1_EINVAL, "ks_tile_parts must be 1, 2, 4, 8 or 16");
        c->ks_tile_parts = (int)value;
        return 0;
    }
    if (!strcmp(name, "ks_tile_min")) {
        c->ks_tile_min = value;
        return 0;
    }
    if (!strcmp(name, "ks_lut")) {
        if (value != 0 && value != 2 && value != 4) return fail(PEBA1_EINVAL, "ks_lut must be 0 (off), 2 or 4 digits per table row");
        c->ks_lut = (int)value;
        return 0;
    }
    if (!strcmp(name, "ks_lut_group")) {
        if (value != 1 && value != 4) return fail(PEBA1_EINVAL, "ks_lut_group must be 1 or 4");
        c->ks_lut_group = (int)value;
        return 0;
    }
    if (!strcmp(name, "ks_lut_group_min")) {
        c->ks_lut_group_min = value;
        return 0;
    }
    if (!strcmp(name, "ks_lut_min")) {
        if (value < 1) return fail(PEBA1_EINVAL, "ks_lut_min must be at least 1");
        c->ks_lut_min = value;
        return 0;
    }
    if (!strcmp(name, "ks_split")) {
        if (value > 1 && (value & (value - 1))) return fail(PEBA1_EINVAL, "ks_split must be a power of two");
        if (value > 64) return fail(PEBA1_EINVAL, "ks_split must be at most 64");
        c->ks_split = value;
        return 0;
    }
    if (!strcmp(name, "wide_max_batch")) {
        c->wide_max_batch = value;
        return 0;
    }
    if (!strcmp(name, "tmk_stagger_ns")) {
        c->tmk_stagger_ns = (int)value;
        return 0;
    }
    if (!strcmp(name, "pair_min_batch")) {
        c->pair_min_batch = value;
        return 0;
    }
    if (!strcmp(name, "kernel_variant")) {
        if (value < 0 || value > 11) return fail(PEBA1_EINVAL, "kernel_variant must be 0..11");
#ifndef PEBA1_ALL_VARIANTS
        if (value != 1 && value != 8 && value != 11)
            return fail(PEBA1_EINVAL, "kernel_variant %lld lost against the default and is not compiled into this build "
                                      "(make -C peba1_b200/csrc VARIANTS=all)", (long long)value);
#endif
        c->kernel_variant = (int)value;
        return 0;
    }
    if (!strcmp(name, "max_pbs_per_launch")) {
        if (value == 0) value = 1 << 18;
        if (value < 1) return fail(PEBA1_EINVAL, "max_pbs_per_launch must be positive");
        c->max_pbs_per_launch = value;
        return 0;
    }
    return fail(PEBA1_EINVAL, "unknown option '%s'", name);
}

}  // extern "C"
