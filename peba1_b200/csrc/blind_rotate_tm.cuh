// blind_rotate_tm.cuh — blind rotation with the MAC accumulators and the per-thread twiddles in TENSOR MEMORY.
//
// Why: the first version (blind_rotate.cuh) keeps, per thread, 64 registers of frequency-domain accumulators and 64
// of twiddles, which caps the kernel at 168-255 registers = 8-12 warps per SM, and ncu shows it latency-bound (2-3
// warps per scheduler, "short scoreboard" on every shared-memory exchange).  Shared memory and L1 are the busiest
// pipe already, so spilling there is no option.  Blackwell's 256 KB tensor memory is a separate datapath that the
// kernel does not use otherwise (there is no MMA in this algorithm): tcgen05.st / tcgen05.ld in the 32x32b shape give
// every thread a private row of 32-bit cells (measured: 570 B/clk/SM load+store with 16 warps; this kernel needs
// ~140).  With accumulators and twiddles there the kernel fits 128 registers: 8 teams = 16 warps per SM.
//
// TMEM layout (512 columns allocated, one CTA per SM).  Lane quarter q = warp % 4 is fixed by the hardware; the four
// warps w, w+4, w+8, w+12 that share a quarter use different column blocks:
//   columns [64*(warp/4), +64)   accumulators of this thread: [c (2)][k2 (8)][re.lo re.hi im.lo im.hi]
//   columns [256, 288)           tw1[8] of thread tid = 32*(warp&1) + lane   (the same for all warps of a quarter)
//   columns [288, 320)           tw2[8]
//
// Same arithmetic, in the same order, as blind_rotate.cuh (bit-exact vs the exact-integer oracle); __host__
// __device__ so csrc/emu_team.cu can run it on CPU threads (tensor memory emulated by a per-thread array).
#pragma once

#include "blind_rotate.cuh"

namespace peba1 {

constexpr int kTmSums = 0;      // + 64 * (warp / 4)
constexpr int kTmTw1 = 256;
constexpr int kTmTw2 = 288;
constexpr int kTmZero = 320;    // 16 cells of 0.0: the first row of a step "loads" its accumulators from here
constexpr int kTmBk = 336;      // shared-key variant: 2 row buffers x 64 cells, [row parity][chunk q][4 complex]

// Shared memory of one team in this version: ONE cross-team exchange buffer (two barriers per transform instead of
// double buffering) so that 8 teams fit: 8 KB + 8 KB + 8 KB + 1.25 KB = 25.25 KB per team.
struct __align__(16) TeamSmemTm {
    static constexpr bool kAliased = false, kDouble = false, kShfl = false, kClear = true, kXTm = false;
    static constexpr int kTw1 = kTmTw1, kTw2 = kTmTw2, kZero = kTmZero;  // tensor-memory columns (8 teams: 4 x 64 sums)
    // [320,384): decomposition input, 16 columns per warp group; [384,512): the thread's own accumulator words, 32 per
    // warp group (see blind_rotate_team_tm).  The zero block of the other layouts is not needed: the sums are cleared.
    static constexpr int kTv = 320, kOwn = 384, kX = 0;
    double2  ex1[kM];
    double2  ex2[kM];
    int32_t  acc[2][kN];
    uint16_t bara[640];
    unsigned long long ex1_free;  // mbarrier: "the cross-team exchange buffer may be overwritten" (2 arrivals, one per warp)
    unsigned long long pad_;
    PEBA1_HD double2* e1(int) { return ex1; }
    PEBA1_HD double2* e2(int) { return ex2; }
};

// The same with ONE exchange buffer: 17.3 KB per team instead of 25.3 KB, so 10 teams (20 warps, 96 registers) fit one SM
// instead of 8.  The inner exchange may live in the cross-team buffer because thread (hi, lo) only ever touches region
// [64 hi, 64 hi + 64) there, in both exchanges, and the eight threads of a region sit in one warp: the hazards between
// the two uses are warp-local (__syncwarp), only the "buffer free" signal moves to after the LAST read of a transform.
struct __align__(16) TeamSmemTm1 {
    static constexpr bool kAliased = true, kDouble = false, kShfl = false, kClear = false, kXTm = false;
    static constexpr int kTw1 = 320, kTw2 = 352, kZero = 384;  // 10 teams: 5 x 64 columns of sums come first
    static constexpr int kTv = 400, kOwn = -1, kX = 0;   // no columns left for the own-words copy
    double2  ex1[kM];
    int32_t  acc[2][kN];
    uint16_t bara[640];
    unsigned long long ex1_free;
    unsigned long long pad_;
    PEBA1_HD double2* e1(int) { return ex1; }
    PEBA1_HD double2* e2(int) { return ex1; }
};

// Both exchanges of a transform in one buffer (as TeamSmemTm1), and TWO such buffers used alternately by successive
// transforms: 25.3 KB per team like TeamSmemTm, but no "buffer free" mbarrier at all.  A buffer is rewritten two
// transforms after it was read, and the team barrier of the transform in between separates the two: a thread passes that
// barrier only when every thread of the team has arrived at it, i.e. has finished the reads of the transform before.
struct __align__(16) TeamSmemTm2 {
    static constexpr bool kAliased = true, kDouble = true, kShfl = false, kClear = true, kXTm = false;
    static constexpr int kTw1 = kTmTw1, kTw2 = kTmTw2, kZero = kTmZero;
    static constexpr int kTv = 320, kOwn = 384, kX = 0;
    double2  ex[2][kM];
    int32_t  acc[2][kN];
    uint16_t bara[640];
    unsigned long long ex1_free;  // unused (kept so that the kernel's set-up code is the same for all layouts)
    unsigned long long pad_;
    PEBA1_HD double2* e1(int par) { return ex[par]; }
    PEBA1_HD double2* e2(int par) { return ex[par]; }
};

// TeamSmemTm2 with the INNER exchange of a transform (the 8 x 8 transposition between the lane index inside an 8-lane
// group and the register index) done by warp shuffles instead of shared memory: three butterfly stages, each swapping one
// bit between the two indices with __shfl_xor (north_star: "warp-shuffle butterflies").  A complex double is four 32-bit
// shuffles, and every stage moves half of a thread's 8 values: 48 SHFL + the selects that pick what to send and where to
// put what arrives (a lane-dependent choice between two registers, per value), against 8 STS.128 + 8 LDS.128.  Same
// arithmetic, same words.  kernel_variant 10; not measured yet (written after the round's GPU budget was spent).
struct __align__(16) TeamSmemTm2S : TeamSmemTm2 {
    static constexpr bool kShfl = true;
};

// The CROSS-TEAM exchange of every transform through TENSOR MEMORY instead of shared memory (kXTm; kernel_variant 11).
// The shared-memory / L1 data pipe is this kernel's busiest unit, and the cross-team exchange is a third of its traffic
// (8 STS.128 + 8 LDS.128 per thread and transform).  Tensor memory is a separate datapath: a thread stores its 8 complex
// values with tcgen05.st.32x32b (row = lane, 32 columns) and reads with tcgen05.ld.16x256b, whose fragment layout hands
// thread t the 64-bit units (t & 3) of lanes (t >> 2) and (t >> 2) + 8 — a transposition between two lane bits and the
// unit index that never touches the LSU (layout probed on the device: scripts/ubench/tmem_xchg_probe.cu).
//
// For that, the two warps of a team must share a tensor-memory lane quarter (warp % 4): a team is warps (w, w + 4), thread
// tid = 32 wb + lane.  Exchange 1 moves register index k0 <-> thread bits (wb, L4, L3) = b:
//     store   thread (wb, lane), value k0  ->  own block, column 16 (k0 >> 2) + 8 isIm + 2 (k0 & 3) + word
//     load    warp wb' reads columns [16 wb', +16) of BOTH blocks (S = 0, 1), lane halves H = 0, 1: thread t receives
//             b = 4 S + 2 H + b0 of  m = 8 b + (t >> 2),  k0 = 4 wb' + (t & 3)
// so AFTER the exchange physical thread (wb', t) is the logical thread hi = k0 = 4 wb' + (t & 3), lo = c = t >> 2 of the
// radix-8 network for the SECOND pass only: the inner exchange, which stays in shared memory and is warp-local, is written
// in that permuted position and read back by the plain thread (tid >> 3, tid & 7), so transform outputs, key columns and
// sums sit where they sit in every other layout (a permuted position there would un-coalesce the key loads: 8
// consecutive lanes would touch 4 different 128-byte lines).  Its swizzle gains the term 2 (h & 3): a 16-byte access is
// served 8 lanes at a time and 8 consecutive lanes of the permuted side hold 4 different h.  The inverse transform runs
// both trips backwards (tcgen05.st.16x256b into both blocks, tcgen05.ld.32x32b of the own block).
// Tensor-memory columns: [0,256) sums, [256,320) tw1 of wb = 0 / 1, [320,352) tw2, [352,480) exchange (32 per warp).
// The decomposition input and the thread-own accumulator words live in shared memory again (no columns left).
struct __align__(16) TeamSmemTmX {
    static constexpr bool kAliased = true, kDouble = false, kShfl = false, kClear = true, kXTm = true;
    static constexpr int kTw1 = kTmTw1, kTw2 = 320, kZero = -1;
    static constexpr int kTv = -1, kOwn = -1, kX = 352;
    double2  ex[kM];          // inner exchange only (warp-local regions)
    int32_t  acc[2][kN];
    uint4    tv[4][kTeam];    // decomposition input between the digit levels: tv[q][tid] = words 4q..4q+3 of thread tid
    uint16_t bara[640];
    unsigned long long ex1_free;  // unused
    unsigned long long pad_;
    PEBA1_HD double2* e1(int) { return ex; }
    PEBA1_HD double2* e2(int) { return ex; }
};
// word index of value k0 (component part = 0 re.lo, 1 re.hi, 2 im.lo, 3 im.hi) in the 32-column block a thread stores
// before / loads after the 32x32b side of the trip
PEBA1_HD constexpr int xtm_col(int k0, int part) { return 16 * (k0 >> 2) + 8 * (part >> 1) + 2 * (k0 & 3) + (part & 1); }
// register index of value b on the 16x256b side: four accesses q = b >> 1 (block S = q >> 1, lane half H = q & 1) of 8 words
PEBA1_HD constexpr int xtm_reg(int b, int part) { return 8 * (b >> 1) + 4 * (part >> 1) + 2 * (b & 1) + (part & 1); }

// In-place transposition x[j] of lane l  <->  x[l] of lane j over the 8 lanes that share tid >> 3 (lo = tid & 7).
template <class Env>
PEBA1_HD void lane8_transpose(Env& env, double2 (&x)[8], int lo) {
#pragma unroll
    for (int m = 4; m >= 1; m >>= 1) {
        const bool up = (lo & m) != 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j & m) continue;
            // lanes with bit m clear keep x[j] and trade x[j | m]; their partners (bit m set) keep x[j | m] and trade x[j]
            const double2 out = up ? x[j] : x[j | m];
            const double2 in = env.shfl_xor_c(out, m);
            if (up) x[j] = in; else x[j | m] = in;
        }
    }
}

PEBA1_HD double2 words_to_c(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
#ifdef __CUDA_ARCH__
    return make_double2(__hiloint2double((int)b, (int)a), __hiloint2double((int)d, (int)c));
#else
    uint64_t re = ((uint64_t)b << 32) | a, im = ((uint64_t)d << 32) | c;
    double2 r;
    memcpy(&r.x, &re, 8);
    memcpy(&r.y, &im, 8);
    return r;
#endif
}
PEBA1_HD void c_to_words(double2 v, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
#ifdef __CUDA_ARCH__
    a = (uint32_t)__double2loint(v.x); b = (uint32_t)__double2hiint(v.x);
    c = (uint32_t)__double2loint(v.y); d = (uint32_t)__double2hiint(v.y);
#else
    uint64_t re, im;
    memcpy(&re, &v.x, 8);
    memcpy(&im, &v.y, 8);
    a = (uint32_t)re; b = (uint32_t)(re >> 32); c = (uint32_t)im; d = (uint32_t)(im >> 32);
#endif
}

// 8 complex values <-> 32 tensor-memory cells
template <class Env>
PEBA1_HD void tm_load8(Env& env, int col, double2 (&v)[8]) {
    uint32_t w[32];
    env.tmem_ld32(col, w);
    env.tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = words_to_c(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
}
template <class Env>
PEBA1_HD void tm_store8(Env& env, int col, const double2 (&v)[8]) {
    uint32_t w[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) c_to_words(v[j], w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
    env.tmem_st32(col, w);
}

// Forward transform; twiddles come from tensor memory right before their use.
template <class Env, class SM>
PEBA1_HD void team_fft_forward_tm(Env& env, double2 (&x)[8], SM& sm) {
    const int tid = env.tid;
    const int par = SM::kDouble ? env.flip() : 0;
    double2* const ex1 = sm.e1(par);
    double2* const ex2 = sm.e2(par);
    uint32_t w[32];
    if constexpr (SM::kXTm) env.tmem_ld32(env.tw1_col(), w);
    else env.tmem_ld32(SM::kTw1, w);  // in flight while the first pass computes
#pragma unroll
    for (int a = 1; a < 8; ++a) x[a] = cmul(x[a], w32(a));
    radix8<false>(x);
    env.tmem_wait_ld();
    const int hi = tid >> 3, lo = tid & 7;
    if constexpr (SM::kXTm) {
        uint32_t v[32];
#pragma unroll
        for (int k0 = 0; k0 < 8; ++k0)
            c_to_words(cmul(x[k0], words_to_c(w[4 * k0], w[4 * k0 + 1], w[4 * k0 + 2], w[4 * k0 + 3])), v[xtm_col(k0, 0)],
                       v[xtm_col(k0, 1)], v[xtm_col(k0, 2)], v[xtm_col(k0, 3)]);
        env.xtm_exchange_fwd(v);
        env.tmem_ld32(SM::kTw2, w);
#pragma unroll
        for (int b = 0; b < 8; ++b) x[b] = words_to_c(v[xtm_reg(b, 0)], v[xtm_reg(b, 1)], v[xtm_reg(b, 2)], v[xtm_reg(b, 3)]);
        radix8<false>(x);
        env.tmem_wait_ld();
        // Inner exchange.  This thread is logical (hp, lp) of the network now (see TeamSmemTmX); the READER side is the
        // plain thread (hi, lo) = (tid >> 3, tid & 7) again, so everything after the transform — key columns, sums, the
        // inverse — sees the same positions as in the other layouts.  Writers and readers of a region [64 h, +64) are
        // lanes of one warp.  cell(h, k1, c) = 64 h + 8 k1 + (c ^ k1 ^ 2 (h & 3)) is conflict-free for both lane patterns.
        const int hp = env.xhi(), lp = env.xlo(), sp = 2 * (hp & 3), ss = 2 * (hi & 3);
        env.warp_sync();  // the region's readers of the previous transform are done
        ex2[64 * hp + (lp ^ sp)] = x[0];
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1)
            ex2[64 * hp + 8 * k1 + (lp ^ k1 ^ sp)] = cmul(x[k1], words_to_c(w[4 * k1], w[4 * k1 + 1], w[4 * k1 + 2], w[4 * k1 + 3]));
        env.warp_sync();
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = ex2[64 * hi + 8 * lo + (c ^ lo ^ ss)];
        radix8<false>(x);
        return;
    } else {
        if (!SM::kDouble) env.ex1_acquire();  // every reader of the previous transform's ex1 is done (split-phase)
#pragma unroll
        for (int k0 = 0; k0 < 8; ++k0)
            ex1[64 * k0 + tid] = cmul(x[k0], words_to_c(w[4 * k0], w[4 * k0 + 1], w[4 * k0 + 2], w[4 * k0 + 3]));
        env.tmem_ld32(SM::kTw2, w);  // in flight across the barrier, the loads and the second pass
        env.team_sync();
#pragma unroll
        for (int b = 0; b < 8; ++b) x[b] = ex1[64 * hi + 8 * b + lo];
        if (!SM::kAliased && !SM::kDouble) env.ex1_release();
    }
    radix8<false>(x);
    env.tmem_wait_ld();
#if defined(PEBA1_ABLATE) && PEBA1_ABLATE == 2  /* what-if build: no inner exchange in the forward transform */
#pragma unroll
    for (int k1 = 1; k1 < 8; ++k1) x[k1] = cmul(x[k1], words_to_c(w[4 * k1], w[4 * k1 + 1], w[4 * k1 + 2], w[4 * k1 + 3]));
#else
    if (SM::kShfl) {
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) x[k1] = cmul(x[k1], words_to_c(w[4 * k1], w[4 * k1 + 1], w[4 * k1 + 2], w[4 * k1 + 3]));
        lane8_transpose(env, x, lo);
    } else {
        if (SM::kAliased) env.warp_sync();  // the region's eight readers above are done before it is rewritten
        ex2[64 * hi + lo] = x[0];
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1)
            ex2[64 * hi + 8 * k1 + (lo ^ k1)] = cmul(x[k1], words_to_c(w[4 * k1], w[4 * k1 + 1], w[4 * k1 + 2], w[4 * k1 + 3]));
        env.warp_sync();
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = ex2[64 * hi + 8 * lo + (c ^ lo)];
    }
#endif
    if (SM::kAliased && !SM::kDouble) env.ex1_release();
    radix8<false>(x);
}

template <class Env, class SM>
PEBA1_HD void team_fft_inverse_tm(Env& env, double2 (&x)[8], SM& sm) {
    const int tid = env.tid;
    const int hi = tid >> 3, lo = tid & 7;
    const int par = SM::kDouble ? env.flip() : 0;
    double2* const ex1 = sm.e1(par);
    double2* const ex2 = sm.e2(par);
    uint32_t w[32];
    env.tmem_ld32(SM::kTw2, w);
    radix8<true>(x);
    if constexpr (SM::kXTm) {
        // the forward transform's inner exchange backwards: written by the plain thread (hi, lo), read by logical (hp, lp);
        // the twiddle of element (k1, c) is applied by the reader (it holds w64^(c k1) for c = lp): the same product
        const int hp = env.xhi(), lp = env.xlo(), sp = 2 * (hp & 3), ss = 2 * (hi & 3);
        env.warp_sync();  // the region's readers of the previous transform are done
#pragma unroll
        for (int c = 0; c < 8; ++c) ex2[64 * hi + 8 * lo + (c ^ lo ^ ss)] = x[c];
        env.warp_sync();
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) x[k1] = ex2[64 * hp + 8 * k1 + (lp ^ k1 ^ sp)];
        env.tmem_wait_ld();
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) x[k1] = cmul_conj(x[k1], words_to_c(w[4 * k1], w[4 * k1 + 1], w[4 * k1 + 2], w[4 * k1 + 3]));
        radix8<true>(x);
        uint32_t v[32];
#pragma unroll
        for (int b = 0; b < 8; ++b) c_to_words(x[b], v[xtm_reg(b, 0)], v[xtm_reg(b, 1)], v[xtm_reg(b, 2)], v[xtm_reg(b, 3)]);
        env.xtm_exchange_inv(v, w);  // also fetches tw1 into w (one tcgen05.wait::ld for both): w is not live across the trip
#pragma unroll
        for (int k0 = 0; k0 < 8; ++k0)
            x[k0] = cmul_conj(words_to_c(v[xtm_col(k0, 0)], v[xtm_col(k0, 1)], v[xtm_col(k0, 2)], v[xtm_col(k0, 3)]),
                              words_to_c(w[4 * k0], w[4 * k0 + 1], w[4 * k0 + 2], w[4 * k0 + 3]));
        radix8<true>(x);
#pragma unroll
        for (int a = 1; a < 8; ++a) x[a] = cmul_conj(x[a], w32(a));
        return;
    }
    env.tmem_wait_ld();
    if (SM::kShfl) {
#pragma unroll
        for (int c = 1; c < 8; ++c) x[c] = cmul_conj(x[c], words_to_c(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]));
        env.tmem_ld32(SM::kTw1, w);
        lane8_transpose(env, x, lo);
    } else {
        if (SM::kDouble) {}                        // a fresh buffer: nobody reads it any more (see TeamSmemTm2)
        else if (SM::kAliased) env.ex1_acquire();  // the previous transform's cross-team readers of this region are done
        else env.warp_sync();                      // previous readers of ex2 are done
        ex2[64 * hi + 8 * lo + lo] = x[0];
#pragma unroll
        for (int c = 1; c < 8; ++c)
            ex2[64 * hi + 8 * lo + (c ^ lo)] = cmul_conj(x[c], words_to_c(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]));
        env.tmem_ld32(SM::kTw1, w);
        env.warp_sync();
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) x[k1] = ex2[64 * hi + 8 * k1 + (lo ^ k1)];
    }
    radix8<true>(x);
    if (SM::kShfl) {}                     // kDouble layout, inner exchange not in memory: a fresh buffer
    else if (SM::kAliased) env.warp_sync();    // the region's readers above are done before it is rewritten
    else env.ex1_acquire();               // previous readers of ex1 are done
#pragma unroll
    for (int b = 0; b < 8; ++b) ex1[64 * hi + 8 * b + lo] = x[b];
    env.team_sync();
    env.tmem_wait_ld();
#pragma unroll
    for (int k0 = 0; k0 < 8; ++k0)
        x[k0] = cmul_conj(ex1[64 * k0 + tid], words_to_c(w[4 * k0], w[4 * k0 + 1], w[4 * k0 + 2], w[4 * k0 + 3]));
    if (!SM::kDouble) env.ex1_release();
    radix8<true>(x);
#pragma unroll
    for (int a = 1; a < 8; ++a) x[a] = cmul_conj(x[a], w32(a));
}

// Writes this thread's twiddles into tensor memory (called once per kernel by the warps with warp/4 == 0).
template <class SM = TeamSmemTm, class Env>
PEBA1_HD void tm_store_twiddles(Env& env, const TeamTwiddles& tw) {
    tm_store8(env, SM::kTw1, tw.tw1);
    tm_store8(env, SM::kTw2, tw.tw2);
    uint32_t z[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) z[j] = 0u;
    env.tmem_st16(SM::kZero, z);
    env.tmem_wait_st();
}

template <int L, int BGBIT, class Env, class SM>
PEBA1_HD void blind_rotate_team_tm(Env& env, SM& sm, const double2* __restrict__ bkfft, int n,
                                   const int32_t* __restrict__ pool, int64_t pool_stride, PbsJob job,
                                   int32_t* __restrict__ u_out) {
    constexpr int ROWS = 2 * L;
    constexpr uint32_t HALF_BG = 1u << (BGBIT - 1);
    constexpr uint32_t MASK_BG = (1u << BGBIT) - 1;
    uint32_t offset = 0;
#pragma unroll
    for (int p = 1; p <= L; ++p) offset += HALF_BG << (32 - p * BGBIT);
    const int tid = env.tid;
    const int sums = env.sums_col;  // kTmSums + 64 * (warp / 4)

    // ---- prologue + mod-switch (identical to blind_rotate_team)
    gate_prologue_modswitch(env, sm.bara, n, pool, pool_stride, job, tid, kTeam);
    env.team_sync();
    {
        const int rot = (2 * kN - (int)sm.bara[n]) & (2 * kN - 1);
        const int32_t mu = 1 << 29;
        // Thread tid is the only one that ever UPDATES words tid + 64 k of the accumulator (k < 8: real slots, k >= 8:
        // imaginary slots of its 8 transform points).  Its copy of those 16 words of the BODY polynomial also lives in
        // tensor memory (SM::kOwn): the un-rotated operand of (X^abar - 1) ACC and the read half of the update then cost
        // two tcgen05.ld per step instead of 32 shared-memory loads (-64 of ~3,075 L1 wavefronts per ciphertext-step).
        uint32_t own[16], zero[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int j = tid + kTeam * k;
            const int idx = (j - rot) & (2 * kN - 1);
            sm.acc[0][j] = 0;
            sm.acc[1][j] = (idx & kN) ? -mu : mu;
            own[k] = (uint32_t)((idx & kN) ? -mu : mu);
            zero[k] = 0u;
        }
        if (SM::kOwn >= 0) {
            env.tmem_st16(env.own_col, zero);
            env.tmem_st16(env.own_col + 16, own);
        }
        if (SM::kClear) {
#pragma unroll
            for (int q = 0; q < 4; ++q) env.tmem_st16(sums + 16 * q, zero);  // the MAC accumulators start every step at 0
        }
    }
    env.team_sync();

    for (int i = 0; i < n; ++i) {
        const int abar = sm.bara[i];
        if (abar == 0) continue;
#ifdef PEBA1_TV_IN_REGS
        uint32_t tv[16];
#endif
#pragma unroll 1
        for (int r = 0; r < ROWS; ++r) {
            const int poly = r >= L ? 1 : 0;
            const int p = r - poly * L;
#ifndef PEBA1_TV_IN_REGS  /* the default; -DPEBA1_TV_IN_REGS rebuilds the r1f kernel for comparison */
            // The rotated differences feed L digit levels.  Held in registers across the transforms they cost 16
            // registers and push the kernel into spilling (LDL/STL = 6 % of its L1 wavefronts); parked in tensor
            // memory between the levels they cost one store and L-1 loads per polynomial.
            uint32_t tv[16];
            if (p == 0) {
                if (SM::kOwn >= 0) {
                    uint32_t own[16];
                    env.tmem_ld16(env.own_col + 16 * poly, own);  // stored at the last update, several tcgen05.wait::st ago
                    env.tmem_wait_ld();
#pragma unroll
                    for (int a = 0; a < 8; ++a) {
                        tv[a] = (uint32_t)rot_only<Env>(sm.acc[poly], 64 * a + tid, abar) - own[a] + offset;
                        tv[8 + a] = (uint32_t)rot_only<Env>(sm.acc[poly], 64 * a + tid + 512, abar) - own[8 + a] + offset;
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < 8; ++a) {
                        tv[a] = (uint32_t)rot_diff<Env>(sm.acc[poly], 64 * a + tid, abar) + offset;
                        tv[8 + a] = (uint32_t)rot_diff<Env>(sm.acc[poly], 64 * a + tid + 512, abar) + offset;
                    }
                }
                if constexpr (SM::kXTm) {  // no tensor-memory columns left: the thread's own 64 bytes of shared memory
#pragma unroll
                    for (int q = 0; q < 4; ++q) sm.tv[q][tid] = make_uint4(tv[4 * q], tv[4 * q + 1], tv[4 * q + 2], tv[4 * q + 3]);
                } else {
                    env.tmem_st16(env.tv_col, tv);
                }
            } else {
                if constexpr (SM::kXTm) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 t4 = sm.tv[q][tid];
                        tv[4 * q] = t4.x; tv[4 * q + 1] = t4.y; tv[4 * q + 2] = t4.z; tv[4 * q + 3] = t4.w;
                    }
                } else {
                    env.tmem_wait_st();
                    env.tmem_ld16(env.tv_col, tv);
                    env.tmem_wait_ld();
                }
            }
#else
            if (p == 0) {
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    tv[a] = (uint32_t)rot_diff<Env>(sm.acc[poly], 64 * a + tid, abar) + offset;
                    tv[8 + a] = (uint32_t)rot_diff<Env>(sm.acc[poly], 64 * a + tid + 512, abar) + offset;
                }
            }
#endif
            const int shift = 32 - (p + 1) * BGBIT;
            double2 x[8];
#pragma unroll
            for (int a = 0; a < 8; ++a)
                x[a] = make_double2(digit2d((tv[a] >> shift) & MASK_BG, HALF_BG), digit2d((tv[8 + a] >> shift) & MASK_BG, HALF_BG));
            team_fft_forward_tm(env, x, sm);
            // pointwise MAC against BK_i row r.  Accumulators live in tensor memory; chunk q = (column c, 4 bins).
            // Software pipeline: the accumulators and key values of chunk q+1 are in flight while chunk q is
            // multiplied.  Row 0 reads its accumulators from the zero block (no branch, no register clearing).
            const double2* bk = bkfft + bkfft_index(ROWS, i, r, 0, 0, tid);
            if (SM::kClear || r > 0) env.tmem_wait_st();  // the previous row's (or the clearing) stores have landed
            uint32_t w0[16], w1[16];
            double2 b0[4], b1[4];
            const int zsel = (SM::kClear || r > 0) ? 1 : 0;  // layouts with a zero block read row 0's accumulators there
#define PEBA1_CHUNK_COL(q) (zsel ? sums + 32 * ((q) >> 1) + 16 * ((q) & 1) : SM::kZero)
#if defined(PEBA1_ABLATE) && PEBA1_ABLATE == 1  /* what-if build (scripts/ablate.sh): no key loads */
#define PEBA1_CHUNK_LDG(q, b)                                                                              \
    _Pragma("unroll") for (int e = 0; e < 4; ++e) b[e] = make_double2(1.0 + (q), 0.5 + e);
#else
#define PEBA1_CHUNK_LDG(q, b)                                                                              \
    _Pragma("unroll") for (int e = 0; e < 4; ++e)                                                          \
        b[e] = env.ldg2(bk + (size_t)(2 * (4 * ((q) & 1) + e) + ((q) >> 1)) * 64);
#endif
#define PEBA1_CHUNK_MAC(q, w, b)                                                                           \
    _Pragma("unroll") for (int e = 0; e < 4; ++e) {                                                        \
        double2 sacc = words_to_c(w[4 * e], w[4 * e + 1], w[4 * e + 2], w[4 * e + 3]);                      \
        cfma(sacc, x[4 * ((q) & 1) + e], b[e]);                                                            \
        c_to_words(sacc, w[4 * e], w[4 * e + 1], w[4 * e + 2], w[4 * e + 3]);                               \
    }                                                                                                      \
    env.tmem_st16(sums + 32 * ((q) >> 1) + 16 * ((q) & 1), w);
            env.tmem_ld16(PEBA1_CHUNK_COL(0), w0);
            PEBA1_CHUNK_LDG(0, b0)
            env.tmem_wait_ld();
            env.tmem_ld16(PEBA1_CHUNK_COL(1), w1);
            PEBA1_CHUNK_LDG(1, b1)
            PEBA1_CHUNK_MAC(0, w0, b0)
            env.tmem_wait_ld();
            env.tmem_ld16(PEBA1_CHUNK_COL(2), w0);
            PEBA1_CHUNK_LDG(2, b0)
            PEBA1_CHUNK_MAC(1, w1, b1)
            env.tmem_wait_ld();
            env.tmem_ld16(PEBA1_CHUNK_COL(3), w1);
            PEBA1_CHUNK_LDG(3, b1)
            PEBA1_CHUNK_MAC(2, w0, b0)
            env.tmem_wait_ld();
            PEBA1_CHUNK_MAC(3, w1, b1)
#undef PEBA1_CHUNK_COL
#undef PEBA1_CHUNK_LDG
#undef PEBA1_CHUNK_MAC
        }
        env.tmem_wait_st();
        // inverse transforms and accumulator update (round to nearest, wrap mod 2^32)
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            double2 s[8];
            tm_load8(env, sums + 32 * c, s);
            if (SM::kClear) {  // clear the column for the next step
                uint32_t zero[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) zero[k] = 0u;
                env.tmem_st16(sums + 32 * c, zero);
                env.tmem_st16(sums + 32 * c + 16, zero);
            }
            team_fft_inverse_tm(env, s, sm);
            if (SM::kOwn >= 0) {
                uint32_t own[16];
                env.tmem_ld16(env.own_col + 16 * c, own);
                env.tmem_wait_ld();
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    own[a] += d2u_rn(s[a].x);
                    own[8 + a] += d2u_rn(s[a].y);
                    sm.acc[c][64 * a + tid] = (int32_t)own[a];
                    sm.acc[c][64 * a + tid + 512] = (int32_t)own[8 + a];
                }
                env.tmem_st16(env.own_col + 16 * c, own);
            } else {
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    sm.acc[c][64 * a + tid] = (int32_t)((uint32_t)sm.acc[c][64 * a + tid] + d2u_rn(s[a].x));
                    sm.acc[c][64 * a + tid + 512] = (int32_t)((uint32_t)sm.acc[c][64 * a + tid + 512] + d2u_rn(s[a].y));
                }
            }
        }
        env.team_sync();
    }

    for (int j = tid; j < kN; j += kTeam)
        u_out[j] = (j == 0) ? sm.acc[0][0] : (int32_t)(0u - (uint32_t)sm.acc[0][kN - j]);
    if (tid == 0) u_out[kN] = sm.acc[1][0];
    env.team_sync();
}

}  // namespace peba1
