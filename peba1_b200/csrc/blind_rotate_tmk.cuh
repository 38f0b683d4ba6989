// blind_rotate_tmk.cuh — blind rotation with the bootstrapping-key rows delivered through TENSOR MEMORY:
//   L2 --(cp.async.bulk, TMA 1-D, one 8 KB half-row per copy)--> shared memory staging (3 slots)
//      --(tcgen05.cp.64x128b.warpx2::02_13, 8 per half-row)--> tensor memory ring (4 half-row slots, all 128 lanes)
//      --(tcgen05.ld.32x32b.x16 by every warp)--> registers of the 8 teams of the CTA.
//
// Why: in blind_rotate_team_tm every team pulls every key row from L2 through the L1 data pipe for itself: 16
// LDG.128 per thread and row = 770 of the 3,170 L1 wavefronts per ciphertext-step (ncu r1k: L1 pipe 78 % busy, the
// binding unit), each with a uniform-register descriptor reload, and the L2 latency shows up as long-scoreboard stalls
// at the first DFMA of a chunk (ptxas sinks the loads next to their use).  The what-if build without key loads runs
// 17 % faster.  Here ONE thread of the CTA moves the row once for all 8 teams with two asynchronous engines that do not
// touch the LSU/L1 data pipe at all, and the teams read it from tensor memory, the datapath that already holds their
// accumulators.  The global key layout [i][r][k2][c][tid] of double2 is, per (k2, c), exactly the 64-row x 16-byte
// matrix tcgen05.cp wants: thread-index tid = row; the 02_13 flavour sends rows 0-31 to lane quarters 0 and 2 and rows
// 32-63 to quarters 1 and 3, which is where the warps with tid 0-31 / 32-63 live (quarter = warp % 4, team = 2 warps).
//
// Protocol (half-row h = 2*(i*ROWS + r) + half; half = k2 / 4).  Four mbarrier families:
//   tma_full[h%3]   TMA of half-row h landed in staging slot h%3           (producer -> producer)
//   smem_free[h%3]  the tcgen05.cp's of half-row h have read the slot      (tcgen05.commit -> producer)
//   tmem_full[h&3]  half-row h is readable in tensor-memory slot h&3       (tcgen05.commit -> all warps)
//   tmem_free[h&3]  all 16 warps have finished their tcgen05.ld of half-row h (one arrival per warp -> producer)
// The producer is nobody in particular.  At two points of every row (before the forward transform, before the MAC)
// and while waiting for a half-row, thread 0 of every team TRIES to advance the pipeline: under a lock it tests, without
// blocking, whether the next tcgen05.cp (TMA landed, tensor-memory slot released) and the next TMA (staging slot read
// out) can be issued, and issues them.  The pipeline therefore runs as far ahead as its rings allow (4 half-rows of
// tensor memory, 3 of staging) and no team is slower than the others.  Measured alternatives: one fixed producer
// thread makes its team the laggard that the other seven wait for (80 k PBS/s); a role rotating over the teams in a
// fixed order makes every team wait for its predecessor (92 k).  A team without work (void step, or no job) still walks
// the rows as a consumer.
//
// Arithmetic per ciphertext: that of blind_rotate_team_tm with the four (column, half) chunks of a row visited in
// the order (c0,h0) (c1,h0) (c0,h1) (c1,h1); the chunks touch disjoint accumulators, so results are bit-identical.
#pragma once

#include "blind_rotate_tm.cuh"

namespace peba1 {

constexpr int kTmKey = 336;          // 4 half-row slots x 32 columns: [slot][c (2)][k2 & 3 (4)][4 words]
constexpr int kKeyHalfBytes = 8192;  // one half-row: k2 in [4*half, 4*half + 4), both columns, 64 thread indices

template <int L, int BGBIT, class Env>
PEBA1_HD void blind_rotate_team_tmk(Env& env, TeamSmemTm& sm, int n, const int32_t* __restrict__ pool,
                                    int64_t pool_stride, PbsJob job, bool live, int32_t* __restrict__ u_out) {
    constexpr int ROWS = 2 * L;
    constexpr uint32_t HALF_BG = 1u << (BGBIT - 1);
    constexpr uint32_t MASK_BG = (1u << BGBIT) - 1;
    uint32_t offset = 0;
#pragma unroll
    for (int p = 1; p <= L; ++p) offset += HALF_BG << (32 - p * BGBIT);
    const int tid = env.tid;
    const int sums = env.sums_col;

    gate_prologue_modswitch(env, sm.bara, n, pool, pool_stride, job, tid, kTeam);
    env.team_sync();
    {
        const int rot = (2 * kN - (int)sm.bara[n]) & (2 * kN - 1);
        const int32_t mu = 1 << 29;
        for (int j = tid; j < kN; j += kTeam) {
            const int idx = (j - rot) & (2 * kN - 1);
            sm.acc[0][j] = 0;
            sm.acc[1][j] = (idx & kN) ? -mu : mu;
        }
    }
    env.team_sync();

    env.key_pump(-1);  // start the pipeline (the first TMA)

    int srow = 0;
    for (int i = 0; i < n; ++i) {
        const int abar = sm.bara[i];
        const bool work = abar != 0;  // team-uniform; a void step still consumes (and releases) its key rows
        uint32_t tv[16];
#pragma unroll 1
        for (int r = 0; r < ROWS; ++r, ++srow) {
            env.key_pump(2 * srow);
            double2 x[8];
            if (work) {
                const int poly = r >= L ? 1 : 0;
                const int p = r - poly * L;
                if (p == 0) {
#pragma unroll
                    for (int a = 0; a < 8; ++a) {
                        tv[a] = (uint32_t)rot_diff<Env>(sm.acc[poly], 64 * a + tid, abar) + offset;
                        tv[8 + a] = (uint32_t)rot_diff<Env>(sm.acc[poly], 64 * a + tid + 512, abar) + offset;
                    }
                }
                const int shift = 32 - (p + 1) * BGBIT;
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const int32_t dr = (int32_t)((tv[a] >> shift) & MASK_BG) - (int32_t)HALF_BG;
                    const int32_t di = (int32_t)((tv[8 + a] >> shift) & MASK_BG) - (int32_t)HALF_BG;
                    x[a] = make_double2(i2d(dr), i2d(di));
                }
                team_fft_forward_tm(env, x, sm);
            }
            env.key_pump(2 * srow + 1);
            const int h0 = 2 * srow;
            if (work) {
                if (r > 0) env.tmem_wait_st();  // the previous row's accumulator stores have landed
                uint32_t w0[16], w1[16], k0[16], k1[16];
                const int zsel = r > 0 ? 1 : 0;
                // accumulator block of (column c, half hf): bins 4*hf .. 4*hf+3 of column c
#define PEBA1_KCOL(c, hf) (zsel ? sums + 32 * (c) + 16 * (hf) : kTmZero)
#define PEBA1_KMAC(c, hf, w, k)                                                                            \
    _Pragma("unroll") for (int e = 0; e < 4; ++e) {                                                        \
        double2 sacc = words_to_c(w[4 * e], w[4 * e + 1], w[4 * e + 2], w[4 * e + 3]);                      \
        cfma(sacc, x[4 * (hf) + e], words_to_c(k[4 * e], k[4 * e + 1], k[4 * e + 2], k[4 * e + 3]));        \
        c_to_words(sacc, w[4 * e], w[4 * e + 1], w[4 * e + 2], w[4 * e + 3]);                               \
    }                                                                                                      \
    env.tmem_st16(sums + 32 * (c) + 16 * (hf), w);
                env.key_wait(h0);
                env.tmem_ld16(PEBA1_KCOL(0, 0), w0);
                env.key_ld16(h0, 0, k0);
                env.tmem_wait_ld();
                env.tmem_ld16(PEBA1_KCOL(1, 0), w1);
                env.key_ld16(h0, 1, k1);
                PEBA1_KMAC(0, 0, w0, k0)
                env.tmem_wait_ld();
                env.key_release(h0);
                env.key_wait(h0 + 1);
                env.tmem_ld16(PEBA1_KCOL(0, 1), w0);
                env.key_ld16(h0 + 1, 0, k0);
                PEBA1_KMAC(1, 0, w1, k1)
                env.tmem_wait_ld();
                env.tmem_ld16(PEBA1_KCOL(1, 1), w1);
                env.key_ld16(h0 + 1, 1, k1);
                PEBA1_KMAC(0, 1, w0, k0)
                env.tmem_wait_ld();
                env.key_release(h0 + 1);
                PEBA1_KMAC(1, 1, w1, k1)
#undef PEBA1_KCOL
#undef PEBA1_KMAC
            } else {
                env.key_wait(h0);
                env.key_release(h0);
                env.key_wait(h0 + 1);
                env.key_release(h0 + 1);
            }
        }
        if (work) {
            env.tmem_wait_st();
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                double2 s[8];
                tm_load8(env, sums + 32 * c, s);
                team_fft_inverse_tm(env, s, sm);
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    sm.acc[c][64 * a + tid] = (int32_t)((uint32_t)sm.acc[c][64 * a + tid] + d2u_rn(s[a].x));
                    sm.acc[c][64 * a + tid + 512] = (int32_t)((uint32_t)sm.acc[c][64 * a + tid + 512] + d2u_rn(s[a].y));
                }
            }
            env.team_sync();
        }
    }

    if (live) {
        for (int j = tid; j < kN; j += kTeam)
            u_out[j] = (j == 0) ? sm.acc[0][0] : (int32_t)(0u - (uint32_t)sm.acc[0][kN - j]);
        if (tid == 0) u_out[kN] = sm.acc[1][0];
    }
    env.team_sync();
}

}  // namespace peba1
