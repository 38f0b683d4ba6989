// blind_rotate_tm_variants.cuh — measured alternatives of blind_rotate_team_tm that LOST against the default and are not
// part of the product build (make VARIANTS=all compiles them; DESIGN.md §5 has the numbers): the register-accumulator
// version for 6 teams per SM (kernel_variant 9) and the shared-key version (kernel_variant 2).  Same arithmetic, same words.
#pragma once

#include "blind_rotate_tm.cuh"

namespace peba1 {

// Register-accumulator variant of blind_rotate_team_tm for 6 teams per SM (12 warps, up to 168 registers): the MAC
// accumulators stay in registers for the whole step (no tcgen05.ld/st per row and none of the register moves ptxas wraps
// around them), twiddles and the decomposition input stay in tensor memory.  Measured basis: 6 teams per SM deliver the
// same throughput as 8 with the tensor-memory MAC (one wave: 8.3 ms for 888 vs 10.9 ms for 1,184).
template <int L, int BGBIT, class Env, class SM>
PEBA1_HD void blind_rotate_team_tmr(Env& env, SM& sm, const double2* __restrict__ bkfft, int n,
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
        for (int j = tid; j < kN; j += kTeam) {
            const int idx = (j - rot) & (2 * kN - 1);
            sm.acc[0][j] = 0;
            sm.acc[1][j] = (idx & kN) ? -mu : mu;
        }
    }
    env.team_sync();

    for (int i = 0; i < n; ++i) {
        const int abar = sm.bara[i];
        if (abar == 0) continue;
        double2 sum0[8], sum1[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { sum0[q] = make_double2(0.0, 0.0); sum1[q] = make_double2(0.0, 0.0); }
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
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    tv[a] = (uint32_t)rot_diff<Env>(sm.acc[poly], 64 * a + tid, abar) + offset;
                    tv[8 + a] = (uint32_t)rot_diff<Env>(sm.acc[poly], 64 * a + tid + 512, abar) + offset;
                }
                env.tmem_st16(env.tv_col, tv);
            } else {
                env.tmem_wait_st();
                env.tmem_ld16(env.tv_col, tv);
                env.tmem_wait_ld();
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
            // pointwise MAC against BK_i row r, accumulators in registers (6 teams per SM leave 168 registers per thread)
            const double2* bk = bkfft + bkfft_index(ROWS, i, r, 0, 0, tid);
            // chunk q = (column q >> 1, bins 4 (q & 1) ..): the key values of chunk q+1 are requested before chunk q is
            // multiplied (the same software pipeline as the tensor-memory MAC)
            double2 b0[4], b1[4];
#define PEBA1_RLDG(q, b)                                                                                   \
    _Pragma("unroll") for (int e = 0; e < 4; ++e)                                                          \
        b[e] = env.ldg2(bk + (size_t)(2 * (4 * ((q) & 1) + e) + ((q) >> 1)) * 64);
#define PEBA1_RMAC(q, b)                                                                                   \
    _Pragma("unroll") for (int e = 0; e < 4; ++e) {                                                        \
        if ((q) >> 1) cfma(sum1[4 * ((q) & 1) + e], x[4 * ((q) & 1) + e], b[e]);                           \
        else cfma(sum0[4 * ((q) & 1) + e], x[4 * ((q) & 1) + e], b[e]);                                    \
    }
            PEBA1_RLDG(0, b0)
            PEBA1_RLDG(1, b1)
            PEBA1_RMAC(0, b0)
            PEBA1_RLDG(2, b0)
            PEBA1_RMAC(1, b1)
            PEBA1_RLDG(3, b1)
            PEBA1_RMAC(2, b0)
            PEBA1_RMAC(3, b1)
#undef PEBA1_RLDG
#undef PEBA1_RMAC
        }
        // inverse transforms and accumulator update (round to nearest, wrap mod 2^32)
        team_fft_inverse_tm(env, sum0, sm);
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            sm.acc[0][64 * a + tid] = (int32_t)((uint32_t)sm.acc[0][64 * a + tid] + d2u_rn(sum0[a].x));
            sm.acc[0][64 * a + tid + 512] = (int32_t)((uint32_t)sm.acc[0][64 * a + tid + 512] + d2u_rn(sum0[a].y));
        }
        team_fft_inverse_tm(env, sum1, sm);
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            sm.acc[1][64 * a + tid] = (int32_t)((uint32_t)sm.acc[1][64 * a + tid] + d2u_rn(sum1[a].x));
            sm.acc[1][64 * a + tid + 512] = (int32_t)((uint32_t)sm.acc[1][64 * a + tid + 512] + d2u_rn(sum1[a].y));
        }
        env.team_sync();
    }

    for (int j = tid; j < kN; j += kTeam)
        u_out[j] = (j == 0) ? sm.acc[0][0] : (int32_t)(0u - (uint32_t)sm.acc[0][kN - j]);
    if (tid == 0) u_out[kN] = sm.acc[1][0];
    env.team_sync();
}


// ---------------------------------------------------------------------------------------------------------------
// Shared-key variant.  In blind_rotate_team_tm every team streams the whole key row (16 KB) from L2 through the L1
// data pipe for itself (ncu r1f: L1 hit rate 15 %, key loads = 23 points of the 78 % L1 pipe utilisation, 6.5 TB/s
// of L2->SM traffic).  The four teams whose warps share a tensor-memory lane quarter (teams t, t+2, t+4, t+6) hold,
// lane for lane, the same thread index, hence need the very same key values.  Here they form a GROUP: for each row
// each team loads ONE quarter of the row (4 values per thread, issued before its transform so the L2 latency is
// hidden), publishes it in tensor memory, and after a group barrier all four read the whole row from there.  Key
// traffic through L1/L2 drops 4x.  Price: the group advances roughly in step (split-phase mbarrier per row buffer,
// two buffers), and a team whose step is void (abar == 0) or that has no job still walks the rows as a producer.
template <int L, int BGBIT, class Env>
PEBA1_HD void blind_rotate_team_tmsh(Env& env, TeamSmemTm& sm, const double2* __restrict__ bkfft, int n,
                                     const int32_t* __restrict__ pool, int64_t pool_stride, PbsJob job, bool live,
                                     int32_t* __restrict__ u_out) {
    constexpr int ROWS = 2 * L;
    constexpr uint32_t HALF_BG = 1u << (BGBIT - 1);
    constexpr uint32_t MASK_BG = (1u << BGBIT) - 1;
    uint32_t offset = 0;
#pragma unroll
    for (int p = 1; p <= L; ++p) offset += HALF_BG << (32 - p * BGBIT);
    const int tid = env.tid;
    const int sums = env.sums_col;
    const int myq = env.chunk_id;  // which quarter of every key row this team fetches: q = (column c, half)

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

    // Row s = i*ROWS + r of the key is published (by all four teams, a quarter each) into buffer s & 1 and announced
    // on the split-phase barrier of that buffer: a team ARRIVES right after publishing its quarter of row s+1 (end of
    // its row s) and WAITS only just before its MAC of row s+1, i.e. one transform later, so the group tolerates about
    // one transform of skew instead of meeting at every row.  Overwriting buffer (s+1)&1 is safe: a team gets past its
    // wait for row s only after every team arrived for row s, which each does after its MAC of row s-1 (the last
    // reader of that buffer).
    const int total_rows = n * ROWS;
    {
        const double2* bk = bkfft + bkfft_index(ROWS, 0, 0, 0, 0, tid);
        uint32_t w[16];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const double2 v = env.ldg2(bk + (size_t)(2 * (4 * (myq & 1) + e) + (myq >> 1)) * 64);
            c_to_words(v, w[4 * e], w[4 * e + 1], w[4 * e + 2], w[4 * e + 3]);
        }
        env.tmem_st16(kTmBk + 16 * myq, w);
        env.tmem_wait_st();
        env.group_arrive(0);
    }
    int srow = 0;
    for (int i = 0; i < n; ++i) {
        const int abar = sm.bara[i];
        const bool work = abar != 0;  // team-uniform; a void step still produces its share of the key rows
        uint32_t tv[16];
#pragma unroll 1
        for (int r = 0; r < ROWS; ++r, ++srow) {
            // 1. my quarter of the NEXT key row: in flight during this row's transform
            const bool more = srow + 1 < total_rows;
            const int ni = r + 1 < ROWS ? i : i + 1, nr = r + 1 < ROWS ? r + 1 : 0;
            double2 mine[4];
            if (more) {
                const double2* bkn = bkfft + bkfft_index(ROWS, ni, nr, 0, 0, tid);
#pragma unroll
                for (int e = 0; e < 4; ++e) mine[e] = env.ldg2(bkn + (size_t)(2 * (4 * (myq & 1) + e) + (myq >> 1)) * 64);
            }
            // 2. decompose + forward transform
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
            // 3. row s is complete in tensor memory once every team has arrived for it
            const int buf = kTmBk + 64 * (srow & 1);
            env.group_wait(srow & 1, (srow >> 1) & 1);
            // 4. pointwise MAC, accumulators and key both from tensor memory, chunk q+1 in flight during chunk q
            if (work) {
                if (r > 0) env.tmem_wait_st();  // the previous row's accumulator stores have landed
                uint32_t w0[16], w1[16], k0[16], k1[16];
                const int zsel = r > 0 ? 1 : 0;
#define PEBA1_CHUNK_COL(q) (zsel ? sums + 16 * (q) : kTmZero)
#define PEBA1_CHUNK_MAC(q, w, k)                                                                           \
    _Pragma("unroll") for (int e = 0; e < 4; ++e) {                                                        \
        double2 sacc = words_to_c(w[4 * e], w[4 * e + 1], w[4 * e + 2], w[4 * e + 3]);                      \
        cfma(sacc, x[4 * ((q) & 1) + e], words_to_c(k[4 * e], k[4 * e + 1], k[4 * e + 2], k[4 * e + 3]));   \
        c_to_words(sacc, w[4 * e], w[4 * e + 1], w[4 * e + 2], w[4 * e + 3]);                               \
    }                                                                                                      \
    env.tmem_st16(sums + 16 * (q), w);
                env.tmem_ld16(PEBA1_CHUNK_COL(0), w0);
                env.tmem_ld16(buf, k0);
                env.tmem_wait_ld();
                env.tmem_ld16(PEBA1_CHUNK_COL(1), w1);
                env.tmem_ld16(buf + 16, k1);
                PEBA1_CHUNK_MAC(0, w0, k0)
                env.tmem_wait_ld();
                env.tmem_ld16(PEBA1_CHUNK_COL(2), w0);
                env.tmem_ld16(buf + 32, k0);
                PEBA1_CHUNK_MAC(1, w1, k1)
                env.tmem_wait_ld();
                env.tmem_ld16(PEBA1_CHUNK_COL(3), w1);
                env.tmem_ld16(buf + 48, k1);
                PEBA1_CHUNK_MAC(2, w0, k0)
                env.tmem_wait_ld();
                PEBA1_CHUNK_MAC(3, w1, k1)
#undef PEBA1_CHUNK_COL
#undef PEBA1_CHUNK_MAC
            }
            // 5. publish my quarter of row s+1 and announce it
            if (more) {
                uint32_t w[16];
#pragma unroll
                for (int e = 0; e < 4; ++e) c_to_words(mine[e], w[4 * e], w[4 * e + 1], w[4 * e + 2], w[4 * e + 3]);
                env.tmem_st16(kTmBk + 64 * ((srow + 1) & 1) + 16 * myq, w);
                env.tmem_wait_st();
                env.group_arrive((srow + 1) & 1);
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
