// blind_rotate_pair.cuh — blind rotation where one TEAM (64 threads) carries TWO bootstraps at once.
//
// Why: ncu (profiles/r1f, r1k) puts blind_rotate_team_tm on the L1/shared-memory data pipe (78 % busy) with the FP64
// pipe at 55 %, and 8 warps per SM run as fast as 16.  So the kernel does not need 16 warps of 128 registers; it needs
// fewer wavefronts through L1.  With 8 warps of up to 255 registers each thread holds the 8 points of two different
// ciphertexts and runs both transforms in lock step:
//   * every key value fetched from L2 is used for two ciphertexts (key wavefronts per ciphertext-step 770 -> 385);
//   * one tensor-memory twiddle load, one named barrier and one mbarrier round per PAIR of transforms;
//   * two independent butterfly networks per thread hide the FP64 latency that the second set of warps used to hide.
// Arithmetic per ciphertext is instruction for instruction that of blind_rotate_team_tm (bit-exact vs the
// exact-integer oracle; csrc/emu_team.cu runs this code on CPU threads).
//
// TMEM layout (512 columns, 8 warps): lane quarter = warp % 4 (hardware); columns
//   [128*(warp/4) + 64*ct, +64)   accumulators of ciphertext ct (0/1) of this thread: [c (2)][k2 (8)][4 words]
//   [256, 320) twiddles, [320, 336) zero block   (as in blind_rotate_tm.cuh)
#pragma once

#include "blind_rotate_tm.cuh"

namespace peba1 {

struct __align__(16) TeamSmemPair {
    double2  ex1[2][kM];       // [ciphertext]: cross-team exchange
    double2  ex2[2][kM];       // [ciphertext]: exchange inside groups of 8 lanes
    int32_t  acc[2][2][kN];    // [ciphertext][mask, body]
    uint16_t bara[2][640];
    unsigned long long ex1_free;  // mbarrier, 2 arrivals (one per warp), guards both ex1 buffers
    unsigned long long pad_;
};

#define PEBA1_TW(w, j) words_to_c(w[4 * (j)], w[4 * (j) + 1], w[4 * (j) + 2], w[4 * (j) + 3])

template <class Env>
PEBA1_HD void pair_fft_forward_tm(Env& env, double2 (&x)[2][8], TeamSmemPair& sm) {
    const int tid = env.tid;
    uint32_t w[32];
    env.tmem_ld32(kTmTw1, w);
#pragma unroll
    for (int ct = 0; ct < 2; ++ct) {
#pragma unroll
        for (int a = 1; a < 8; ++a) x[ct][a] = cmul(x[ct][a], w32(a));
        radix8<false>(x[ct]);
    }
    env.tmem_wait_ld();
    env.ex1_acquire();
#pragma unroll
    for (int k0 = 0; k0 < 8; ++k0) {
        const double2 t = PEBA1_TW(w, k0);
        sm.ex1[0][64 * k0 + tid] = cmul(x[0][k0], t);
        sm.ex1[1][64 * k0 + tid] = cmul(x[1][k0], t);
    }
    env.tmem_ld32(kTmTw2, w);
    env.team_sync();
    const int hi = tid >> 3, lo = tid & 7;
#pragma unroll
    for (int ct = 0; ct < 2; ++ct) {
#pragma unroll
        for (int b = 0; b < 8; ++b) x[ct][b] = sm.ex1[ct][64 * hi + 8 * b + lo];
    }
    env.ex1_release();
    radix8<false>(x[0]);
    radix8<false>(x[1]);
    env.tmem_wait_ld();
    sm.ex2[0][64 * hi + lo] = x[0][0];
    sm.ex2[1][64 * hi + lo] = x[1][0];
#pragma unroll
    for (int k1 = 1; k1 < 8; ++k1) {
        const double2 t = PEBA1_TW(w, k1);
        sm.ex2[0][64 * hi + 8 * k1 + (lo ^ k1)] = cmul(x[0][k1], t);
        sm.ex2[1][64 * hi + 8 * k1 + (lo ^ k1)] = cmul(x[1][k1], t);
    }
    env.warp_sync();
#pragma unroll
    for (int ct = 0; ct < 2; ++ct) {
#pragma unroll
        for (int c = 0; c < 8; ++c) x[ct][c] = sm.ex2[ct][64 * hi + 8 * lo + (c ^ lo)];
    }
    radix8<false>(x[0]);
    radix8<false>(x[1]);
}

template <class Env>
PEBA1_HD void pair_fft_inverse_tm(Env& env, double2 (&x)[2][8], TeamSmemPair& sm) {
    const int tid = env.tid;
    const int hi = tid >> 3, lo = tid & 7;
    uint32_t w[32];
    env.tmem_ld32(kTmTw2, w);
    radix8<true>(x[0]);
    radix8<true>(x[1]);
    env.tmem_wait_ld();
    env.warp_sync();  // previous readers of ex2 are done
    sm.ex2[0][64 * hi + 8 * lo + lo] = x[0][0];
    sm.ex2[1][64 * hi + 8 * lo + lo] = x[1][0];
#pragma unroll
    for (int c = 1; c < 8; ++c) {
        const double2 t = PEBA1_TW(w, c);
        sm.ex2[0][64 * hi + 8 * lo + (c ^ lo)] = cmul_conj(x[0][c], t);
        sm.ex2[1][64 * hi + 8 * lo + (c ^ lo)] = cmul_conj(x[1][c], t);
    }
    env.tmem_ld32(kTmTw1, w);
    env.warp_sync();
#pragma unroll
    for (int ct = 0; ct < 2; ++ct) {
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) x[ct][k1] = sm.ex2[ct][64 * hi + 8 * k1 + (lo ^ k1)];
    }
    radix8<true>(x[0]);
    radix8<true>(x[1]);
    env.ex1_acquire();  // previous readers of ex1 are done
#pragma unroll
    for (int ct = 0; ct < 2; ++ct) {
#pragma unroll
        for (int b = 0; b < 8; ++b) sm.ex1[ct][64 * hi + 8 * b + lo] = x[ct][b];
    }
    env.team_sync();
    env.tmem_wait_ld();
#pragma unroll
    for (int k0 = 0; k0 < 8; ++k0) {
        const double2 t = PEBA1_TW(w, k0);
        x[0][k0] = cmul_conj(sm.ex1[0][64 * k0 + tid], t);
        x[1][k0] = cmul_conj(sm.ex1[1][64 * k0 + tid], t);
    }
    env.ex1_release();
#pragma unroll
    for (int ct = 0; ct < 2; ++ct) {
        radix8<true>(x[ct]);
#pragma unroll
        for (int a = 1; a < 8; ++a) x[ct][a] = cmul_conj(x[ct][a], w32(a));
    }
}

// job0 / job1: the two bootstraps of this team; u1 == nullptr: the second one is a shadow (odd batch) whose result is
// dropped.  A void step of one ciphertext (abar == 0) is NOT special-cased: (X^0 - 1) * ACC = 0 decomposes into all-zero
// digits, every transform value is +-0 and the accumulator is left unchanged, exactly as if the step were skipped.
template <int L, int BGBIT, class Env>
PEBA1_HD void blind_rotate_pair_tm(Env& env, TeamSmemPair& sm, const double2* __restrict__ bkfft, int n,
                                   const int32_t* __restrict__ pool, int64_t pool_stride, PbsJob job0, PbsJob job1,
                                   int32_t* __restrict__ u0, int32_t* __restrict__ u1) {
    constexpr int ROWS = 2 * L;
    constexpr uint32_t HALF_BG = 1u << (BGBIT - 1);
    constexpr uint32_t MASK_BG = (1u << BGBIT) - 1;
    uint32_t offset = 0;
#pragma unroll
    for (int p = 1; p <= L; ++p) offset += HALF_BG << (32 - p * BGBIT);
    const int tid = env.tid;
    const int sums = env.sums_col;  // 128 * (warp / 4); ciphertext ct at + 64 * ct

    // ---- prologue + mod-switch
#pragma unroll 1
    for (int ct = 0; ct < 2; ++ct) {
        const PbsJob job = ct ? job1 : job0;
        gate_prologue_modswitch(env, sm.bara[ct], n, pool, pool_stride, job, tid, kTeam);
    }
    env.team_sync();
#pragma unroll 1
    for (int ct = 0; ct < 2; ++ct) {
        const int rot = (2 * kN - (int)sm.bara[ct][n]) & (2 * kN - 1);
        const int32_t mu = 1 << 29;
        for (int j = tid; j < kN; j += kTeam) {
            const int idx = (j - rot) & (2 * kN - 1);
            sm.acc[ct][0][j] = 0;
            sm.acc[ct][1][j] = (idx & kN) ? -mu : mu;
        }
    }
    env.team_sync();

    for (int i = 0; i < n; ++i) {
        const int abar0 = sm.bara[0][i], abar1 = sm.bara[1][i];
        if ((abar0 | abar1) == 0) continue;
        uint32_t tv[2][16];
#pragma unroll 1
        for (int r = 0; r < ROWS; ++r) {
            const int poly = r >= L ? 1 : 0;
            const int p = r - poly * L;
            if (p == 0) {
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    tv[0][a] = (uint32_t)rot_diff<Env>(sm.acc[0][poly], 64 * a + tid, abar0) + offset;
                    tv[0][8 + a] = (uint32_t)rot_diff<Env>(sm.acc[0][poly], 64 * a + tid + 512, abar0) + offset;
                    tv[1][a] = (uint32_t)rot_diff<Env>(sm.acc[1][poly], 64 * a + tid, abar1) + offset;
                    tv[1][8 + a] = (uint32_t)rot_diff<Env>(sm.acc[1][poly], 64 * a + tid + 512, abar1) + offset;
                }
            }
            const int shift = 32 - (p + 1) * BGBIT;
            double2 x[2][8];
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) {
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const int32_t dr = (int32_t)((tv[ct][a] >> shift) & MASK_BG) - (int32_t)HALF_BG;
                    const int32_t di = (int32_t)((tv[ct][8 + a] >> shift) & MASK_BG) - (int32_t)HALF_BG;
                    x[ct][a] = make_double2(i2d(dr), i2d(di));
                }
            }
            pair_fft_forward_tm(env, x, sm);
            // pointwise MAC against BK_i row r: one key fetch serves both ciphertexts.  Chunk q = (column c = q>>1,
            // bins 4*(q&1)..+3); accumulators and key of chunk q+1 are in flight while chunk q is multiplied.
            const double2* bk = bkfft + bkfft_index(ROWS, i, r, 0, 0, tid);
            if (r > 0) env.tmem_wait_st();
            uint32_t wa0[16], wa1[16], wb0[16], wb1[16];
            double2 b0[4], b1[4];
            const int zsel = r > 0 ? 1 : 0;
#define PEBA1_PCOL(q, ct) (zsel ? sums + 64 * (ct) + 32 * ((q) >> 1) + 16 * ((q) & 1) : kTmZero)
#define PEBA1_PLDG(q, b)                                                                                   \
    _Pragma("unroll") for (int e = 0; e < 4; ++e)                                                          \
        b[e] = env.ldg2(bk + (size_t)(2 * (4 * ((q) & 1) + e) + ((q) >> 1)) * 64);
#define PEBA1_PMAC(q, w, b, ct)                                                                            \
    _Pragma("unroll") for (int e = 0; e < 4; ++e) {                                                        \
        double2 sacc = words_to_c(w[4 * e], w[4 * e + 1], w[4 * e + 2], w[4 * e + 3]);                      \
        cfma(sacc, x[ct][4 * ((q) & 1) + e], b[e]);                                                        \
        c_to_words(sacc, w[4 * e], w[4 * e + 1], w[4 * e + 2], w[4 * e + 3]);                               \
    }                                                                                                      \
    env.tmem_st16(sums + 64 * (ct) + 32 * ((q) >> 1) + 16 * ((q) & 1), w);
            env.tmem_ld16(PEBA1_PCOL(0, 0), wa0);
            env.tmem_ld16(PEBA1_PCOL(0, 1), wb0);
            PEBA1_PLDG(0, b0)
            env.tmem_wait_ld();
            env.tmem_ld16(PEBA1_PCOL(1, 0), wa1);
            env.tmem_ld16(PEBA1_PCOL(1, 1), wb1);
            PEBA1_PLDG(1, b1)
            PEBA1_PMAC(0, wa0, b0, 0)
            PEBA1_PMAC(0, wb0, b0, 1)
            env.tmem_wait_ld();
            env.tmem_ld16(PEBA1_PCOL(2, 0), wa0);
            env.tmem_ld16(PEBA1_PCOL(2, 1), wb0);
            PEBA1_PLDG(2, b0)
            PEBA1_PMAC(1, wa1, b1, 0)
            PEBA1_PMAC(1, wb1, b1, 1)
            env.tmem_wait_ld();
            env.tmem_ld16(PEBA1_PCOL(3, 0), wa1);
            env.tmem_ld16(PEBA1_PCOL(3, 1), wb1);
            PEBA1_PLDG(3, b1)
            PEBA1_PMAC(2, wa0, b0, 0)
            PEBA1_PMAC(2, wb0, b0, 1)
            env.tmem_wait_ld();
            PEBA1_PMAC(3, wa1, b1, 0)
            PEBA1_PMAC(3, wb1, b1, 1)
#undef PEBA1_PCOL
#undef PEBA1_PLDG
#undef PEBA1_PMAC
        }
        env.tmem_wait_st();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            double2 s[2][8];
            {
                uint32_t w0[32], w1[32];
                env.tmem_ld32(sums + 32 * c, w0);
                env.tmem_ld32(sums + 64 + 32 * c, w1);
                env.tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    s[0][j] = words_to_c(w0[4 * j], w0[4 * j + 1], w0[4 * j + 2], w0[4 * j + 3]);
                    s[1][j] = words_to_c(w1[4 * j], w1[4 * j + 1], w1[4 * j + 2], w1[4 * j + 3]);
                }
            }
            pair_fft_inverse_tm(env, s, sm);
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) {
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    sm.acc[ct][c][64 * a + tid] = (int32_t)((uint32_t)sm.acc[ct][c][64 * a + tid] + d2u_rn(s[ct][a].x));
                    sm.acc[ct][c][64 * a + tid + 512] =
                        (int32_t)((uint32_t)sm.acc[ct][c][64 * a + tid + 512] + d2u_rn(s[ct][a].y));
                }
            }
        }
        env.team_sync();
    }

#pragma unroll 1
    for (int ct = 0; ct < 2; ++ct) {
        int32_t* u = ct ? u1 : u0;
        if (u == nullptr) continue;
        for (int j = tid; j < kN; j += kTeam)
            u[j] = (j == 0) ? sm.acc[ct][0][0] : (int32_t)(0u - (uint32_t)sm.acc[ct][0][kN - j]);
        if (tid == 0) u[kN] = sm.acc[ct][1][0];
    }
    env.team_sync();
}

#undef PEBA1_TW

}  // namespace peba1
