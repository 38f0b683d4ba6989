// blind_rotate.cuh — one TEAM (64 threads) bootstraps one LWE sample:
//   K0  gate prologue + mod-switch      (SURVEY.md §8c steps 1-2;  TFHE boots* linear part + modSwitchFromTorus32)
//   K1  blind rotation, n CMux steps    (steps 3-4; tfhe_blindRotate_FFT / tfhe_MuxRotate_FFT /
//                                        tGswFFTExternMulToTLwe — decomposition, 2l forward transforms,
//                                        pointwise MAC against BK_i, 2 inverse transforms, all fused)
//       sample extraction               (step 5; tLweExtractLweSample)
// The accumulator lives in shared memory as int32 for the whole bootstrap; every transform runs in registers
// (negacyclic_fft.cuh); the transformed key is read straight from L2 in a team-coalesced layout.
//
// __host__ __device__: emulated with CPU threads by csrc/emu_team.cu for the CPU-side tests.
#pragma once

#include "negacyclic_fft.cuh"

namespace peba1 {

// One bootstrap job: t = (0, c0) + s0 * pool[in0] + s1 * pool[in1] + s2 * pool[in2]   (in < 0: operand absent).
// The third operand serves the 3-input gates (XOR3, MAJ3: a full adder in two bootstraps).
struct __align__(8) PbsJob {
    int32_t in0, in1;
    int32_t c0;
    int16_t s0, s1;
    int32_t in2 = -1;
    int16_t s2 = 0, pad = 0;
};
static_assert(sizeof(PbsJob) == 24, "PbsJob layout");

template <class Env>
PEBA1_HD int32_t rot_diff(const int32_t* acc, int j, int abar) {
    // ((X^abar - 1) * acc)[j] with negacyclic wrap; 0 < abar < 2N
    const int idx = (j - abar) & (2 * kN - 1);
    int32_t v = acc[idx & (kN - 1)];
    v = (idx & kN) ? (int32_t)(0u - (uint32_t)v) : v;
    return (int32_t)((uint32_t)v - (uint32_t)acc[j]);
}

template <class Env>
PEBA1_HD int32_t rot_only(const int32_t* acc, int j, int abar) {
    // (X^abar * acc)[j] with negacyclic wrap; 0 < abar < 2N
    const int idx = (j - abar) & (2 * kN - 1);
    const int32_t v = acc[idx & (kN - 1)];
    return (idx & kN) ? (int32_t)(0u - (uint32_t)v) : v;
}

// K0, shared by every blind-rotation variant: gate prologue t = (0, c0) + s0 pool[in0] + s1 pool[in1] (+ s2 pool[in2]) and
// mod-switch to Z_2N, bar = ((uint32)t + 2^20) >> 21, written to `bara[0..n]` by `nthreads`
// threads (index `t0`).  The caller synchronises afterwards.
template <class Env>
PEBA1_HD void gate_prologue_modswitch(Env& env, uint16_t* bara, int n, const int32_t* __restrict__ pool, int64_t pool_stride,
                                      const PbsJob& job, int t0, int nthreads) {
    for (int w = t0; w <= n; w += nthreads) {
        uint32_t t = (w == n) ? (uint32_t)job.c0 : 0u;
        if (job.in0 >= 0) t += (uint32_t)(int32_t)job.s0 * (uint32_t)env.ldg(pool + (int64_t)job.in0 * pool_stride + w);
        if (job.in1 >= 0) t += (uint32_t)(int32_t)job.s1 * (uint32_t)env.ldg(pool + (int64_t)job.in1 * pool_stride + w);
        if (job.in2 >= 0) t += (uint32_t)(int32_t)job.s2 * (uint32_t)env.ldg(pool + (int64_t)job.in2 * pool_stride + w);
        bara[w] = (uint16_t)((t + (1u << 20)) >> 21);
    }
}

// L = gadget length, BGBIT = log2 of the gadget base.  `u_out` receives the extracted sample (N+1 words).
template <int L, int BGBIT, class Env>
PEBA1_HD void blind_rotate_team(Env& env, TeamSmem& sm, const TeamTwiddles& tw, const double2* __restrict__ bkfft,
                                int n, const int32_t* __restrict__ pool, int64_t pool_stride, PbsJob job,
                                int32_t* __restrict__ u_out) {
    constexpr int ROWS = 2 * L;
    constexpr uint32_t HALF_BG = 1u << (BGBIT - 1);
    constexpr uint32_t MASK_BG = (1u << BGBIT) - 1;
    uint32_t offset = 0;
#pragma unroll
    for (int p = 1; p <= L; ++p) offset += HALF_BG << (32 - p * BGBIT);
    const int tid = env.tid;

    // ---- K0: prologue + mod-switch to Z_2N: bar = ((uint32)t + 2^20) >> 21
    gate_prologue_modswitch(env, sm.bara, n, pool, pool_stride, job, tid, kTeam);
    env.team_sync();

    // ---- step 3: ACC = (0, X^(2N - barb) * (mu, ..., mu)), mu = 1/8
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

    // ---- step 4: n CMux steps
    for (int i = 0; i < n; ++i) {
        const int abar = sm.bara[i];
        if (abar == 0) continue;  // team-uniform; (X^0 - 1) * ACC = 0 contributes nothing
        double2 sum0[8], sum1[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { sum0[q] = make_double2(0.0, 0.0); sum1[q] = make_double2(0.0, 0.0); }

        // Rows are a real loop (not unrolled): six copies of the transform overflow the instruction cache
        // (ncu r1a: icc hit rate 76%, "no instruction" stalls).  The digit shift is a run-time value instead.
        uint32_t tv[16];
#pragma unroll 1
        for (int r = 0; r < ROWS; ++r) {
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
            double2 x[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int32_t dr = (int32_t)((tv[a] >> shift) & MASK_BG) - (int32_t)HALF_BG;
                const int32_t di = (int32_t)((tv[8 + a] >> shift) & MASK_BG) - (int32_t)HALF_BG;
                x[a] = make_double2(i2d(dr), i2d(di));
            }
            team_fft_forward(env, x, tw, sm, r & 1);
            const double2* bk = bkfft + bkfft_index(ROWS, i, r, 0, 0, tid);
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                const double2 b0 = env.ldg2(bk + (size_t)(2 * k2 + 0) * 64);
                const double2 b1 = env.ldg2(bk + (size_t)(2 * k2 + 1) * 64);
                cfma(sum0[k2], x[k2], b0);
                cfma(sum1[k2], x[k2], b1);
            }
        }
        // inverse transforms and accumulator update (round to nearest, wrap mod 2^32)
        team_fft_inverse(env, sum0, tw, sm, (ROWS + 0) & 1);
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            sm.acc[0][64 * a + tid] = (int32_t)((uint32_t)sm.acc[0][64 * a + tid] + d2u_rn(sum0[a].x));
            sm.acc[0][64 * a + tid + 512] = (int32_t)((uint32_t)sm.acc[0][64 * a + tid + 512] + d2u_rn(sum0[a].y));
        }
        team_fft_inverse(env, sum1, tw, sm, (ROWS + 1) & 1);
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            sm.acc[1][64 * a + tid] = (int32_t)((uint32_t)sm.acc[1][64 * a + tid] + d2u_rn(sum1[a].x));
            sm.acc[1][64 * a + tid + 512] = (int32_t)((uint32_t)sm.acc[1][64 * a + tid + 512] + d2u_rn(sum1[a].y));
        }
        env.team_sync();
    }

    // ---- step 5: sample extraction at index 0
    for (int j = tid; j < kN; j += kTeam)
        u_out[j] = (j == 0) ? sm.acc[0][0] : (int32_t)(0u - (uint32_t)sm.acc[0][kN - j]);
    if (tid == 0) u_out[kN] = sm.acc[1][0];
    env.team_sync();  // smem is reused by the team's next job
}

// K3: transform one integer key polynomial into the layout the MAC reads, pre-scaled by 1/512.
template <class Env>
PEBA1_HD void key_transform_team(Env& env, TeamSmem& sm, const TeamTwiddles& tw, const int32_t* __restrict__ poly,
                                 double2* __restrict__ out_bins /* + bkfft_index(rows,i,r,0,c,0) */) {
    const int tid = env.tid;
    double2 x[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) x[a] = make_double2((double)poly[64 * a + tid], (double)poly[64 * a + tid + 512]);
    team_fft_forward(env, x, tw, sm, 0);
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2)
        out_bins[(size_t)(2 * k2) * 64 + tid] = make_double2(x[k2].x * (1.0 / 512.0), x[k2].y * (1.0 / 512.0));
    env.team_sync();
}

// Host-computed twiddle tables (long double), uploaded once per context.
struct TwiddleTables {
    double2 tw1[64][8];
    double2 tw2[8][8];
};

}  // namespace peba1
