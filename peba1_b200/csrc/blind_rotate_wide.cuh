// blind_rotate_wide.cuh — LOW-LATENCY blind rotation: one CTA of 2l teams (6 x 64 = 384 threads) bootstraps ONE sample.
//
// Why: the throughput kernel gives a bootstrap to one team of 64 threads, which runs its 2l forward transforms one
// after the other: 4.4 ms per bootstrap when the team has an SM to itself.  That is what a narrow circuit level costs
// (the reference's demo runs thousands of levels of 1-7 gates, one face match alone 234 levels of 81 gates), whatever
// the batch throughput is.  The 2l rows of a CMux step are independent until the accumulation, so here team r
// decomposes its own digit level, transforms it and multiplies it with key row r at the same time as the others; the 2l
// partial products are summed by teams 0 and 1 (one output column each), which run the two inverse transforms in
// parallel as well.  Per step: one forward transform + one inverse transform of latency instead of 2l + 2.
//
// 384 threads with up to 168 registers: the (single-row) products stay in registers.  Tensor memory holds the twiddles and
// carries the hand-over of the partial products to the two summing teams (64 columns per warp).
// Shared memory: the accumulator (8 KB, shared by the CTA), the mod-switched input, the key rows (TMA) and 16 KB of
// exchange buffers per team.
//
// Exactness: each partial product is rounded once more than in the fused multiply-add chain of the throughput kernel
// and the 2l partials are added group by group (even teams, odd teams, then the two groups); the error stays ~2^-8
// against the 0.5 that would change the rounded integer, so the outputs are the same words (checked against the
// exact-integer oracle like every other variant).
#pragma once

#include "blind_rotate_tm.cuh"

namespace peba1 {

// Tensor-memory columns of this kernel (256 allocated, per lane quarter): [0,192) the partial products of the up to three
// warps of a quarter (64 each), [192,224) tw1 and [224,256) tw2 of thread tid = 32 (warp & 1) + lane (all warps of a
// quarter share tid).  The twiddles used to sit in 64 registers; next to the product hand-over that made ptxas
// rematerialise some of them from constant memory inside the step loop (thread-indexed LDC).
constexpr int kWideTw1 = 192, kWideTw2 = 224;
// A team's two 8 KB exchange buffers, used alternately by successive transforms like TeamSmemTm2 (both exchanges of a
// transform in one buffer); buffer 0 doubles as the hand-over area of the two summing teams.
struct __align__(16) WideTeamSmem {
    static constexpr bool kAliased = true, kDouble = true, kShfl = false, kClear = false, kXTm = false;
    static constexpr int kTw1 = kWideTw1, kTw2 = kWideTw2;
    double2 ex[2][kM];
    PEBA1_HD double2* e1(int par) { return ex[par]; }
    PEBA1_HD double2* e2(int par) { return ex[par]; }
};
template <int ROWS>
struct __align__(128) WideSmem {
    double2  key[ROWS][2 * 8 * kTeam];  // key row of each team for the coming step: [k2][c][tid], 16 KB, filled by TMA
    WideTeamSmem team[ROWS];
    int32_t  acc[2][kN];
    uint16_t bara[640];
    unsigned long long key_full[ROWS];  // mbarrier per team: "the row has landed"
};

// env: DevEnv-like (tid, team_sync, warp_sync, ldg, ldg2) plus cta_sync(), cta_sync_tm() (a CTA barrier that also orders
// tensor-memory stores before it against loads after it), prod_store(p0, p1) / prod_accumulate(team, p0, p1) (a team's 16
// products to its warps' tensor-memory columns / another team's added to mine), pair_sync() (teams 0 and 1 together),
// key_fetch(dst, src, bytes, mbar) (one thread: TMA bulk copy global -> shared, completion on the mbarrier) and
// key_wait(mbar, parity).
template <int L, int BGBIT, class Env>
PEBA1_HD void blind_rotate_wide(Env& env, WideSmem<2 * L>& sm, int team,
                                const double2* __restrict__ bkfft, int n, const int32_t* __restrict__ pool,
                                int64_t pool_stride, PbsJob job, int32_t* __restrict__ u_out) {
    constexpr int ROWS = 2 * L;
    constexpr uint32_t HALF_BG = 1u << (BGBIT - 1);
    constexpr uint32_t MASK_BG = (1u << BGBIT) - 1;
    uint32_t offset = 0;
#pragma unroll
    for (int p = 1; p <= L; ++p) offset += HALF_BG << (32 - p * BGBIT);
    const int tid = env.tid;
    const int ctid = team * kTeam + tid;  // thread index in the CTA
    constexpr int kCta = ROWS * kTeam;
    WideTeamSmem& mine = sm.team[team];

    gate_prologue_modswitch(env, sm.bara, n, pool, pool_stride, job, ctid, kCta);
    env.cta_sync();
    {
        const int rot = (2 * kN - (int)sm.bara[n]) & (2 * kN - 1);
        const int32_t mu = 1 << 29;
        for (int j = ctid; j < kN; j += kCta) {
            const int idx = (j - rot) & (2 * kN - 1);
            sm.acc[0][j] = 0;
            sm.acc[1][j] = (idx & kN) ? -mu : mu;
        }
    }
    env.cta_sync();

    if (tid == 0) env.key_fetch(sm.key[team], bkfft + bkfft_index(ROWS, 0, team, 0, 0, 0), sizeof(sm.key[team]), &sm.key_full[team]);
    const int poly = team >= L ? 1 : 0;
    const int shift = 32 - (team - poly * L + 1) * BGBIT;
    for (int i = 0; i < n; ++i) {
        // (a void step, abar == 0, is not skipped: its digits are all zero and it leaves the accumulator unchanged; that
        // keeps the one-row-ahead key pipeline free of special cases)
        const int abar = sm.bara[i];
        // ---- phase 1, all teams: own digit level -> forward transform -> product with key row `team`
        double2 x[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const uint32_t tr = (uint32_t)rot_diff<Env>(sm.acc[poly], 64 * a + tid, abar) + offset;
            const uint32_t ti = (uint32_t)rot_diff<Env>(sm.acc[poly], 64 * a + tid + 512, abar) + offset;
            x[a] = make_double2(digit2d((tr >> shift) & MASK_BG, HALF_BG), digit2d((ti >> shift) & MASK_BG, HALF_BG));
        }
        team_fft_forward_tm(env, x, mine);
        // With one warp per scheduler nothing hides the L2 (or HBM) latency of the key values — ncu: a third of all
        // stall samples sat on the first multiplication — and ptxas sinks loads next to their use whatever the source
        // order (an L1 prefetch did not help either).  So each team's row comes by TMA into shared memory one step
        // ahead: requested right after the previous step's products, it lands during the inverse transforms.
        env.key_wait(&sm.key_full[team], (unsigned)i & 1u);
        const double2* kb = sm.key[team];
        double2 p0[8], p1[8];
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
            p0[k2] = cmul(x[k2], kb[(2 * k2 + 0) * kTeam + tid]);
            p1[k2] = cmul(x[k2], kb[(2 * k2 + 1) * kTeam + tid]);
        }
        env.team_sync();  // the team is done reading ex2 (last exchange of the transform) and its key row
        if (tid == 0 && i + 1 < n)
            env.key_fetch(sm.key[team], bkfft + bkfft_index(ROWS, i + 1, team, 0, 0, 0), sizeof(sm.key[team]), &sm.key_full[team]);
        // Hand-over of the partial products THROUGH TENSOR MEMORY (ncu r2g: the shared-memory pipe of an SM that runs this
        // kernel is 70 % busy — 5,500 wavefronts per step, 1,536 of them the products written by 2l teams and read back by
        // two).  Teams of equal parity share their lane quarters (team r = warps 2r, 2r+1), and thread tid of every team
        // holds the same bins: teams 2.. park their 64 words in their warp's columns, team 0 sums the even teams and team 1
        // the odd ones (both columns each), and the two swap one column's partial sum through shared memory (256
        // wavefronts).  Column c = (P_c + P_c+2 + ...) summed in team order inside a group, then even group + odd group.
        if (team >= 2) env.prod_store(p0, p1);
        env.cta_sync_tm();
        // ---- phase 2, teams 0 and 1: column `team` = sum of the 2l partial products, inverse transform, ACC update
        if (team < 2) {
#pragma unroll 1
            for (int r = team + 2; r < ROWS; r += 2) env.prod_accumulate(r, p0, p1);
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) mine.ex[0][64 * k2 + tid] = team == 0 ? p1[k2] : p0[k2];
            env.pair_sync();
            double2 s[8];
            const double2* other = sm.team[team ^ 1].ex[0];
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) s[k2] = team == 0 ? cadd(p0[k2], other[64 * k2 + tid]) : cadd(other[64 * k2 + tid], p1[k2]);
            env.pair_sync();  // both summing teams have read each other's hand-over area before it serves as exchange buffer
            team_fft_inverse_tm(env, s, mine);
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                sm.acc[team][64 * a + tid] = (int32_t)((uint32_t)sm.acc[team][64 * a + tid] + d2u_rn(s[a].x));
                sm.acc[team][64 * a + tid + 512] = (int32_t)((uint32_t)sm.acc[team][64 * a + tid + 512] + d2u_rn(s[a].y));
            }
        }
        env.cta_sync();
    }

    for (int j = ctid; j < kN; j += kCta)
        u_out[j] = (j == 0) ? sm.acc[0][0] : (int32_t)(0u - (uint32_t)sm.acc[0][kN - j]);
    if (ctid == 0) u_out[kN] = sm.acc[1][0];
}

}  // namespace peba1
