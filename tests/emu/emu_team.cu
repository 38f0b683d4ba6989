// emu_team.cu — CPU emulation of ONE team (64 threads) of the blind-rotation kernel (and of one 2l-team CTA of the
// low-latency kernel).
//
// TEST INFRASTRUCTURE: this container has no GPU, so the `-m "not gpu"` tests run the very same
// __host__ __device__ team code (negacyclic_fft.cuh, blind_rotate.cuh) with 64 CPU threads and pthread
// barriers, and compare it with the oracle.  Never loaded by the product path.
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <vector>
#include "blind_rotate.cuh"
#include "blind_rotate_tm.cuh"
#include "blind_rotate_tm_variants.cuh"
#include "blind_rotate_pair.cuh"
#include "blind_rotate_tmk.cuh"
#include "blind_rotate_wide.cuh"
#include "twiddles_host.h"

#include <atomic>

using namespace peba1;

namespace {

struct HostEnv {
    int tid;
    pthread_barrier_t* team;
    pthread_barrier_t* warp;
    void team_sync() { pthread_barrier_wait(team); }
    void warp_sync() { pthread_barrier_wait(warp); }
    // __shfl_xor of a complex double: publish, meet, read the partner's value, meet again before the slot is reused
    double2* shfl_slots = nullptr;  // one per lane of this warp
    double2 shfl_xor_c(double2 v, int mask) {
        const int lane = tid & 31;
        shfl_slots[lane] = v;
        pthread_barrier_wait(warp);
        const double2 r = shfl_slots[lane ^ mask];
        pthread_barrier_wait(warp);
        return r;
    }
    int32_t ldg(const int32_t* p) const { return *p; }
    double2 ldg2(const double2* p) const { return *p; }
    // tensor memory of the tensor-memory variant, emulated as a private array of 32-bit cells per thread
    int sums_col = 0;
    int tv_col = 340;
    int own_col = 400;
    int xpar = 0;
    int flip() { const int p = xpar; xpar ^= 1; return p; }
    uint32_t* tm = nullptr;
    void tmem_ld4(int col, uint32_t (&w)[4]) const { memcpy(w, tm + col, sizeof w); }
    void tmem_st4(int col, const uint32_t (&w)[4]) const { memcpy(tm + col, w, sizeof w); }
    void tmem_ld8(int col, uint32_t (&w)[8]) const { memcpy(w, tm + col, sizeof w); }
    void tmem_st8(int col, const uint32_t (&w)[8]) const { memcpy(tm + col, w, sizeof w); }
    void tmem_ld16(int col, uint32_t (&w)[16]) const { memcpy(w, tm + col, sizeof w); }
    void tmem_st16(int col, const uint32_t (&w)[16]) const { memcpy(tm + col, w, sizeof w); }
    void tmem_ld32(int col, uint32_t (&w)[32]) const { memcpy(w, tm + col, sizeof w); }
    void tmem_st32(int col, const uint32_t (&w)[32]) const { memcpy(tm + col, w, sizeof w); }
    void tmem_wait_ld() const {}
    void tmem_wait_st() const {}
    // key pipeline of blind_rotate_tmk.cuh: the emulation reads the half-row straight from the transformed key
    const double2* key = nullptr;
    int key_rows = 0;
    void key_pump(int) const {}
    void key_wait(int) const {}
    void key_release(int) const {}
    void key_ld16(int h, int c, uint32_t (&k)[16]) const {
        const int srow = h >> 1, half = h & 1;
        for (int e = 0; e < 4; ++e) {
            const double2 v = key[bkfft_index(key_rows, srow / key_rows, srow % key_rows, 4 * half + e, c, tid)];
            c_to_words(v, k[4 * e], k[4 * e + 1], k[4 * e + 2], k[4 * e + 3]);
        }
    }
    // split-phase "exchange buffer free" barrier of the device code: a full team barrier is a valid (stronger) stand-in
    void ex1_acquire() { pthread_barrier_wait(team); }
    void ex1_release() {}
    // ---- the cross-team exchange through tensor memory (TeamSmemTmX): the two 32-lane x 32-column blocks of a team and
    // the fragment layout of tcgen05.ld/st.16x256b.x2 as the device probe (scripts/ubench/tmem_xchg_probe.cu) printed it:
    // thread t, register j of the access at lane half H: lane (t >> 2) + 8 ((j >> 1) & 1) + 16 H, column 8 (j >> 2) + 2 (t & 3) + (j & 1)
    uint32_t (*xt)[32][32] = nullptr;  // [block = the team's warp][lane][column]
    int wb = 0, tw1_col_ = 0, xhi_ = 0, xlo_ = 0;
    int tw1_col() const { return tw1_col_; }
    int xhi() const { return xhi_; }
    int xlo() const { return xlo_; }
    void xtm_exchange_fwd(uint32_t (&v)[32]) {
        const int t = tid & 31;
        pthread_barrier_wait(team);
        memcpy(xt[wb][t], v, sizeof v);
        pthread_barrier_wait(team);
        for (int q = 0; q < 4; ++q) {
            const int S = q >> 1, H = q & 1;
            for (int j = 0; j < 8; ++j)
                v[8 * q + j] = xt[S][(t >> 2) + 8 * ((j >> 1) & 1) + 16 * H][16 * wb + 8 * (j >> 2) + 2 * (t & 3) + (j & 1)];
        }
    }
    void xtm_exchange_inv(uint32_t (&v)[32], uint32_t (&tw)[32]) {
        const int t = tid & 31;
        pthread_barrier_wait(team);
        for (int q = 0; q < 4; ++q) {
            const int S = q >> 1, H = q & 1;
            for (int j = 0; j < 8; ++j)
                xt[S][(t >> 2) + 8 * ((j >> 1) & 1) + 16 * H][16 * wb + 8 * (j >> 2) + 2 * (t & 3) + (j & 1)] = v[8 * q + j];
        }
        pthread_barrier_wait(team);
        memcpy(v, xt[wb][t], sizeof v);
        memcpy(tw, tm + tw1_col_, sizeof tw);
    }
};

struct Shared {
    TeamSmem sm;
    TeamSmemTm sm_tm;
    TeamSmemTm1 sm_tm1;
    TeamSmemTm2 sm_tm2;
    TeamSmemTm2S sm_tm2s;
    TeamSmemTmX sm_tmx;
    uint32_t xt[2][32][32];
    double2 shfl_slots[2][32];
    TeamSmemPair sm_pair;
    TwiddleTables tables;
    pthread_barrier_t team, warp[2];
    // work description
    int mode;  // 0 key transform, 1 bootstrap
    int L, Bgbit, n, rows;
    const int32_t* bk;
    double2* bkfft;
    const int32_t* in_ct;
    int32_t* u_out;
    int64_t count;
};

struct ThreadArg { Shared* S; int tid; };

void* worker(void* p) {
    ThreadArg* A = (ThreadArg*)p;
    Shared* S = A->S;
    HostEnv env{A->tid, &S->team, &S->warp[A->tid >> 5]};
    env.shfl_slots = S->shfl_slots[A->tid >> 5];
    TeamTwiddles tw;
    load_team_twiddles(S->tables, A->tid, tw);
    if (S->mode == 0) {
        const int polys = S->n * S->rows * 2;
        for (int q = 0; q < polys; ++q) {
            const int c = q & 1, ir = q >> 1;
            key_transform_team(env, S->sm, tw, S->bk + (size_t)q * kN,
                               S->bkfft + bkfft_index(S->rows, ir / S->rows, ir % S->rows, 0, c, 0));
        }
    } else if (S->mode == 2) {
        std::vector<uint32_t> tmem(512, 0xDEADBEEFu);
        env.tm = tmem.data();
        env.sums_col = 64 * (A->tid % 3);  // any column block must work
        tm_store_twiddles(env, tw);
        for (int64_t g = 0; g < S->count; ++g) {
            PbsJob job{0, -1, 0, 1, 0};
            const int32_t* pool = S->in_ct + (size_t)g * (S->n + 1);
            int32_t* u = S->u_out + (size_t)g * (kN + 1);
            if (S->L == 3 && S->Bgbit == 7)
                blind_rotate_team_tm<3, 7>(env, S->sm_tm, S->bkfft, S->n, pool, S->n + 1, job, u);
            else
                blind_rotate_team_tm<2, 10>(env, S->sm_tm, S->bkfft, S->n, pool, S->n + 1, job, u);
        }
    } else if (S->mode == 6) {
        std::vector<uint32_t> tmem(512, 0xDEADBEEFu);
        env.tm = tmem.data();
        env.sums_col = 64 * (A->tid % 4);
        tm_store_twiddles<TeamSmemTm2>(env, tw);
        for (int64_t g = 0; g < S->count; ++g) {
            PbsJob job{(int32_t)g, -1, 0, 1, 0};
            int32_t* u = S->u_out + (size_t)g * (kN + 1);
            if (S->L == 3 && S->Bgbit == 7)
                blind_rotate_team_tm<3, 7>(env, S->sm_tm2, S->bkfft, S->n, S->in_ct, S->n + 1, job, u);
            else
                blind_rotate_team_tm<2, 10>(env, S->sm_tm2, S->bkfft, S->n, S->in_ct, S->n + 1, job, u);
        }
    } else if (S->mode == 8) {
        std::vector<uint32_t> tmem(512, 0xDEADBEEFu);
        env.tm = tmem.data();
        env.sums_col = 64 * (A->tid % 4);
        tm_store_twiddles<TeamSmemTm2S>(env, tw);
        for (int64_t g = 0; g < S->count; ++g) {
            PbsJob job{(int32_t)g, -1, 0, 1, 0};
            int32_t* u = S->u_out + (size_t)g * (kN + 1);
            if (S->L == 3 && S->Bgbit == 7)
                blind_rotate_team_tm<3, 7>(env, S->sm_tm2s, S->bkfft, S->n, S->in_ct, S->n + 1, job, u);
            else
                blind_rotate_team_tm<2, 10>(env, S->sm_tm2s, S->bkfft, S->n, S->in_ct, S->n + 1, job, u);
        }
    } else if (S->mode == 9) {
        // cross-team exchange through tensor memory: tid = 32 wb + lane; twiddles as the kernel publishes them
        std::vector<uint32_t> tmem(512, 0xDEADBEEFu);
        env.tm = tmem.data();
        env.sums_col = 64 * (A->tid % 4);
        env.xt = S->xt;
        env.wb = A->tid >> 5;
        env.tw1_col_ = TeamSmemTmX::kTw1 + 32 * env.wb;
        env.xhi_ = 4 * env.wb + (A->tid & 3);
        env.xlo_ = (A->tid & 31) >> 2;
        tm_store8(env, env.tw1_col_, tw.tw1);
        double2 t2[8];
        for (int j = 0; j < 8; ++j) t2[j] = S->tables.tw2[env.xlo_][j];
        tm_store8(env, TeamSmemTmX::kTw2, t2);
        for (int64_t g = 0; g < S->count; ++g) {
            PbsJob job{(int32_t)g, -1, 0, 1, 0};
            int32_t* u = S->u_out + (size_t)g * (kN + 1);
            if (S->L == 3 && S->Bgbit == 7)
                blind_rotate_team_tm<3, 7>(env, S->sm_tmx, S->bkfft, S->n, S->in_ct, S->n + 1, job, u);
            else
                blind_rotate_team_tm<2, 10>(env, S->sm_tmx, S->bkfft, S->n, S->in_ct, S->n + 1, job, u);
        }
    } else if (S->mode == 5) {
        std::vector<uint32_t> tmem(512, 0xDEADBEEFu);
        env.tm = tmem.data();
        env.sums_col = 64 * (A->tid % 5);
        env.tv_col = TeamSmemTm1::kTv + 16 * (A->tid % 5);  // this layout keeps its twiddles in [320, 384)
        tm_store_twiddles<TeamSmemTm1>(env, tw);
        for (int64_t g = 0; g < S->count; ++g) {
            PbsJob job{(int32_t)g, -1, 0, 1, 0};
            int32_t* u = S->u_out + (size_t)g * (kN + 1);
            if (S->L == 3 && S->Bgbit == 7)
                blind_rotate_team_tm<3, 7>(env, S->sm_tm1, S->bkfft, S->n, S->in_ct, S->n + 1, job, u);
            else
                blind_rotate_team_tm<2, 10>(env, S->sm_tm1, S->bkfft, S->n, S->in_ct, S->n + 1, job, u);
        }
    } else if (S->mode == 4) {
        std::vector<uint32_t> tmem(512, 0xDEADBEEFu);
        env.tm = tmem.data();
        env.sums_col = 64 * (A->tid % 4);
        env.key = S->bkfft;
        env.key_rows = S->rows;
        tm_store_twiddles(env, tw);
        for (int64_t g = 0; g < S->count; ++g) {
            PbsJob job{(int32_t)g, -1, 0, 1, 0};
            int32_t* u = S->u_out + (size_t)g * (kN + 1);
            if (S->L == 3 && S->Bgbit == 7)
                blind_rotate_team_tmk<3, 7>(env, S->sm_tm, S->n, S->in_ct, S->n + 1, job, true, u);
            else
                blind_rotate_team_tmk<2, 10>(env, S->sm_tm, S->n, S->in_ct, S->n + 1, job, true, u);
        }
    } else if (S->mode == 3) {
        std::vector<uint32_t> tmem(512, 0xDEADBEEFu);
        env.tm = tmem.data();
        env.sums_col = 128 * (A->tid % 2);
        tm_store_twiddles(env, tw);
        for (int64_t g = 0; g < S->count; g += 2) {  // pool = the input samples, job g reads slot g
            PbsJob job0{(int32_t)g, -1, 0, 1, 0};
            const bool two = g + 1 < S->count;
            PbsJob job1{(int32_t)(two ? g + 1 : g), -1, 0, 1, 0};
            int32_t* u0 = S->u_out + (size_t)g * (kN + 1);
            int32_t* u1 = two ? u0 + (kN + 1) : nullptr;
            if (S->L == 3 && S->Bgbit == 7)
                blind_rotate_pair_tm<3, 7>(env, S->sm_pair, S->bkfft, S->n, S->in_ct, S->n + 1, job0, job1, u0, u1);
            else
                blind_rotate_pair_tm<2, 10>(env, S->sm_pair, S->bkfft, S->n, S->in_ct, S->n + 1, job0, job1, u0, u1);
        }
    } else {
        for (int64_t g = 0; g < S->count; ++g) {
            PbsJob job{0, -1, 0, 1, 0};
            const int32_t* pool = S->in_ct + (size_t)g * (S->n + 1);
            int32_t* u = S->u_out + (size_t)g * (kN + 1);
            if (S->L == 3 && S->Bgbit == 7)
                blind_rotate_team<3, 7>(env, S->sm, tw, S->bkfft, S->n, pool, S->n + 1, job, u);
            else
                blind_rotate_team<2, 10>(env, S->sm, tw, S->bkfft, S->n, pool, S->n + 1, job, u);
        }
    }
    return nullptr;
}

int run(Shared* S) {
    build_twiddle_tables(S->tables);
    pthread_barrier_init(&S->team, nullptr, kTeam);
    pthread_barrier_init(&S->warp[0], nullptr, 32);
    pthread_barrier_init(&S->warp[1], nullptr, 32);
    std::vector<pthread_t> th(kTeam);
    std::vector<ThreadArg> args(kTeam);
    for (int t = 0; t < kTeam; ++t) { args[t] = {S, t}; pthread_create(&th[t], nullptr, worker, &args[t]); }
    for (int t = 0; t < kTeam; ++t) pthread_join(th[t], nullptr);
    pthread_barrier_destroy(&S->team);
    pthread_barrier_destroy(&S->warp[0]);
    pthread_barrier_destroy(&S->warp[1]);
    return 0;
}

// ---- the low-latency kernel code (blind_rotate_wide.cuh): one CTA of 2l teams per bootstrap, 2l x 64 CPU threads
template <int ROWS>
struct WideShared {
    WideSmem<ROWS> sm;
    TwiddleTables tables;
    pthread_barrier_t cta, pair, team[ROWS], warp[2 * ROWS];
    std::atomic<int> fetched[ROWS];  // key rows that have "landed" per team: stands in for the TMA mbarrier
    double2 prod[ROWS][kTeam][16];    // the teams' tensor-memory columns with their partial products
    int n;
    const double2* bkfft;
    const int32_t* in_ct;
    int32_t* u_out;
    int64_t count;
};
template <int ROWS>
struct WideEnv : HostEnv {
    WideShared<ROWS>* S;
    int my_team;
    int waited = 0;  // key rows this thread has consumed
    void cta_sync() { pthread_barrier_wait(&S->cta); }
    void cta_sync_tm() { pthread_barrier_wait(&S->cta); }
    void prod_store(const double2 (&p0)[8], const double2 (&p1)[8]) {
        for (int k = 0; k < 8; ++k) { S->prod[my_team][tid][k] = p0[k]; S->prod[my_team][tid][8 + k] = p1[k]; }
    }
    void prod_accumulate(int r, double2 (&p0)[8], double2 (&p1)[8]) {
        for (int k = 0; k < 8; ++k) { p0[k] = cadd(p0[k], S->prod[r][tid][k]); p1[k] = cadd(p1[k], S->prod[r][tid][8 + k]); }
    }
    void pair_sync() { pthread_barrier_wait(&S->pair); }
    // the device issues an asynchronous bulk copy that completes on the mbarrier; here the copy is done on the spot and
    // published by a release increment, and key_wait spins until as many rows have landed as this thread has asked for
    void key_fetch(void* dst, const void* src, size_t bytes, unsigned long long*) {
        memcpy(dst, src, bytes);
        S->fetched[my_team].fetch_add(1, std::memory_order_release);
    }
    void key_wait(unsigned long long*, unsigned) {
        ++waited;
        while (S->fetched[my_team].load(std::memory_order_acquire) < waited) sched_yield();
    }
};
template <int ROWS> struct WideArg { WideShared<ROWS>* S; int ctid; };

template <int L, int BGBIT>
void* wide_worker(void* p) {
    constexpr int ROWS = 2 * L;
    auto* A = (WideArg<ROWS>*)p;
    WideShared<ROWS>* S = A->S;
    const int team = A->ctid / kTeam, tid = A->ctid % kTeam;
    WideEnv<ROWS> env;
    env.tid = tid;
    env.team = &S->team[team];
    env.warp = &S->warp[2 * team + (tid >> 5)];
    env.S = S;
    env.my_team = team;
    TeamTwiddles tw;
    load_team_twiddles(S->tables, tid, tw);
    std::vector<uint32_t> tmem(512, 0xDEADBEEFu);  // the thread's view of its lane's twiddle columns
    env.tm = tmem.data();
    tm_store8(env, kWideTw1, tw.tw1);
    tm_store8(env, kWideTw2, tw.tw2);
    for (int64_t g = 0; g < S->count; ++g) {
        PbsJob job{(int32_t)g, -1, 0, 1, 0};
        blind_rotate_wide<L, BGBIT>(env, S->sm, team, S->bkfft, S->n, S->in_ct, S->n + 1, job,
                                    S->u_out + (size_t)g * (kN + 1));
        env.cta_sync();  // the kernel's job loop: nobody starts the next prologue before the outputs are written
    }
    return nullptr;
}

template <int L, int BGBIT>
int run_wide(int n, const double2* bkfft, const int32_t* in_ct, int32_t* u_out, int64_t count) {
    constexpr int ROWS = 2 * L;
    auto* S = new WideShared<ROWS>();
    S->n = n; S->bkfft = bkfft; S->in_ct = in_ct; S->u_out = u_out; S->count = count;
    build_twiddle_tables(S->tables);
    pthread_barrier_init(&S->cta, nullptr, ROWS * kTeam);
    pthread_barrier_init(&S->pair, nullptr, 2 * kTeam);
    for (int r = 0; r < ROWS; ++r) {
        pthread_barrier_init(&S->team[r], nullptr, kTeam);
        pthread_barrier_init(&S->warp[2 * r], nullptr, 32);
        pthread_barrier_init(&S->warp[2 * r + 1], nullptr, 32);
        S->fetched[r].store(0);
    }
    std::vector<pthread_t> th(ROWS * kTeam);
    std::vector<WideArg<ROWS>> args(ROWS * kTeam);
    for (int t = 0; t < ROWS * kTeam; ++t) { args[t] = {S, t}; pthread_create(&th[t], nullptr, wide_worker<L, BGBIT>, &args[t]); }
    for (int t = 0; t < ROWS * kTeam; ++t) pthread_join(th[t], nullptr);
    pthread_barrier_destroy(&S->cta);
    pthread_barrier_destroy(&S->pair);
    for (int r = 0; r < ROWS; ++r) {
        pthread_barrier_destroy(&S->team[r]);
        pthread_barrier_destroy(&S->warp[2 * r]);
        pthread_barrier_destroy(&S->warp[2 * r + 1]);
    }
    delete S;
    return 0;
}

}  // namespace

extern "C" {

// the low-latency kernel code (blind_rotate_wide.cuh): a CTA of 2l teams per bootstrap
int peba1_emu_bootstrap_woks_wide(int L, int Bgbit, int n, const double* bkfft, const int32_t* in_ct, int32_t* u_out,
                                  int64_t count) {
    if (L == 3 && Bgbit == 7) return run_wide<3, 7>(n, (const double2*)bkfft, in_ct, u_out, count);
    if (L == 2 && Bgbit == 10) return run_wide<2, 10>(n, (const double2*)bkfft, in_ct, u_out, count);
    return -1;
}

// bk: int32 [n][rows][2][1024] -> bkfft: double2 [n][rows][8][2][64]
int peba1_emu_key_transform(const int32_t* bk, int n, int rows, double* bkfft_out) {
    Shared* S = new Shared();
    S->mode = 0; S->n = n; S->rows = rows; S->bk = bk; S->bkfft = (double2*)bkfft_out;
    int rc = run(S);
    delete S;
    return rc;
}

// in: count samples of n+1 words; out: count extracted samples of 1025 words
int peba1_emu_bootstrap_woks(int L, int Bgbit, int n, const double* bkfft, const int32_t* in_ct, int32_t* u_out,
                             int64_t count) {
    if (!((L == 3 && Bgbit == 7) || (L == 2 && Bgbit == 10))) return -1;
    Shared* S = new Shared();
    S->mode = 1; S->L = L; S->Bgbit = Bgbit; S->n = n; S->rows = 2 * L;
    S->bkfft = (double2*)bkfft; S->in_ct = in_ct; S->u_out = u_out; S->count = count;
    int rc = run(S);
    delete S;
    return rc;
}

// the tensor-memory variant of the kernel code (blind_rotate_tm.cuh)
int peba1_emu_bootstrap_woks_tm(int L, int Bgbit, int n, const double* bkfft, const int32_t* in_ct, int32_t* u_out,
                                int64_t count) {
    if (!((L == 3 && Bgbit == 7) || (L == 2 && Bgbit == 10))) return -1;
    Shared* S = new Shared();
    S->mode = 2; S->L = L; S->Bgbit = Bgbit; S->n = n; S->rows = 2 * L;
    S->bkfft = (double2*)bkfft; S->in_ct = in_ct; S->u_out = u_out; S->count = count;
    int rc = run(S);
    delete S;
    return rc;
}

// two bootstraps per team (blind_rotate_pair.cuh)
int peba1_emu_bootstrap_woks_pair(int L, int Bgbit, int n, const double* bkfft, const int32_t* in_ct, int32_t* u_out,
                                  int64_t count) {
    if (!((L == 3 && Bgbit == 7) || (L == 2 && Bgbit == 10))) return -1;
    Shared* S = new Shared();
    S->mode = 3; S->L = L; S->Bgbit = Bgbit; S->n = n; S->rows = 2 * L;
    S->bkfft = (double2*)bkfft; S->in_ct = in_ct; S->u_out = u_out; S->count = count;
    int rc = run(S);
    delete S;
    return rc;
}

// key rows through tensor memory (blind_rotate_tmk.cuh): consumer side only
int peba1_emu_bootstrap_woks_tmk(int L, int Bgbit, int n, const double* bkfft, const int32_t* in_ct, int32_t* u_out,
                                 int64_t count) {
    if (!((L == 3 && Bgbit == 7) || (L == 2 && Bgbit == 10))) return -1;
    Shared* S = new Shared();
    S->mode = 4; S->L = L; S->Bgbit = Bgbit; S->n = n; S->rows = 2 * L;
    S->bkfft = (double2*)bkfft; S->in_ct = in_ct; S->u_out = u_out; S->count = count;
    int rc = run(S);
    delete S;
    return rc;
}

// single exchange buffer (TeamSmemTm1: the layout of the 10-teams-per-SM kernel)
int peba1_emu_bootstrap_woks_tm1(int L, int Bgbit, int n, const double* bkfft, const int32_t* in_ct, int32_t* u_out,
                                 int64_t count) {
    if (!((L == 3 && Bgbit == 7) || (L == 2 && Bgbit == 10))) return -1;
    Shared* S = new Shared();
    S->mode = 5; S->L = L; S->Bgbit = Bgbit; S->n = n; S->rows = 2 * L;
    S->bkfft = (double2*)bkfft; S->in_ct = in_ct; S->u_out = u_out; S->count = count;
    int rc = run(S);
    delete S;
    return rc;
}

// double-buffered single exchange buffer, no mbarrier (TeamSmemTm2: the default kernel's layout)
int peba1_emu_bootstrap_woks_tm2(int L, int Bgbit, int n, const double* bkfft, const int32_t* in_ct, int32_t* u_out,
                                 int64_t count) {
    if (!((L == 3 && Bgbit == 7) || (L == 2 && Bgbit == 10))) return -1;
    Shared* S = new Shared();
    S->mode = 6; S->L = L; S->Bgbit = Bgbit; S->n = n; S->rows = 2 * L;
    S->bkfft = (double2*)bkfft; S->in_ct = in_ct; S->u_out = u_out; S->count = count;
    int rc = run(S);
    delete S;
    return rc;
}

// TeamSmemTm2S: two exchange buffers, inner exchange by warp shuffles (kernel_variant 10)
int peba1_emu_bootstrap_woks_tm2s(int L, int Bgbit, int n, const double* bkfft, const int32_t* in_ct, int32_t* u_out,
                                  int64_t count) {
    if (!((L == 3 && Bgbit == 7) || (L == 2 && Bgbit == 10))) return -1;
    Shared* S = new Shared();
    S->mode = 8; S->L = L; S->Bgbit = Bgbit; S->n = n; S->rows = 2 * L;
    S->bkfft = (double2*)bkfft; S->in_ct = in_ct; S->u_out = u_out; S->count = count;
    int rc = run(S);
    delete S;
    return rc;
}

// TeamSmemTmX: the cross-team exchange of every transform through tensor memory (kernel_variant 11)
int peba1_emu_bootstrap_woks_tmx(int L, int Bgbit, int n, const double* bkfft, const int32_t* in_ct, int32_t* u_out,
                                 int64_t count) {
    if (!((L == 3 && Bgbit == 7) || (L == 2 && Bgbit == 10))) return -1;
    Shared* S = new Shared();
    S->mode = 9; S->L = L; S->Bgbit = Bgbit; S->n = n; S->rows = 2 * L;
    S->bkfft = (double2*)bkfft; S->in_ct = in_ct; S->u_out = u_out; S->count = count;
    int rc = run(S);
    delete S;
    return rc;
}

}  // extern "C"
