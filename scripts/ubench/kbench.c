/* kbench.c — the gate-bootstrap micro-benchmark of bench.py without Python: links libpeba1_cuda only, so one run costs a
 * few seconds of GPU-box time instead of a minute of interpreter and torch start-up.  A TUNING aid (kernel variants,
 * options, batch sizes); the numbers that count are bench.py's.
 *
 *   make -C peba1_b200/csrc libpeba1_cuda.so
 *   gcc -O2 -Iinclude -o scripts/ubench/kbench_bin scripts/ubench/kbench.c -Lpeba1_b200/csrc -lpeba1_cuda \
 *       -Wl,-rpath,'$ORIGIN/../../peba1_b200/csrc'
 *   scripts/ubench/kbench_bin [batch=16384] [steps=3] [warmup=2] [op=0] [name=value ...]     (name = a set_option name)
 *   e.g.  kbench_bin kernel_variant=8        kbench_bin batch=81 steps=20
 *
 * Keys and inputs are pseudo-random words (xorshift, fixed seed): the timing does not depend on their values, and the
 * printed checksum of the outputs is the same for every kernel variant and option (they all compute the same words), so
 * two runs can be compared for equality without an oracle. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "peba1_cuda.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd32(void) {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 16);
}

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_ != 0) {                                                               \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, peba1_cuda_last_error());   \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

int main(int argc, char** argv) {
    int64_t batch = 16384, steps = 3, warmup = 2, op = PEBA1_OP_NAND;
    peba1_cuda_params p = {630, 1024, 1, 3, 7, 8, 2};
    peba1_cuda_ctx* c = NULL;
    CHECK(peba1_cuda_create(0, &p, &c));
    for (int i = 1; i < argc; ++i) {
        char* eq = strchr(argv[i], '=');
        if (!eq) { fprintf(stderr, "expected name=value, got %s\n", argv[i]); return 2; }
        *eq = 0;
        const int64_t v = strtoll(eq + 1, NULL, 0);
        if (!strcmp(argv[i], "batch")) batch = v;
        else if (!strcmp(argv[i], "steps")) steps = v;
        else if (!strcmp(argv[i], "warmup")) warmup = v;
        else if (!strcmp(argv[i], "op")) op = v;
        else CHECK(peba1_cuda_set_option(c, argv[i], v));
    }
    const int nin = op >= PEBA1_OP_MUX ? 3 : 2;

    const size_t bk_words = (size_t)p.n * 2 * p.l * 2 * p.N, ksk_words = (size_t)p.N * p.ks_t * 4 * (p.n + 1);
    int32_t* buf = (int32_t*)malloc(sizeof(int32_t) * (ksk_words > bk_words ? ksk_words : bk_words));
    if (!buf) return 3;
    for (size_t i = 0; i < bk_words; ++i) buf[i] = (int32_t)rnd32();
    CHECK(peba1_cuda_load_bootstrap_key(c, buf));
    for (size_t i = 0; i < ksk_words; ++i) buf[i] = (int32_t)rnd32();
    CHECK(peba1_cuda_load_keyswitch_key(c, buf));

    CHECK(peba1_cuda_pool_reserve(c, (nin + 1) * batch));
    const size_t ct = (size_t)p.n + 1;
    int32_t* in = (int32_t*)malloc(sizeof(int32_t) * ct * (size_t)batch);
    for (int q = 0; q < nin; ++q) {
        for (size_t i = 0; i < ct * (size_t)batch; ++i) in[i] = (int32_t)rnd32();
        CHECK(peba1_cuda_upload(c, q * batch, batch, in));
    }
    int64_t* slots = (int64_t*)malloc(sizeof(int64_t) * 4 * (size_t)batch);
    for (int64_t i = 0; i < 4 * batch; ++i) slots[i] = i;

    CHECK(peba1_cuda_enable_timing(c, 1));
    for (int64_t s = 0; s < warmup + steps; ++s) {
        if (s == warmup) CHECK(peba1_cuda_reset_stats(c));
        CHECK(peba1_cuda_gate_batch(c, (int32_t)op, slots, slots + batch, nin == 3 ? slots + 2 * batch : NULL,
                                    slots + nin * batch, batch));
        CHECK(peba1_cuda_sync(c));
    }
    peba1_cuda_stats st;
    CHECK(peba1_cuda_get_stats(c, &st));
    CHECK(peba1_cuda_download(c, nin * batch, batch, in));
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < ct * (size_t)batch; ++i) h = (h ^ (uint32_t)in[i]) * 1099511628211ull;

    const double br = st.sum_blind_rotate_ms / (double)steps, ks = st.sum_keyswitch_ms / (double)steps;
    printf("{\"batch\": %lld, \"op\": %lld, \"steps\": %lld, \"blind_rotate_ms\": %.4f, \"keyswitch_ms\": %.4f, "
           "\"pbs_per_s\": %.1f, \"gates_per_s\": %.1f, \"launches_per_step\": %.1f, \"checksum\": \"%016llx\"}\n",
           (long long)batch, (long long)op, (long long)steps, br, ks, 1e3 * (double)st.pbs / (double)steps / br,
           1e3 * (double)batch / (br + ks), (double)st.launches / (double)steps, (unsigned long long)h);
    free(buf); free(in); free(slots);
    CHECK(peba1_cuda_destroy(c));
    return 0;
}
