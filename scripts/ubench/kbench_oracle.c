/* kbench_oracle.c — the CPU oracle on EXACTLY the workload of kbench.c (same xorshift stream: bootstrapping key,
 * keyswitch key, inputs), printing the same FNV-1a checksum of the output ciphertexts.  Equal checksums mean every one
 * of the batch x (n+1) output words of the GPU path equals the oracle's, at BASELINE configs[1]'s full size.
 * TEST INFRASTRUCTURE (it links oracle/): run by hand or by tests/, never by the product.
 *
 *   make -C oracle
 *   gcc -O2 -Ioracle -o scripts/ubench/kbench_oracle_bin scripts/ubench/kbench_oracle.c -Loracle -ltfhe_oracle \
 *       -Wl,-rpath,'$ORIGIN/../../oracle' -lpthread -lm
 *   scripts/ubench/kbench_oracle_bin [batch=16384] [op=0] [threads=8] [mode=1]        (mode 0: exact integers, slow) */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tfhe_oracle.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd32(void) {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 16);
}

int main(int argc, char** argv) {
    int64_t batch = 16384, op = ORACLE_OP_NAND, threads = 8, mode = ORACLE_FFT;
    for (int i = 1; i < argc; ++i) {
        char* eq = strchr(argv[i], '=');
        if (!eq) return 2;
        *eq = 0;
        const int64_t v = strtoll(eq + 1, NULL, 0);
        if (!strcmp(argv[i], "batch")) batch = v;
        else if (!strcmp(argv[i], "op")) op = v;
        else if (!strcmp(argv[i], "threads")) threads = v;
        else if (!strcmp(argv[i], "mode")) mode = v;
        else return 2;
    }
    oracle_params p;
    if (oracle_default_params(128, &p)) return 3;
    const size_t bk_words = oracle_bk_words(&p), ksk_words = oracle_ksk_words(&p), ct = (size_t)p.n + 1;
    int32_t* bk = (int32_t*)malloc(4 * bk_words);
    int32_t* ksk = (int32_t*)malloc(4 * ksk_words);
    for (size_t i = 0; i < bk_words; ++i) bk[i] = (int32_t)rnd32();
    for (size_t i = 0; i < ksk_words; ++i) ksk[i] = (int32_t)rnd32();
    int32_t* zeros = (int32_t*)calloc((size_t)p.N, 4);  /* the secret keys play no part in a gate */
    oracle_keys* K = oracle_keys_from_arrays(&p, zeros, zeros, bk, ksk);
    if (!K) return 4;
    const int nin = op >= ORACLE_OP_MUX ? 3 : 2;
    int32_t* in[3] = {NULL, NULL, NULL};
    for (int q = 0; q < nin; ++q) {
        in[q] = (int32_t*)malloc(4 * ct * (size_t)batch);
        for (size_t i = 0; i < ct * (size_t)batch; ++i) in[q][i] = (int32_t)rnd32();
    }
    int32_t* out = (int32_t*)malloc(4 * ct * (size_t)batch);
    if (oracle_gate_batch(K, (int)mode, (int)op, in[0], in[1], in[2], out, batch, (int)threads)) return 5;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < ct * (size_t)batch; ++i) h = (h ^ (uint32_t)out[i]) * 1099511628211ull;
    printf("{\"batch\": %lld, \"op\": %lld, \"oracle_mode\": %lld, \"checksum\": \"%016llx\"}\n", (long long)batch,
           (long long)op, (long long)mode, (unsigned long long)h);
    oracle_keys_free(K);
    return 0;
}
