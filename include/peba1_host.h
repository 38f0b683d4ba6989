/*
 * include/peba1_host.h — extensions of the TFHE-compatible front end (libpeba1_tfhe.so) beyond upstream's API:
 * flush control, batching of many circuit instances, statistics, and raw access to samples and keys for bindings
 * and tests.  Everything is plain C ABI.
 */
#ifndef PEBA1_HOST_H
#define PEBA1_HOST_H

#include <stdint.h>
#include "tfhe/tfhe_core.h"

#ifdef __cplusplus
extern "C" {
#endif

/* GPU used by the engines this THREAD creates after the call (and the default of threads that never call it; else
   $PEBA1_DEVICE, else $LOCAL_RANK, else 0).  Recorder state is per key set (one engine = one GPU context, its own wire
   table and lock; a ciphertext handle names its engine), so one process can drive several GPUs with one host thread and
   one key set each; peba1_flush / peba1_batch_begin / _end act on the calling thread's engines only.  A gate whose
   operands belong to different key sets aborts. */
void peba1_host_set_device(int device);

/* TEST HOOK — not a fallback.  Engines created after this call hand the four device operations of the recorder to the
   given table instead of libpeba1_cuda, so the CPU test-suite can check the schedule (levels, widths, slot reuse,
   aliasing) with a bit-level simulator.  Nothing in the product installs a backend; without one (the default) the
   first gate needs a CUDA device and aborts otherwise. */
typedef struct {
    void* user;
    int (*reserve)(void* user, int64_t slots);
    int (*upload_slots)(void* user, const int64_t* slots, int64_t count, const int32_t* words);
    int (*download)(void* user, int64_t slot, int32_t* words);
    int (*gates)(void* user, const void* gates /* const peba1_gate* */, int64_t count);
} peba1_test_backend;
void peba1_host_set_test_backend(const peba1_test_backend* backend);

/* Executes every recorded gate: one peba1_cuda_gates() call (2 kernel launches) per circuit level. */
void peba1_flush(void);
/* While a batch is open, Math.h routines do not flush on exit, so many instances share each level's launch. */
void peba1_batch_begin(void);
void peba1_batch_end(void); /* closes the batch and flushes when the outermost one ends */
/* eager = 1 (default): the outermost Math.h routine flushes before returning, which keeps the reference's per-routine
   wall-clock timings (src/main.cpp:310-312) meaningful.  eager = 0: flush only on decrypt / peba1_flush. */
void peba1_set_eager(int eager);
/* Circuit mode.  0 = FAITHFUL (default): every gate the reference issues is bootstrapped, also
   on trivial inputs (quirk Q6) and when its result is never used, so raw ciphertexts equal a per-gate evaluation.
   1 = FOLD: gates with constant, repeated or complementary operands become constants / aliases / cheaper gates at
   record time, and gates nobody consumes are dropped at flush.  2 = FOLD + the full adder's carry as one MUX
   (2 bootstraps + 1 keyswitch) instead of 3 AND + 2 XOR.  3 = FOLD + 3-input gates: a full adder is XOR3 + MAJ3, one
   bootstrap each (PEBA1_OP_XOR3 / PEBA1_OP_MAJ3: 2(a+b+c)+1/2 and a+b+c, not in upstream's gate set), and the
   squared distance is ONE carry-save column compression over all slots (squarer partial products + accumulator) with
   a single final ripple add, fed by a borrow-chain |a - b| that reproduces the reference's subtraction defect.  Modes 1-3 keep every DECRYPTED result, quirks included; raw ciphertexts differ.
   Changing the mode flushes first. */
void peba1_set_mode(int mode);
int peba1_get_mode(void);

typedef struct {
    uint64_t gates;        /* bootstrapped gates recorded (MUX counts once) */
    uint64_t pbs;          /* bootstraps those gates need (MUX = 2) */
    uint64_t linear;       /* NOT / COPY / CONSTANT calls (never launched: resolved as aliases) */
    uint64_t levels;       /* circuit levels executed = batched launch pairs */
    uint64_t flushes;
    uint64_t max_level_width;
    uint64_t uploads, downloads; /* ciphertexts moved host->device / device->host */
    uint64_t live_slots, peak_slots;
    double flush_seconds;  /* wall time spent inside flushes */
    uint64_t executed_gates, executed_pbs; /* what was actually launched (differs from gates/pbs in fold mode) */
    uint64_t folded;       /* gates replaced by a constant, an alias or a cheaper gate at record time (fold mode) */
    uint64_t dead;         /* recorded gates dropped at flush because nothing consumed them (fold mode) */
    uint64_t device_encryptions; /* ciphertexts produced by peba1_encrypt_array_device */
} peba1_host_stats;
void peba1_get_stats(peba1_host_stats* out);
void peba1_reset_stats(void);

/* Raw sample access (n+1 words: a[0..n-1], b).  _get flushes if the value is still pending. */
int32_t peba1_sample_words(const TFheGateBootstrappingParameterSet* params);
void peba1_sample_get(const LweSample* sample, const TFheGateBootstrappingCloudKeySet* bk, int32_t* words);
void peba1_sample_get_words(const LweSample* sample, int32_t* words); /* as _get, without naming the key */
void peba1_sample_set(LweSample* sample, const int32_t* words);
LweSample* peba1_sample_at(LweSample* array, int32_t index); /* &array[index] for bindings */

/* RANDOMNESS.  Key bits, masks and noise come from ChaCha20 (host/chacha_rng.hpp).  The process-wide generator behind
   new_random_gate_bootstrapping_secret_keyset and bootsSymEncrypt is keyed with 256 bits from getrandom() at first use
   (upstream's default engine is unseeded and the reference never seeds it: a faithful copy of that would give every
   process the same secret key); tfhe_random_generator_setSeed makes it deterministic, as upstream's hook does.
   The functions below that take a `seed` are DETERMINISTIC BY REQUEST — tests and benchmarks; a 64-bit seed cannot carry
   128-bit security. */
/* one ChaCha20 block (8 x 64 bits) for known-answer tests: key[8], 64-bit block counter (state words 12-13), 64-bit
   stream id (state words 14-15) */
void peba1_host_chacha_block(const uint32_t key[8], uint64_t block, uint64_t stream, uint64_t out[8]);
/* Deterministic key generation (new_random_gate_bootstrapping_secret_keyset uses the process-wide generator). */
TFheGateBootstrappingSecretKeySet* peba1_new_secret_keyset_seeded(const TFheGateBootstrappingParameterSet* params, uint64_t seed);
/* A keyset around caller-provided key arrays (copied): lwe_key[n], tlwe_key[N] may be NULL for a cloud-only use. */
TFheGateBootstrappingSecretKeySet* peba1_new_keyset_from_arrays(const TFheGateBootstrappingParameterSet* params,
                                                                const int32_t* lwe_key, const int32_t* tlwe_key,
                                                                const int32_t* bk, const int32_t* ksk);
/* A stand-alone cloud keyset (what a server holds); free it with delete_gate_bootstrapping_cloud_keyset. */
TFheGateBootstrappingCloudKeySet* peba1_new_cloud_keyset_from_arrays(const TFheGateBootstrappingParameterSet* params,
                                                                     const int32_t* bk, const int32_t* ksk);
/* Multi-GPU replicas: a keyset with no key material.  Its GPU context is created without keys; the caller fills the
   device key buffers (peba1_cuda_key_device_buffers on peba1_cuda_context_of) with an NCCL broadcast from the rank that
   holds the real keys, then calls peba1_cuda_keys_mark_loaded.  It cannot encrypt or decrypt. */
TFheGateBootstrappingSecretKeySet* peba1_new_placeholder_keyset(const TFheGateBootstrappingParameterSet* params);
const int32_t* peba1_keyset_bk(const TFheGateBootstrappingCloudKeySet* bk);   /* [n][2l][2][N] */
const int32_t* peba1_keyset_ksk(const TFheGateBootstrappingCloudKeySet* bk);  /* [N][t][2^basebit][n+1] */
/* Batch encryption with an explicit seed: bits[count] -> words[count][n+1]. */
void peba1_encrypt_bits(const TFheGateBootstrappingSecretKeySet* key, const int32_t* bits, int64_t count, uint64_t seed, int32_t* words);
/* array[i] <- fresh encryption of bits[i] made ON THE DEVICE (peba1_cuda_encrypt_bits): synthetic galleries and benchmark
   inputs without host->device traffic.  Deterministic per (seed, i).  Hands the secret LWE key to the GPU: benchmark
   and test use only. */
void peba1_encrypt_array_device(LweSample* array, int32_t nbelems, const int32_t* bits,
                                const TFheGateBootstrappingSecretKeySet* key, uint64_t seed);
/* The same for a list of arrays (arrays[k][i] <- bits[k * nbelems + i]) in ONE launch: a wave of templates. */
void peba1_encrypt_array_list_device(LweSample* const* arrays, int32_t narrays, int32_t nbelems, const int32_t* bits,
                                     const TFheGateBootstrappingSecretKeySet* key, uint64_t seed);
/* arrays[k][i] <- sample (k*nbelems + i) of a dense DEVICE buffer of (n+1)-word samples (e.g. the receive buffer of an
   NCCL broadcast / scatter): one device-side scatter into the ciphertext pool, no bounce through host memory. */
void peba1_sample_set_list_device(LweSample* const* arrays, int32_t narrays, int32_t nbelems, const void* device_words,
                                  const TFheGateBootstrappingCloudKeySet* bk);
/* Pool row of a sample whose value lives on the device (flushes pending work first); -1 for host values, constants and
   negated aliases.  With peba1_cuda_pool_info this lets a caller hand result ciphertexts to a collective in place. */
int64_t peba1_sample_device_slot(const LweSample* sample, const TFheGateBootstrappingCloudKeySet* bk);
void peba1_decrypt_words(const TFheGateBootstrappingSecretKeySet* key, const int32_t* words, int64_t count, int32_t* bits);
/* The libpeba1_cuda context behind a cloud key (created on first use); NULL without a device. */
void* peba1_cuda_context_of(const TFheGateBootstrappingCloudKeySet* bk);

#ifdef __cplusplus
}
#endif
#endif
