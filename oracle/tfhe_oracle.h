/*
 * oracle/tfhe_oracle.h — CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the TFHE gate-bootstrapping algorithm that the
 * reference (lab-incert/peba1) reaches through every `boots*` call in
 * /root/reference/src/Math.cpp (e.g. :34-43, :84, :153, :158, :165, :170,
 * :233, :260-261, :277) and /root/reference/src/main.cpp:21-23,64,78.
 *
 * The arithmetic itself lives in the third-party library `tfhe/tfhe`
 * (https://github.com/tfhe/tfhe), linked as libtfhe-nayuki-portable.so by
 * /root/reference/CMakeLists.txt:10-15.  Its version is NOT pinned by the
 * reference (README.md:17 only links the project page); the 128-bit parameter
 * set n=630 used by main.cpp:20 first appears in upstream v1.1, which is the
 * version restated here.  The library is absent from /root/reference and from
 * this image, so this file restates its published algorithm (SURVEY.md §8c,
 * steps 1-8).
 *
 * PARITY STATUS: "parity unpinned" at the raw-coefficient level — the
 * reference holds no golden vectors, known-answer ciphertexts or seeds
 * (SURVEY.md §4, §8c).  What IS pinned are the decrypted-bit identities that
 * main.cpp checks (:86 round trip, :320 ADD, :352 two's complement, :383 ABS,
 * :422 SUB, :456 MUL, :578 protocol verdict); tests/ checks the oracle against
 * those and against the gate truth tables.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may call into this file.  The product
 * (peba1_b200/, libpeba1_cuda) never links or loads it.
 *
 * Two external-product back ends are provided:
 *   ORACLE_EXACT : exact integer negacyclic convolution mod 2^32 (ground truth)
 *   ORACLE_FFT   : FP64 folded negacyclic FFT with round-to-nearest (fast; it
 *                  is checked against ORACLE_EXACT by tests/test_oracle.py)
 */
#ifndef PEBA1_TFHE_ORACLE_H
#define PEBA1_TFHE_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t Torus32o;

/* Parameter set.  [UPSTREAM-RECALL] new_default_gate_bootstrapping_parameters(). */
typedef struct {
    int32_t n;          /* LWE dimension                    630 (lambda 128) / 500 (lambda 80) */
    int32_t N;          /* ring degree                      1024 */
    int32_t k;          /* TLWE mask polynomials            1    */
    int32_t l;          /* gadget length                    3 / 2 */
    int32_t Bgbit;      /* gadget base bits                 7 / 10 */
    int32_t ks_t;       /* keyswitch digits                 8    */
    int32_t ks_basebit; /* keyswitch base bits              2    */
    double  ks_stdev;   /* LWE / keyswitch noise            2^-15 */
    double  bk_stdev;   /* bootstrapping-key noise          2^-25 */
} oracle_params;

/* Gate opcodes — shared numbering with include/peba1_cuda.h. */
enum {
    ORACLE_OP_NAND = 0, ORACLE_OP_AND = 1, ORACLE_OP_OR = 2, ORACLE_OP_XOR = 3,
    ORACLE_OP_XNOR = 4, ORACLE_OP_NOR = 5, ORACLE_OP_ANDNY = 6, ORACLE_OP_ANDYN = 7,
    ORACLE_OP_ORNY = 8, ORACLE_OP_ORYN = 9, ORACLE_OP_MUX = 10
};
enum { ORACLE_EXACT = 0, ORACLE_FFT = 1 };

typedef struct oracle_keys oracle_keys;

/* fills the default parameter set for a security level (<=80 → n=500 set, <=128 → n=630 set).
   Returns 0, or -1 when lambda > 128 (upstream aborts there). */
int  oracle_default_params(int32_t minimum_lambda, oracle_params* out);

/* Key generation from a 64-bit seed (deterministic).  SURVEY.md §3.3. */
oracle_keys* oracle_keygen(const oracle_params* p, uint64_t seed);
/* Wraps externally generated key material (copied). bk: [n][(k+1)l][k+1][N], ksk: [N][t][base][n+1]. */
oracle_keys* oracle_keys_from_arrays(const oracle_params* p, const int32_t* lwe_key, const int32_t* tlwe_key,
                                     const int32_t* bk, const int32_t* ksk);
void oracle_keys_free(oracle_keys* k);

const oracle_params* oracle_keys_params(const oracle_keys* k);
const int32_t* oracle_lwe_key(const oracle_keys* k);   /* n  bits          */
const int32_t* oracle_tlwe_key(const oracle_keys* k);  /* N  bits (k=1)    */
const int32_t* oracle_bk(const oracle_keys* k);        /* [n][(k+1)l][k+1][N] Torus32 */
const int32_t* oracle_ksk(const oracle_keys* k);       /* [N][t][base][n+1]    Torus32 */
size_t oracle_bk_words(const oracle_params* p);
size_t oracle_ksk_words(const oracle_params* p);

/* modSwitchToTorus32(mu, Msize) / modSwitchFromTorus32(phase, Msize) [UPSTREAM-RECALL]. */
Torus32o oracle_mod_switch_to_torus32(int32_t mu, int32_t Msize);
int32_t  oracle_mod_switch_from_torus32(Torus32o phase, int32_t Msize);

/* LWE samples are int32[n+1]: a[0..n-1] then b.  SURVEY.md §3.4. */
void    oracle_encrypt_bit(const oracle_keys* k, int32_t bit, uint64_t* rng_state, int32_t* ct);
int32_t oracle_decrypt_bit(const oracle_keys* k, const int32_t* ct);
Torus32o oracle_phase(const oracle_keys* k, const int32_t* ct);
void    oracle_trivial(const oracle_params* p, int32_t bit, int32_t* ct);     /* bootsCONSTANT */
void    oracle_not(const oracle_params* p, const int32_t* in, int32_t* out);  /* bootsNOT */

/* Gate prologue: t = (0,c0) + sa*ca + sb*cb; returns 0 or -1 for a bad op.
   For ORACLE_OP_MUX use oracle_gate directly. */
int oracle_gate_prologue(const oracle_params* p, int op, const int32_t* ca, const int32_t* cb, int32_t* t);

/* Bootstrap without keyswitch (steps 2-5): in int32[n+1] -> out int32[N+1] (a'[0..N-1], b'). */
void oracle_bootstrap_woks(const oracle_keys* k, int mode, const int32_t* in, int32_t* out_extracted);
/* Keyswitch (step 6): in int32[N+1] -> out int32[n+1]. */
void oracle_keyswitch(const oracle_keys* k, const int32_t* in_extracted, int32_t* out);
/* Full gate; cc is only read for MUX (a=sel, b=if1, c=if0).  Output may alias an input. */
int oracle_gate(const oracle_keys* k, int mode, int op, const int32_t* ca, const int32_t* cb, const int32_t* cc,
                int32_t* out);
/* Batch of `count` gates of one op over contiguous ciphertext arrays, `threads` worker threads (>=1). */
int oracle_gate_batch(const oracle_keys* k, int mode, int op, const int32_t* ca, const int32_t* cb,
                      const int32_t* cc, int32_t* out, int64_t count, int threads);

/* One external product step, exposed for unit tests:
   acc (int32[2][N]) <- acc + BK_i (.) ((X^abar - 1) * acc).  SURVEY.md §8c step 4. */
void oracle_cmux_step(const oracle_keys* k, int mode, int32_t i, int32_t abar, int32_t* acc);

/* splitmix64-based helpers (exposed so tests can generate reproducible inputs). */
uint64_t oracle_rng_next(uint64_t* state);

#ifdef __cplusplus
}
#endif
#endif
