/*
 * include/peba1_cuda.h — C ABI of libpeba1_cuda, the B200 (sm_100a) gate-bootstrapping engine.
 *
 * This is the drop-in boundary for the hot path of lab-incert/peba1: everything the reference
 * reaches through the TFHE gate API (`bootsAND/OR/XOR/XNOR/MUX/...`, linked from
 * libtfhe-nayuki-portable.so by /root/reference/CMakeLists.txt:10-15 and called at
 * /root/reference/src/Math.cpp:34-43,84,158,165,170,233,260-261,277) is executed here as batched
 * CUDA launches.  The TFHE-compatible host shim (the headers under include/tfhe, peba1_b200/host/) records gates and
 * calls these entry points once per circuit level; a foreign-language binding (cgo/JNI/ctypes) would bind
 * exactly these symbols — see INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * PEBA1_E* code, with a message available from peba1_cuda_last_error() (thread-local).  There is NO CPU
 * fallback: without a CUDA device every call fails with PEBA1_ENODEVICE.
 *
 * Data layouts (host side, all little-endian int32 = Torus32):
 *   LWE sample        int32[n+1]            a[0..n-1] then b      (upstream LweSample {a,b})
 *   bootstrapping key int32[n][(k+1)l][k+1][N]   row r = bloc*l+p, polynomial 0 = mask, 1 = body
 *   keyswitch key     int32[N][t][2^basebit][n+1] (digit 0 rows are present and ignored)
 *   extracted sample  int32[N+1]            a'[0..N-1] then b'
 */
#ifndef PEBA1_CUDA_H
#define PEBA1_CUDA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PEBA1_OK          0
#define PEBA1_EINVAL     -1  /* bad argument / unsupported parameter set */
#define PEBA1_ENODEVICE  -2  /* no CUDA device, or device is not sm_100 */
#define PEBA1_ECUDA      -3  /* CUDA runtime error (see last_error) */
#define PEBA1_ENOKEY     -4  /* keys not loaded */
#define PEBA1_ENOMEM     -5

/* Gate opcodes: replace bootsNAND/AND/OR/XOR/XNOR/NOR/ANDNY/ANDYN/ORNY/ORYN/MUX of
   <tfhe/tfhe_gate_bootstrapping_functions.h> (upstream boot-gates.cpp; reference call sites above). */
enum peba1_op {
    PEBA1_OP_NAND = 0, PEBA1_OP_AND = 1, PEBA1_OP_OR = 2, PEBA1_OP_XOR = 3, PEBA1_OP_XNOR = 4,
    PEBA1_OP_NOR = 5, PEBA1_OP_ANDNY = 6, PEBA1_OP_ANDYN = 7, PEBA1_OP_ORNY = 8, PEBA1_OP_ORYN = 9,
    PEBA1_OP_MUX = 10,
    /* 3-input gates, ONE bootstrap each (not in upstream TFHE's API; used by the optimised circuit modes only):
       XOR3 = bootstrap(2(a+b+c) + 1/2), MAJ3 = bootstrap(a+b+c): sum and carry of a full adder */
    PEBA1_OP_XOR3 = 11, PEBA1_OP_MAJ3 = 12
};

/* Input modifiers of one gate operand (the recorder folds bootsNOT / bootsCONSTANT into these). */
#define PEBA1_IN_NEG(i)    (1u << (i))        /* operand i is negated (bootsNOT of a slot)              */
#define PEBA1_IN_CONST(i)  (1u << (4 + (i)))  /* operand i is the trivial sample of bit in[i] (0 or 1)  */

typedef struct peba1_cuda_ctx peba1_cuda_ctx;

/* Mirrors TFheGateBootstrappingParameterSet (reference: cloud_key->params, Math.cpp:28). */
typedef struct {
    int32_t n, N, k, l, Bgbit, ks_t, ks_basebit;
} peba1_cuda_params;

/* One recorded gate. out = op(in[0], in[1] [, in[2]]); slots index the device ciphertext pool. */
typedef struct {
    int32_t  op;      /* enum peba1_op */
    uint32_t flags;   /* PEBA1_IN_NEG / PEBA1_IN_CONST */
    int64_t  in[3];   /* pool slots (or constant bits); in[2] only for MUX (sel, if1, if0) */
    int64_t  out;     /* pool slot written; may equal an input slot */
} peba1_gate;

const char* peba1_cuda_last_error(void);
int peba1_cuda_device_count(void);

/* Context = one GPU + one stream + keys + ciphertext pool.  One per GPU / per rank. */
int peba1_cuda_create(int device, const peba1_cuda_params* params, peba1_cuda_ctx** out);
int peba1_cuda_destroy(peba1_cuda_ctx* ctx);
/* Use an existing cudaStream_t (e.g. torch's current stream) for all launches; NULL restores the own stream. */
int peba1_cuda_set_stream(peba1_cuda_ctx* ctx, void* cuda_stream);
int peba1_cuda_sync(peba1_cuda_ctx* ctx);

/* Keys.  Replaces the bk->bkFFT conversion done at the end of new_random_gate_bootstrapping_secret_keyset
   (reference main.cpp:22): the integer bootstrapping key is uploaded once and transformed on the device
   (kernel K3); the keyswitch key is repacked without its all-zero digit-0 rows. */
int peba1_cuda_load_bootstrap_key(peba1_cuda_ctx* ctx, const int32_t* bk_host);
int peba1_cuda_load_keyswitch_key(peba1_cuda_ctx* ctx, const int32_t* ksk_host);
/* Multi-GPU replication: device buffers of the transformed keys, to be filled by an NCCL broadcast done by
   the caller, then declared valid with peba1_cuda_keys_mark_loaded(). */
int peba1_cuda_key_device_buffers(peba1_cuda_ctx* ctx, void** bkfft, uint64_t* bkfft_bytes,
                                  void** ksk, uint64_t* ksk_bytes);
int peba1_cuda_keys_mark_loaded(peba1_cuda_ctx* ctx);

/* Ciphertext pool: `slots` device-resident LWE samples, each (n+1) words at a padded stride.
   Never-written slots read as trivial false (a=0, b=-1/8) — SURVEY.md §8a quirk Q2. */
int peba1_cuda_pool_reserve(peba1_cuda_ctx* ctx, int64_t slots);
int peba1_cuda_pool_info(peba1_cuda_ctx* ctx, void** device_ptr, int64_t* stride_words, int64_t* capacity);
/* contiguous slot range <-> host array of count*(n+1) words (unpadded) */
int peba1_cuda_upload(peba1_cuda_ctx* ctx, int64_t first_slot, int64_t count, const int32_t* host);
int peba1_cuda_download(peba1_cuda_ctx* ctx, int64_t first_slot, int64_t count, int32_t* host);
/* scattered slots */
int peba1_cuda_upload_slots(peba1_cuda_ctx* ctx, const int64_t* slots, int64_t count, const int32_t* host);
int peba1_cuda_download_slots(peba1_cuda_ctx* ctx, const int64_t* slots, int64_t count, int32_t* host);
/* slots[i] (or first_slot + i when slots is NULL) <- sample i of a dense DEVICE buffer of count*(n+1) words, e.g. what an
   NCCL broadcast or scatter delivered: no bounce through host memory.  Asynchronous on the context stream when slots is
   NULL; d_words must stay valid until the copy has run. */
int peba1_cuda_upload_slots_device(peba1_cuda_ctx* ctx, const int64_t* slots, int64_t first_slot, int64_t count,
                                   const int32_t* d_words);

/* Device-side encryption of `count` bits (what bootsSymEncrypt, src/main.cpp:64, does per bit):
   slot slots[i] (or first_slot + i when slots is NULL) <- LWE encryption of bits[i] under lwe_key (n words, 0/1) with
   noise sigma.  Counter-based generator: the result depends only on (seed, index0 + i).  For synthetic galleries and
   benchmark inputs: it places the SECRET key on the device, which a deployment never does (clients encrypt). */
int peba1_cuda_encrypt_bits(peba1_cuda_ctx* ctx, const int32_t* lwe_key, const int32_t* bits, const int64_t* slots,
                            int64_t first_slot, int64_t count, int64_t index0, uint64_t seed, double sigma);

/* One circuit level: `count` independent gates, executed as one blind-rotation launch (K0+K1) and one
   keyswitch launch (K2).  Asynchronous on the context stream.  An output slot may equal an input slot of the same
   level (every bootstrap reads before any keyswitch writes).  A level of more than max_pbs_per_launch bootstraps
   (option, default 2^18) runs as several launch pairs in sequence; if a later pair would read a slot an earlier pair
   of the same call writes, the call is refused with PEBA1_EINVAL instead of computing a chunk-size-dependent result. */
int peba1_cuda_gates(peba1_cuda_ctx* ctx, const peba1_gate* gates, int64_t count);
/* Convenience: the same op for all gates over slot arrays (in2 may be NULL unless op is MUX). */
int peba1_cuda_gate_batch(peba1_cuda_ctx* ctx, int32_t op, const int64_t* in0, const int64_t* in1,
                          const int64_t* in2, const int64_t* out, int64_t count);
/* End-to-end form with HOST buffers (count*(n+1) words each): upload, run, download, synchronise.
   This is what a per-gate `boots*` caller maps to when it owns host ciphertexts.  It works on private device rows:
   no slot of the ciphertext pool is read or written (peba1_cuda_keyswitch_host likewise). */
int peba1_cuda_gate_batch_host(peba1_cuda_ctx* ctx, int32_t op, const int32_t* ca, const int32_t* cb,
                               const int32_t* cc, int32_t* out, int64_t count);
/* bootsNOT / bootsCOPY / bootsCONSTANT on device slots: out = sign*in + (0, cst); in < 0 means absent. */
int peba1_cuda_linear(peba1_cuda_ctx* ctx, const int64_t* in, const int32_t* sign, const int32_t* cst,
                      const int64_t* out, int64_t count);

/* Pieces of the path, exposed for parity tests against the oracle (device-resident intermediate results are
   copied to host): bootstrap without keyswitch (in: count*(n+1) -> out: count*(N+1)) and keyswitch alone. */
int peba1_cuda_bootstrap_woks_host(peba1_cuda_ctx* ctx, const int32_t* in, int32_t* out_extracted, int64_t count);
int peba1_cuda_keyswitch_host(peba1_cuda_ctx* ctx, const int32_t* in_extracted, int32_t* out, int64_t count);

/* Per-level device times without a sync per level: between _begin and _end every launch pair of peba1_cuda_gates is
   bracketed by CUDA events on the context stream; _end waits for the stream and returns, per launch pair, the gates,
   the bootstraps and the milliseconds of the blind rotation and of the keyswitch (at most `capacity` entries are
   written; *entries = how many there were). */
int peba1_cuda_level_log_begin(peba1_cuda_ctx* ctx);
int peba1_cuda_level_log_end(peba1_cuda_ctx* ctx, int64_t* gates, int64_t* pbs, float* blind_rotate_ms,
                             float* keyswitch_ms, int64_t capacity, int64_t* entries);

/* Instrumentation. */
typedef struct {
    uint64_t launches;        /* kernels launched by this context since creation / reset */
    uint64_t pbs;             /* bootstraps executed */
    uint64_t keyswitches;
    float    last_blind_rotate_ms; /* CUDA-event time of the last K1 launch (needs timing enabled) */
    float    last_keyswitch_ms;    /* ... of the last K2 launch */
    double   sum_blind_rotate_ms, sum_keyswitch_ms; /* accumulated since reset (timing enabled) */
} peba1_cuda_stats;
int peba1_cuda_enable_timing(peba1_cuda_ctx* ctx, int enable);  /* per-kernel CUDA events (syncs per call) */
int peba1_cuda_get_stats(peba1_cuda_ctx* ctx, peba1_cuda_stats* out);
int peba1_cuda_reset_stats(peba1_cuda_ctx* ctx);
/* Measured FP64 FMA peak of this device (dependent-free DFMA loop on all SMs), in TFLOP/s. */
int peba1_cuda_measure_fp64_peak(peba1_cuda_ctx* ctx, double* tflops);
/* Kernel configuration knobs: tuning aids and the measured alternatives kept for comparison.  Every
   setting produces the same words.
     kernel_variant   1 (default): tensor-memory accumulators; persistent 8-team CTAs on two alternating exchange buffers
                      for wide batches, narrow shapes and the low-latency kernel (one CTA per bootstrap) for narrow ones.
                      0: register accumulators.  2: key rows shared by four teams through tensor memory.  3: two bootstraps
                      per team.  4: key rows by TMA + tcgen05.cp into a tensor-memory ring.  5 / 6: 10 teams per CTA on one
                      exchange buffer (fixed grid / persistent).  7: persistent 8-team CTAs on one cross-team exchange buffer
                      guarded by an mbarrier (the default before).  8: the default's wide-batch kernel, also below one
                      full wave.  9: 6 teams per CTA with the accumulators in registers.  10: as 8 with the inner exchange
                      of the transforms by warp shuffles.
     tm_teams         teams per CTA of the narrow shapes (0 = by batch size)      teams_per_cta   the same for variant 0
     wide_max_batch   batches up to this size use the low-latency kernel (-1 = one bootstrap per SM at most)
     ks_split         CTAs per ciphertext in the keyswitch of narrow levels (-1 = by batch size, 0 = never)
     ks_lut           digits per row of the keyswitch's table of row sums: 4 (default; 255 sums per group of 4 digits,
                      1.34 GB of device memory per key, built on the device when first needed), 2 (15 sums, 157 MB) or
                      0 (no table: one key row per non-zero digit, as TFHE does it).  Same output words either way.
     ks_lut_group     ciphertexts per CTA of that kernel (1 or 4) on levels at least ks_lut_group_min wide
                      (-1, default: eight waves of single-ciphertext CTAs, 14,208 on 148 SMs)
     ks_lut_min       levels narrower than this use the per-digit kernels (default 1)
     narrow_two_buffers  1: the narrow shapes use the two-buffer layout of the wide-batch kernel too (default 0)
     ks_group, ks_tile, ks_tile_parts, ks_tile_min: keyswitch forms that measured slower (several ciphertexts per CTA, register
                      tiles); accepted by `make VARIANTS=all` builds only
     pair_min_batch, tmk_stagger_ns, max_pbs_per_launch: see peba1_cuda.cu */
int peba1_cuda_set_option(peba1_cuda_ctx* ctx, const char* name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* PEBA1_CUDA_H */
