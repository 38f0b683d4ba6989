/*
 * <tfhe/tfhe_core.h> — types of the TFHE gate-bootstrapping API as PEBA1 uses them, provided by peba1-b200.
 *
 * Written from scratch (upstream tfhe/tfhe is neither vendored by the reference nor installed here).  Only what the
 * reference touches is promised: `Torus32` (include/Math.h:24), `LweSample` as an indexable array element
 * (src/Math.cpp:34 `&tmp[0]`), `TFheGateBootstrappingCloudKeySet::params` (src/Math.cpp:28) and
 * `TFheGateBootstrappingSecretKeySet::cloud` as an embedded member (src/main.cpp:23).
 *
 * The reference relies on these headers to pull in <vector>, <cstdint>, <cstdio>, <cstdlib>, <ctime>, <cmath> and
 * <iostream> (include/Client.h:14 uses std::vector, src/Client.cpp:8 printf, src/main.cpp:37 srand/time).
 */
#ifndef PEBA1_TFHE_CORE_H
#define PEBA1_TFHE_CORE_H

#ifdef __cplusplus
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <iostream>
#include <vector>
#else
#include <stdint.h>
#endif

typedef int32_t Torus32; /* torus element x / 2^32 */

/* One LWE ciphertext.  `a`, `b`, `current_variance` mirror upstream's field names; `peba1_handle` is the recorder
   handle: 0 means the host words (a, b) ARE the value, anything else names a value that lives (or will live) in the
   device ciphertext pool, and the host words are only a cache filled by bootsSymDecrypt / peba1_sample_get. */
typedef struct LweSample {
    Torus32* a;              /* n mask words (host storage, owned by the array allocation) */
    Torus32 b;
    double current_variance; /* not tracked by this engine; kept for layout familiarity */
    uint64_t peba1_handle;
    int32_t peba1_n;         /* number of mask words behind `a` (the parameter set's n) */
    int32_t peba1_pad;
} LweSample;

/* Parameter set: upstream's nested LweParams/TGswParams/TLweParams flattened. */
typedef struct TFheGateBootstrappingParameterSet {
    int32_t ks_t;        /* keyswitch digits          8 */
    int32_t ks_basebit;  /* keyswitch base bits       2 */
    int32_t n;           /* LWE dimension             630 (lambda 128) / 500 (lambda 80) */
    int32_t N;           /* ring degree               1024 */
    int32_t k;           /* TLWE mask polynomials     1 */
    int32_t l;           /* gadget length             3 / 2 */
    int32_t Bgbit;       /* gadget base bits          7 / 10 */
    double ks_stdev;     /* LWE + keyswitch-key noise 2^-15 */
    double bk_stdev;     /* bootstrapping-key noise   2^-25 */
    double max_stdev;    /* largest tolerated output noise (informative) */
    int32_t minimum_lambda;
} TFheGateBootstrappingParameterSet;

struct peba1_key_material; /* opaque: integer keys on the host + the GPU context they are loaded into */

typedef struct TFheGateBootstrappingCloudKeySet {
    const TFheGateBootstrappingParameterSet* params;
    struct peba1_key_material* material; /* bootstrapping key + keyswitch key (public) */
} TFheGateBootstrappingCloudKeySet;

typedef struct TFheGateBootstrappingSecretKeySet {
    const TFheGateBootstrappingParameterSet* params;
    const int32_t* lwe_key;  /* n bits */
    const int32_t* tlwe_key; /* N bits */
    TFheGateBootstrappingCloudKeySet cloud; /* embedded, like upstream (main.cpp:23 takes its address) */
} TFheGateBootstrappingSecretKeySet;

#endif
