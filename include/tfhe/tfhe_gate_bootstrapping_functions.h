/*
 * <tfhe/tfhe_gate_bootstrapping_functions.h> — the TFHE gate API subset PEBA1 calls, implemented by
 * libpeba1_tfhe.so as a RECORDING front end of the B200 engine: boots* gates append nodes to a DAG; the DAG is
 * levelised and executed as one batched CUDA launch pair (blind rotation + keyswitch) per level when a result is
 * needed (bootsSymDecrypt, peba1_flush, or the end of an outermost Math.h routine).
 *
 * Signatures follow the reference's call sites:
 *   new_default_gate_bootstrapping_parameters     src/main.cpp:21
 *   new_random_gate_bootstrapping_secret_keyset   src/main.cpp:22
 *   new/delete_gate_bootstrapping_ciphertext_array src/main.cpp:59, src/Math.cpp:28,47
 *   bootsSymEncrypt / bootsSymDecrypt             src/main.cpp:64,78
 *   bootsCONSTANT/NOT/COPY                        src/Math.cpp:58,153,31
 *   bootsAND/OR/XOR/XNOR                          src/Math.cpp:38,170,34,260
 *   bootsMUX                                      src/Math.cpp:261,277
 * All gates return void and abort on fatal errors, like upstream.  Outputs may alias inputs (src/Math.cpp:272).
 */
#ifndef PEBA1_TFHE_GATE_BOOTSTRAPPING_FUNCTIONS_H
#define PEBA1_TFHE_GATE_BOOTSTRAPPING_FUNCTIONS_H

#include "tfhe_core.h"

#ifdef __cplusplus
extern "C" {
#endif

TFheGateBootstrappingParameterSet* new_default_gate_bootstrapping_parameters(int32_t minimum_lambda);
void delete_gate_bootstrapping_parameters(TFheGateBootstrappingParameterSet* params);

TFheGateBootstrappingSecretKeySet* new_random_gate_bootstrapping_secret_keyset(const TFheGateBootstrappingParameterSet* params);
void delete_gate_bootstrapping_secret_keyset(TFheGateBootstrappingSecretKeySet* keyset);
void delete_gate_bootstrapping_cloud_keyset(TFheGateBootstrappingCloudKeySet* keyset); /* no-op for an embedded keyset */

LweSample* new_gate_bootstrapping_ciphertext(const TFheGateBootstrappingParameterSet* params);
LweSample* new_gate_bootstrapping_ciphertext_array(int32_t nbelems, const TFheGateBootstrappingParameterSet* params);
void delete_gate_bootstrapping_ciphertext(LweSample* sample);
void delete_gate_bootstrapping_ciphertext_array(int32_t nbelems, LweSample* samples);

void bootsSymEncrypt(LweSample* result, int32_t message, const TFheGateBootstrappingSecretKeySet* key);
int32_t bootsSymDecrypt(const LweSample* sample, const TFheGateBootstrappingSecretKeySet* key);

void bootsCONSTANT(LweSample* result, int32_t value, const TFheGateBootstrappingCloudKeySet* bk);
void bootsNOT(LweSample* result, const LweSample* ca, const TFheGateBootstrappingCloudKeySet* bk);
void bootsCOPY(LweSample* result, const LweSample* ca, const TFheGateBootstrappingCloudKeySet* bk);

void bootsNAND(LweSample* result, const LweSample* ca, const LweSample* cb, const TFheGateBootstrappingCloudKeySet* bk);
void bootsOR(LweSample* result, const LweSample* ca, const LweSample* cb, const TFheGateBootstrappingCloudKeySet* bk);
void bootsAND(LweSample* result, const LweSample* ca, const LweSample* cb, const TFheGateBootstrappingCloudKeySet* bk);
void bootsXOR(LweSample* result, const LweSample* ca, const LweSample* cb, const TFheGateBootstrappingCloudKeySet* bk);
void bootsXNOR(LweSample* result, const LweSample* ca, const LweSample* cb, const TFheGateBootstrappingCloudKeySet* bk);
void bootsNOR(LweSample* result, const LweSample* ca, const LweSample* cb, const TFheGateBootstrappingCloudKeySet* bk);
void bootsANDNY(LweSample* result, const LweSample* ca, const LweSample* cb, const TFheGateBootstrappingCloudKeySet* bk);
void bootsANDYN(LweSample* result, const LweSample* ca, const LweSample* cb, const TFheGateBootstrappingCloudKeySet* bk);
void bootsORNY(LweSample* result, const LweSample* ca, const LweSample* cb, const TFheGateBootstrappingCloudKeySet* bk);
void bootsORYN(LweSample* result, const LweSample* ca, const LweSample* cb, const TFheGateBootstrappingCloudKeySet* bk);
/* result = a ? b : c */
void bootsMUX(LweSample* result, const LweSample* a, const LweSample* b, const LweSample* c, const TFheGateBootstrappingCloudKeySet* bk);

/* upstream's seeding hook; the reference never calls it, so keygen/encryption use a fixed default seed */
void tfhe_random_generator_setSeed(uint32_t* values, int32_t size);

#ifdef __cplusplus
}
#endif
#endif
