/*
 * <tfhe/tfhe_io.h> — key and ciphertext serialisation with upstream TFHE's function names.
 *
 * The reference includes this header (include/Math.h:5, include/Client.h:5) but calls none of it: its client and
 * server live in one process (src/main.cpp).  A deployment splits them — the client keeps the secret keyset and ships
 * the cloud keyset and ciphertexts — and that is what these functions are for (the same calls upstream's tutorial
 * alice.c / cloud.c / verif.c makes).
 *
 * Formats.  A ciphertext is written the way upstream's write_lweSample does [UPSTREAM-RECALL, unpinned: upstream is
 * not available here to compare byte for byte]: int32 type id 42, n mask words, the body word, double variance
 * (little endian; 2,536 bytes at n = 630).  Parameter sets and keysets use this build's own self-describing container
 * ("PEBA1IO", version, kind, fields, FNV-1a checksum) — NOT upstream's text/binary mix; loading a file written by a
 * real TFHE install is future work, it needs upstream's source to pin the layout.  Readers validate and abort with a
 * message on a bad magic, kind, size or checksum, like every fatal error in this API.
 */
#ifndef PEBA1_TFHE_IO_H
#define PEBA1_TFHE_IO_H
#include <stdio.h>
#include "tfhe_core.h"

#ifdef __cplusplus
extern "C" {
#endif

void export_tfheGateBootstrappingParameterSet_toFile(FILE* F, const TFheGateBootstrappingParameterSet* params);
TFheGateBootstrappingParameterSet* new_tfheGateBootstrappingParameterSet_fromFile(FILE* F);

/* bootstrapping key + keyswitch key (integer form, 113.7 MB at the 128-bit set); the loader uploads and transforms them
   on first use like any other cloud keyset */
void export_tfheGateBootstrappingCloudKeySet_toFile(FILE* F, const TFheGateBootstrappingCloudKeySet* cloud);
TFheGateBootstrappingCloudKeySet* new_tfheGateBootstrappingCloudKeySet_fromFile(FILE* F);

/* cloud keyset + the LWE and ring secret keys */
void export_tfheGateBootstrappingSecretKeySet_toFile(FILE* F, const TFheGateBootstrappingSecretKeySet* keyset);
TFheGateBootstrappingSecretKeySet* new_tfheGateBootstrappingSecretKeySet_fromFile(FILE* F);

/* one ciphertext; export flushes if the value is still pending on the device */
void export_gate_bootstrapping_ciphertext_toFile(FILE* F, const LweSample* sample,
                                                 const TFheGateBootstrappingParameterSet* params);
void import_gate_bootstrapping_ciphertext_fromFile(FILE* F, LweSample* sample,
                                                   const TFheGateBootstrappingParameterSet* params);

#ifdef __cplusplus
}
#endif
#endif
