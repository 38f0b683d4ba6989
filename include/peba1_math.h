/*
 * include/peba1_math.h — PEBA1's homomorphic circuit library with the reference's own signatures
 * (/root/reference/include/Math.h:11-51), implemented by libpeba1_tfhe.so as a wire-level circuit builder on the
 * gate recorder: the same Boolean function gate for gate (including the reference's quirks Q1-Q6),
 * but every routine only RECORDS its gates; they execute level by level as batched GPU launches.
 *
 * A translation unit that includes the reference's Math.h instead of this file links against the same symbols
 * (main.cpp and Client.cpp build unchanged; Math.cpp is replaced by this library).
 */
#ifndef PEBA1_MATH_H
#define PEBA1_MATH_H

#include <tfhe/tfhe.h>
#include <vector>

void printSlot64(uint64_t n);

/* ---- encrypted domain (reference src/Math.cpp:27-417) ---- */
void bootsADD1bit(LweSample* result, LweSample* a, LweSample* b, LweSample* carry, const TFheGateBootstrappingCloudKeySet* cloud_key);
void bootsADDNbit(LweSample* result, LweSample* a, LweSample* b, LweSample* carry, const int bitsize, const TFheGateBootstrappingCloudKeySet* cloud_key);
void bootsTwoSComplement(LweSample* result, LweSample* a, const int bitsize, const TFheGateBootstrappingCloudKeySet* cloud_key);
void bootsABS(LweSample* result, LweSample* a, const int bitsize, const TFheGateBootstrappingCloudKeySet* cloud_key);
void bootsSUBNbit(LweSample* result, LweSample* a, LweSample* b, const int bitsize, const TFheGateBootstrappingCloudKeySet* cloud_key);
void bootsShiftLeft(LweSample* result, LweSample* a, const int bitsize, const int n, const TFheGateBootstrappingCloudKeySet* cloud_key);
void bootsShiftRight(LweSample* result, LweSample* a, const int bitsize, const int n, const TFheGateBootstrappingCloudKeySet* cloud_key);
void bootsShiftLeftNR(LweSample* a, const int bitsize, const int n, const TFheGateBootstrappingCloudKeySet* cloud_key);
void bootsMultiply(LweSample* result, LweSample* a, LweSample* b, const int bitsize, const TFheGateBootstrappingCloudKeySet* cloud_key);
double approxEquals(Torus32 a, Torus32 b);
void compare_bit(LweSample* result, const LweSample* a, const LweSample* b, const LweSample* lsb_carry, LweSample* tmp, const TFheGateBootstrappingCloudKeySet* bk);
void minimum(LweSample* result, LweSample* bit, const LweSample* a, const LweSample* b, const int nb_bits, const TFheGateBootstrappingCloudKeySet* bk);
void HE_ManhattanDistance(LweSample* result, std::vector<LweSample*> a, std::vector<LweSample*> b, const int bitsize, const TFheGateBootstrappingCloudKeySet* cloud_key);
void HE_EuclideanDistance(LweSample* result, std::vector<LweSample*> a, std::vector<LweSample*> b, const int bitsize, const TFheGateBootstrappingCloudKeySet* cloud_key);
void Function_f(LweSample* result_b, std::vector<LweSample*> a, std::vector<LweSample*> b, LweSample* bound_match, const int bitsize, const TFheGateBootstrappingCloudKeySet* cloud_key);
void Function_g(LweSample* result, LweSample* result_b, LweSample* r0, LweSample* r1, const int bitsize, const TFheGateBootstrappingCloudKeySet* cloud_key);

/* ---- plaintext mirror (reference src/Math.cpp:427-587) ---- */
uint64_t ADDNbit(uint64_t a, uint64_t b, const int bitsize);
uint64_t TwoSComplement(uint64_t a, const int bitsize);
uint64_t ABS(uint64_t a, const int bitsize);
uint64_t SUBNbit(uint64_t a, uint64_t b, const int bitsize);
uint64_t Multiply(uint64_t a, uint64_t b, const int bitsize);
uint64_t ManhattanDistance(std::vector<uint8_t> a, std::vector<uint8_t> b);
uint64_t ManhattanDistance64(std::vector<uint64_t> a, std::vector<uint64_t> b);
uint64_t EuclideanDistance(std::vector<uint8_t> a, std::vector<uint8_t> b);
uint64_t EuclideanDistance64(std::vector<uint64_t> a, std::vector<uint64_t> b);
uint64_t ManhattanDistanceBitwise(std::vector<uint64_t> a, std::vector<uint64_t> b, const int bitsize);
uint64_t EuclideanDistanceBitwise(std::vector<uint64_t> a, std::vector<uint64_t> b, const int bitsize);
uint64_t Function_f_clear(std::vector<uint64_t> a, std::vector<uint64_t> b, uint64_t bound_match_clear, const int bitsize);
uint64_t Function_g_clear(uint64_t result_b, uint64_t r0, uint64_t r1, const int bitsize);

/* ---- C-ABI views of the same routines for bindings (ctypes / cgo / JNI): plain pointers and sizes ---- */
#ifdef __cplusplus
extern "C" {
#endif
void peba1_math_add1bit(LweSample* result, LweSample* a, LweSample* b, LweSample* carry, const TFheGateBootstrappingCloudKeySet* bk);
void peba1_math_add(LweSample* result, LweSample* a, LweSample* b, LweSample* carry, int bitsize, const TFheGateBootstrappingCloudKeySet* bk);
void peba1_math_twos_complement(LweSample* result, LweSample* a, int bitsize, const TFheGateBootstrappingCloudKeySet* bk);
void peba1_math_abs(LweSample* result, LweSample* a, int bitsize, const TFheGateBootstrappingCloudKeySet* bk);
void peba1_math_sub(LweSample* result, LweSample* a, LweSample* b, int bitsize, const TFheGateBootstrappingCloudKeySet* bk);
void peba1_math_shift_left(LweSample* result, LweSample* a, int bitsize, int n, const TFheGateBootstrappingCloudKeySet* bk);
void peba1_math_shift_right(LweSample* result, LweSample* a, int bitsize, int n, const TFheGateBootstrappingCloudKeySet* bk);
void peba1_math_shift_left_nr(LweSample* a, int bitsize, int n, const TFheGateBootstrappingCloudKeySet* bk);
void peba1_math_compare_bit(LweSample* result, const LweSample* a, const LweSample* b, const LweSample* lsb_carry, LweSample* tmp, const TFheGateBootstrappingCloudKeySet* bk);
/* `width` = 0 reproduces the reference (23, valid for bitsize <= 12); otherwise the product width, e.g. 3*bitsize-1 */
void peba1_math_multiply(LweSample* result, LweSample* a, LweSample* b, int bitsize, int width, const TFheGateBootstrappingCloudKeySet* bk);
void peba1_math_minimum(LweSample* result, LweSample* bit, const LweSample* a, const LweSample* b, int nb_bits, const TFheGateBootstrappingCloudKeySet* bk);
/* a, b: `slots` arrays of `bitsize` samples each */
void peba1_math_manhattan(LweSample* result, LweSample* const* a, LweSample* const* b, int slots, int bitsize, const TFheGateBootstrappingCloudKeySet* bk);
void peba1_math_euclidean(LweSample* result, LweSample* const* a, LweSample* const* b, int slots, int bitsize, const TFheGateBootstrappingCloudKeySet* bk);
void peba1_math_function_f(LweSample* result_b, LweSample* const* a, LweSample* const* b, int slots, LweSample* bound_match, int bitsize, const TFheGateBootstrappingCloudKeySet* bk);
void peba1_math_function_g(LweSample* result, LweSample* result_b, LweSample* r0, LweSample* r1, int bitsize, const TFheGateBootstrappingCloudKeySet* bk);
#ifdef __cplusplus
}
#endif

#endif
