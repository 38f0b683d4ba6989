/*
 * oracle/refmath_shim.cpp — CPU ORACLE (test infrastructure, NOT the product).
 *
 * C-ABI views (plain pointers and sizes, for ctypes) of the reference's OWN circuit routines.  This file contains no
 * circuit code: every function below forwards to the routine of the same purpose in /root/reference/src/Math.cpp
 * (declared in /root/reference/include/Math.h:11-33), which `make -C oracle ref` compiles unchanged, from where it lies,
 * into oracle/_ref/libpeba1_refmath_cpu.so together with oracle/tfhe_api_cpu.cpp (the gate API on the CPU oracle).
 * The shapes mirror include/peba1_math.h's `peba1_math_*` views of the product, so a test calls both sides alike.
 */
#include "Math.h" /* the reference's header: -I$(REFERENCE)/include */

#include <vector>

typedef const TFheGateBootstrappingCloudKeySet* CK;

extern "C" {
/* Math.cpp:27 */
void refmath_add1bit(LweSample* r, LweSample* a, LweSample* b, LweSample* carry, CK bk) { bootsADD1bit(r, a, b, carry, bk); }
/* Math.cpp:54 */
void refmath_add(LweSample* r, LweSample* a, LweSample* b, LweSample* carry, int bits, CK bk) { bootsADDNbit(r, a, b, carry, bits, bk); }
/* Math.cpp:71 */
void refmath_twos_complement(LweSample* r, LweSample* a, int bits, CK bk) { bootsTwoSComplement(r, a, bits, bk); }
/* Math.cpp:97 */
void refmath_abs(LweSample* r, LweSample* a, int bits, CK bk) { bootsABS(r, a, bits, bk); }
/* Math.cpp:123 */
void refmath_sub(LweSample* r, LweSample* a, LweSample* b, int bits, CK bk) { bootsSUBNbit(r, a, b, bits, bk); }
/* Math.cpp:182,192,201 */
void refmath_shift_left(LweSample* r, LweSample* a, int bits, int n, CK bk) { bootsShiftLeft(r, a, bits, n, bk); }
void refmath_shift_right(LweSample* r, LweSample* a, int bits, int n, CK bk) { bootsShiftRight(r, a, bits, n, bk); }
void refmath_shift_left_nr(LweSample* a, int bits, int n, CK bk) { bootsShiftLeftNR(a, bits, n, bk); }
/* Math.cpp:214 (the reference's product is always 23 samples wide) */
void refmath_multiply(LweSample* r, LweSample* a, LweSample* b, int bits, CK bk) { bootsMultiply(r, a, b, bits, bk); }
/* Math.cpp:259 */
void refmath_compare_bit(LweSample* r, const LweSample* a, const LweSample* b, const LweSample* lsb_carry, LweSample* tmp, CK bk) {
    compare_bit(r, a, b, lsb_carry, tmp, bk);
}
/* Math.cpp:265 */
void refmath_minimum(LweSample* r, LweSample* bit, const LweSample* a, const LweSample* b, int nb_bits, CK bk) { minimum(r, bit, a, b, nb_bits, bk); }
/* Math.cpp:294 */
void refmath_manhattan(LweSample* r, LweSample* const* a, LweSample* const* b, int slots, int bits, CK bk) {
    HE_ManhattanDistance(r, std::vector<LweSample*>(a, a + slots), std::vector<LweSample*>(b, b + slots), bits, bk);
}
/* Math.cpp:333 */
void refmath_euclidean(LweSample* r, LweSample* const* a, LweSample* const* b, int slots, int bits, CK bk) {
    HE_EuclideanDistance(r, std::vector<LweSample*>(a, a + slots), std::vector<LweSample*>(b, b + slots), bits, bk);
}
/* Math.cpp:379 */
void refmath_function_f(LweSample* rb, LweSample* const* a, LweSample* const* b, int slots, LweSample* bound, int bits, CK bk) {
    Function_f(rb, std::vector<LweSample*>(a, a + slots), std::vector<LweSample*>(b, b + slots), bound, bits, bk);
}
/* Math.cpp:390 */
void refmath_function_g(LweSample* r, LweSample* rb, LweSample* r0, LweSample* r1, int bits, CK bk) { Function_g(r, rb, r0, r1, bits, bk); }

/* the plaintext mirror (Math.cpp:427-587), for the decrypted-value checks */
uint64_t refmath_clear_add(uint64_t a, uint64_t b, int bits) { return ADDNbit(a, b, bits); }
uint64_t refmath_clear_twos(uint64_t a, int bits) { return TwoSComplement(a, bits); }
uint64_t refmath_clear_abs(uint64_t a, int bits) { return ABS(a, bits); }
uint64_t refmath_clear_sub(uint64_t a, uint64_t b, int bits) { return SUBNbit(a, b, bits); }
uint64_t refmath_clear_multiply(uint64_t a, uint64_t b, int bits) { return Multiply(a, b, bits); }
uint64_t refmath_clear_manhattan_bitwise(const uint64_t* a, const uint64_t* b, int n, int bits) {
    return ManhattanDistanceBitwise(std::vector<uint64_t>(a, a + n), std::vector<uint64_t>(b, b + n), bits);
}
uint64_t refmath_clear_euclidean_bitwise(const uint64_t* a, const uint64_t* b, int n, int bits) {
    return EuclideanDistanceBitwise(std::vector<uint64_t>(a, a + n), std::vector<uint64_t>(b, b + n), bits);
}
uint64_t refmath_clear_function_f(const uint64_t* a, const uint64_t* b, int n, uint64_t bound, int bits) {
    return Function_f_clear(std::vector<uint64_t>(a, a + n), std::vector<uint64_t>(b, b + n), bound, bits);
}
uint64_t refmath_clear_function_g(uint64_t rb, uint64_t r0, uint64_t r1, int bits) { return Function_g_clear(rb, r0, r1, bits); }
}
