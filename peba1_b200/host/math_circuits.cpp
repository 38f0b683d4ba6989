// math_circuits.cpp — PEBA1's Math.h routines (reference src/Math.cpp:27-417) as a wire-level circuit builder.
//
// The reference builds every circuit out of LweSample temporaries, COPYs and per-gate library calls.  Here a routine
// reads the wires its inputs currently name, combines them with gate() / mux() on `Bit` values (RAII handles on
// recorder wires), and binds the result wires to the output samples: no temporaries, no copies, nothing executed
// until the recorder flushes.  The Boolean function is the reference's, gate for gate — same gate types, same
// operands, same constants, including gates on trivial inputs (Q6) and the defects documented in SURVEY.md §8a — so
// that decrypted AND raw-ciphertext parity with a faithful per-gate evaluation holds.  Gate counts per routine are
// checked against the reference's structure in tests/test_circuits.py.
#include <math.h>
#include <stdio.h>

#include <deque>
#include <utility>

#include "../../include/peba1_math.h"
#include "recorder.hpp"

using namespace peba1;

namespace {

// One ciphertext value.  Copying shares the wire; destruction drops the reference.
class Bit {
  public:
    Bit() : h_(kFalse) {}
    static Bit constant(bool v) { return Bit(v ? kTrue : kFalse); }
    static Bit adopt(Handle h) { return Bit(h); }  // takes over one reference
    static Bit of(Engine* e, const LweSample* s) {
        const Handle h = sample_handle(e, s);
        retain(h);
        return Bit(h);
    }
    Bit(const Bit& o) : h_(o.h_) { retain(h_); }
    Bit(Bit&& o) noexcept : h_(o.h_) { o.h_ = kFalse; }
    Bit& operator=(Bit o) noexcept {
        std::swap(h_, o.h_);
        return *this;
    }
    ~Bit() { release(h_); }
    Bit operator!() const {  // bootsNOT: free
        retain(h_);
        return Bit(negate(h_));
    }
    Handle handle() const { return h_; }
    void store(LweSample* s) const {  // the sample gets its own reference
        retain(h_);
        sample_assign(s, h_);
    }

  private:
    explicit Bit(Handle h) : h_(h) {}
    Handle h_;
};

using Word = std::vector<Bit>;  // LSB first

struct Builder {
    Engine* e;
    Bit gate(int op, const Bit& a, const Bit& b) const { return Bit::adopt(record_gate(e, op, a.handle(), b.handle())); }
    Bit AND(const Bit& a, const Bit& b) const { return gate(PEBA1_OP_AND, a, b); }
    Bit OR(const Bit& a, const Bit& b) const { return gate(PEBA1_OP_OR, a, b); }
    Bit XOR(const Bit& a, const Bit& b) const { return gate(PEBA1_OP_XOR, a, b); }
    Bit XNOR(const Bit& a, const Bit& b) const { return gate(PEBA1_OP_XNOR, a, b); }
    Bit MUX(const Bit& s, const Bit& t, const Bit& f) const {
        return Bit::adopt(record_mux(e, s.handle(), t.handle(), f.handle()));
    }
    // 3-input gates, one bootstrap each (mode 3 only)
    Bit XOR3(const Bit& a, const Bit& b, const Bit& c) const {
        return Bit::adopt(record_gate3(e, PEBA1_OP_XOR3, a.handle(), b.handle(), c.handle()));
    }
    Bit MAJ3(const Bit& a, const Bit& b, const Bit& c) const {
        return Bit::adopt(record_gate3(e, PEBA1_OP_MAJ3, a.handle(), b.handle(), c.handle()));
    }
    Word load(const LweSample* s, int bits) const {
        Word w;
        w.reserve((size_t)bits);
        for (int i = 0; i < bits; ++i) w.push_back(Bit::of(e, &s[i]));
        return w;
    }

    // reference bootsADD1bit (Math.cpp:27-50): sum = (a^b)^c ; c' = ((a&b) ^ (a&c)) ^ (c&b) — 4 XOR + 3 AND
    Bit full_adder(const Bit& a, const Bit& b, Bit& carry) const {
        const Bit c0 = carry;
        if (recorder_mode() >= 3) {  // sum and carry in one bootstrap each
            carry = MAJ3(a, b, c0);
            return XOR3(a, b, c0);
        }
        if (recorder_mode() >= 2) {  // optimised adder: carry' = (a == b) ? a : carry — one MUX on the shared a^b
            const Bit t = XOR(a, b);
            Bit sum2 = XOR(t, c0);
            carry = MUX(t, c0, a);
            return sum2;
        }
        Bit sum = XOR(XOR(a, b), c0);
        const Bit ab = AND(a, b);
        const Bit ac = AND(a, c0);
        const Bit t = XOR(ab, ac);
        const Bit cb = AND(c0, b);
        carry = XOR(t, cb);
        return sum;
    }
    // reference bootsADDNbit (Math.cpp:54-67): ripple add of `bits` bits, carry-in trivially 0 (Q5); returns the
    // `bits` sum bits and leaves the carry-out in `carry`
    Word add(const Word& a, const Word& b, int bits, Bit& carry) const {
        carry = Bit::constant(false);
        Word r;
        r.reserve((size_t)bits);
        for (int i = 0; i < bits; ++i) r.push_back(full_adder(a[i], b[i], carry));
        return r;
    }
    // reference bootsTwoSComplement (Math.cpp:71-93): XOR every bit with a trivial 1 (bootstrapped, Q6), then + 1
    Word twos_complement(const Word& a, int bits) const {
        const Bit one = Bit::constant(true);
        Word inv, inc;
        for (int i = 0; i < bits; ++i) {
            inv.push_back(XOR(a[i], one));
            inc.push_back(Bit::constant(i == 0));
        }
        Bit carry;
        return add(inv, inc, bits, carry);
    }
    // reference bootsABS (Math.cpp:97-119): mask = sign bit replicated; (a + mask) ^ mask
    Word abs(const Word& a, int bits) const {
        const Word mask((size_t)bits, a[bits - 1]);
        Bit carry;
        const Word t = add(a, mask, bits, carry);
        Word r;
        for (int i = 0; i < bits; ++i) r.push_back(XOR(t[i], mask[i]));
        return r;
    }
    // reference bootsSUBNbit (Math.cpp:123-180): |a - b| on bits+1 bits via two's complement and a carry select.
    // Quirk Q1 is reproduced: bit `bits` of the complement is forced to 1 even when b = 0.
    Word sub(const Word& a, const Word& b, int bits) const {
        Word a_long(a.begin(), a.begin() + bits);
        a_long.push_back(Bit::constant(false));
        Word b_neg = twos_complement(b, bits);
        b_neg.push_back(Bit::constant(true));
        Bit carry;
        const Word s = add(a_long, b_neg, bits + 1, carry);
        const Bit not_carry = !carry;
        Word neg_part;
        for (int i = 0; i <= bits; ++i) neg_part.push_back(AND(s[i], not_carry));
        const Word neg_fixed = twos_complement(neg_part, bits + 1);
        Word r;
        Word pos_part;
        for (int i = 0; i <= bits; ++i) pos_part.push_back(AND(s[i], carry));
        for (int i = 0; i <= bits; ++i) r.push_back(OR(neg_fixed[i], pos_part[i]));
        return r;
    }
    // Mode 3 only: the low `bits` bits of bootsSUBNbit(a, b) — what the squared distance feeds to the multiplier
    // (Math.cpp:352-356) — without the reference's two two's-complement passes.  d = a - b by a borrow chain
    // (XOR3 / MAJ3), sign s = borrow out, result = (d ^ s) + s.  The reference's defect Q1 (for b = 0 it yields
    // 2^bits - a, whose low bits are -a) is reproduced by forcing s when b is all zero.  Checked against the reference
    // formula on every (a, 0), (0, b) and thousands of random pairs (tests/test_circuits.py).
    Word absdiff_low(const Word& a, const Word& b, int bits) const {
        Bit borrow = Bit::constant(false);
        Word d;
        for (int i = 0; i < bits; ++i) {
            d.push_back(XOR3(a[i], b[i], borrow));
            borrow = MAJ3(!a[i], b[i], borrow);
        }
        Bit any_b = b[0];
        for (int i = 1; i < bits; ++i) any_b = OR(any_b, b[i]);
        const Bit s = OR(borrow, !any_b);
        Word r;
        Bit carry = s;
        for (int i = 0; i < bits; ++i) {
            const Bit t = XOR(d[i], s);
            r.push_back(XOR(t, carry));
            if (i + 1 < bits) carry = AND(t, carry);
        }
        return r;
    }
    // reference bootsMultiply (Math.cpp:214-250): shift-and-add over `width` bits (23 in the reference); each of the
    // `bits` partial products is accumulated with a (width-1)-bit ripple add, so bit width-1 stays trivially 0 (Q5)
    Word multiply(const Word& a, const Word& b, int bits, int width) const {
        Word total((size_t)width, Bit::constant(false));
        Word partial((size_t)width, Bit::constant(false));
        for (int i = 0; i < bits; ++i) {
            for (int j = 0; j < i; ++j) partial[j] = Bit::constant(false);
            for (int j = 0; j < bits; ++j) partial[j + i] = AND(a[j], b[i]);
            const Word ongoing = total;
            Bit carry;
            const Word s = add(partial, ongoing, width - 1, carry);
            for (int j = 0; j < width - 1; ++j) total[j] = s[j];
        }
        return total;
    }
    // reference compare_bit (Math.cpp:259-262): equal ? carry : a
    Bit compare_bit(const Bit& a, const Bit& b, const Bit& carry) const { return MUX(XNOR(a, b), carry, a); }
    // reference minimum (Math.cpp:265-286): returns the comparator bit (1 iff a > b, Q3) and fills `mn` with min(a,b)
    Bit minimum(const Word& a, const Word& b, int bits, Word& mn) const {
        Bit c = Bit::constant(false);
        for (int i = 0; i < bits; ++i) c = compare_bit(a[i], b[i], c);
        mn.clear();
        for (int i = 0; i < bits; ++i) mn.push_back(MUX(c, b[i], a[i]));
        return c;
    }
};

void store_word(LweSample* out, const Word& w, int bits) {
    for (int i = 0; i < bits; ++i) w[i].store(&out[i]);
}

[[noreturn]] void bad_size(const char* routine, const char* what, int bitsize, int width) {
    // The reference overruns its heap arrays in these cases (Math.cpp:216 hard-codes 23 result samples, :337 24 accumulator
    // samples); a drop-in must not turn that into silent undefined behaviour of its own.
    fprintf(stderr, "peba1: fatal: %s: %s (bitsize %d, width %d)\n", routine, what, bitsize, width);
    abort();
}

int multiply_width(int bitsize, int width) {
    if (bitsize < 1) bad_size("bootsMultiply", "bitsize must be positive", bitsize, width);
    const int w = width > 0 ? width : 23;  // the reference's hard-coded length (Math.cpp:216)
    // partial product i touches result samples i .. i + bitsize - 1: the last needs 2 * bitsize - 1 samples
    if (2 * bitsize - 1 > w) bad_size("bootsMultiply", "the product needs 2 * bitsize - 1 samples, more than the result width", bitsize, w);
    return w;
}

// the distance circuits accumulate on 23 samples (24 allocated), squares of (bitsize + 1)-bit differences: bitsize 8 in
// the reference (main.cpp:20 `bitsize = 8`); anything that makes the multiplier or the 3 * bitsize comparator index
// outside those arrays is refused
void check_distance_bitsize(const char* routine, int bitsize) {
    if (bitsize < 8 || 2 * bitsize - 1 > 23) bad_size(routine, "bitsize must be 8..12 (23-sample accumulator, 3 * bitsize >= 24 comparator)", bitsize, 23);
}

void check_vectors(const char* what, size_t na, size_t nb) {
    // the reference only reports these conditions and carries on (Math.cpp:295-300)
    if (na == 0 || nb == 0) perror(what);
    if (na != nb) perror("The two ciphertexts do not have the same size in a distance computation\n");
}

// Mode 3 body of the squared distance: every term of every slot's square (the diagonal bits d_i at weight 2^2i, the
// cross products d_i d_j, i < j, at weight 2^(i+j+1)) and the incoming accumulator bits are thrown into 23 weight
// columns, the columns are compressed with full adders (XOR3 + MAJ3: 3 bits of a column -> 1 bit + 1 carry) until at
// most two bits are left in each, and ONE ripple add produces the sum.  Same value as the reference's
// multiply-then-add chain: the distance is < 2^23 (128 * 255^2), so no carry leaves column 22 and bit 23 stays
// untouched (Q5); d = the low bits of the reference's SUBNbit result, quirk Q1 included (Builder::absdiff_low).
void euclid_columns(const Builder& B, Word& acc, LweSample* const* a, LweSample* const* b, int slots, int bitsize) {
    const int width = 23;
    std::vector<std::deque<Bit>> col((size_t)width);
    for (int j = 0; j < width; ++j)
        if (acc[j].handle() != kFalse) col[j].push_back(acc[j]);
    for (int s = 0; s < slots; ++s) {
        const Word d = B.absdiff_low(B.load(b[s], bitsize), B.load(a[s], bitsize), bitsize);
        for (int i = 0; i < bitsize; ++i) {
            if (2 * i < width) col[2 * i].push_back(d[i]);
            for (int j = i + 1; j < bitsize; ++j)
                if (i + j + 1 < width) col[i + j + 1].push_back(B.AND(d[i], d[j]));
        }
    }
    // Wallace passes: every pass compresses ALL columns by one layer of full adders (bits of the current layer only; the
    // carries join the next layer of the column above), so the depth is ~log_1.5 of the tallest column (15 passes for
    // 128 slots) instead of a chain that walks the columns one after the other (84 narrow levels).  Same adder count.
    for (bool more = true; more;) {
        std::vector<std::deque<Bit>> next((size_t)width);
        for (int j = 0; j < width; ++j) {
            std::deque<Bit>& c = col[j];
            while (c.size() >= 3) {
                const Bit x = c.front(); c.pop_front();
                const Bit y = c.front(); c.pop_front();
                const Bit z = c.front(); c.pop_front();
                next[j].push_back(B.XOR3(x, y, z));
                if (j + 1 < width) next[j + 1].push_back(B.MAJ3(x, y, z));
            }
            while (!c.empty()) { next[j].push_back(c.front()); c.pop_front(); }  // 0..2 leftovers move on unchanged
        }
        col.swap(next);
        more = false;
        for (int j = 0; j < width; ++j) more |= col[j].size() > 2;
    }
    Word p, q;
    for (int j = 0; j < width; ++j) {
        p.push_back(col[j].size() > 0 ? col[j][0] : Bit::constant(false));
        q.push_back(col[j].size() > 1 ? col[j][1] : Bit::constant(false));
    }
    Bit carry;
    const Word sum = B.add(p, q, width, carry);
    for (int j = 0; j < width; ++j) acc[j] = sum[j];
}

// shared body of HE_EuclideanDistance / Function_f: `acc` holds the 24 accumulator bits, 23 of which are summed
void euclid_into(const Builder& B, Word& acc, LweSample* const* a, LweSample* const* b, int slots, int bitsize) {
    if (recorder_mode() >= 3 && 2 * bitsize <= 23) return euclid_columns(B, acc, a, b, slots, bitsize);
    const int max_bits = 24;
    for (int i = 0; i < slots; ++i) {
        const Word d = B.sub(B.load(b[i], bitsize), B.load(a[i], bitsize), bitsize);
        const Word sq = B.multiply(d, d, bitsize, 23);  // only d[0..bitsize-1] enter the product
        Word sq24 = sq;
        sq24.push_back(Bit::constant(false));  // tmp_square[23] is never written in the reference
        Bit carry;
        const Word s = B.add(sq24, acc, max_bits - 1, carry);
        for (int j = 0; j < max_bits - 1; ++j) acc[j] = s[j];
    }
}

}  // namespace

// ================================================================================================ Math.h API
void printSlot64(uint64_t n) {
    printf("[%lu %ld ", (unsigned long)n, (long)n);
    for (int i = 0; i < 64; ++i) putchar('0' + (int)((n >> i) & 1));
    printf("]");
    fflush(stdout);
}

void bootsADD1bit(LweSample* result, LweSample* a, LweSample* b, LweSample* carry, const TFheGateBootstrappingCloudKeySet* bk) {
    RoutineScope scope(bk);
    Builder B{scope.e};
    Bit c = Bit::of(B.e, carry);
    const Bit s = B.full_adder(Bit::of(B.e, a), Bit::of(B.e, b), c);
    s.store(result);
    c.store(carry);
}
void peba1_math_add1bit(LweSample* r, LweSample* a, LweSample* b, LweSample* c, const TFheGateBootstrappingCloudKeySet* bk) { bootsADD1bit(r, a, b, c, bk); }

void bootsADDNbit(LweSample* result, LweSample* a, LweSample* b, LweSample* carry, const int bitsize, const TFheGateBootstrappingCloudKeySet* bk) {
    RoutineScope scope(bk);
    Builder B{scope.e};
    Bit c;
    const Word r = B.add(B.load(a, bitsize), B.load(b, bitsize), bitsize, c);
    store_word(result, r, bitsize);
    c.store(&carry[0]);
}
void peba1_math_add(LweSample* r, LweSample* a, LweSample* b, LweSample* c, int bits, const TFheGateBootstrappingCloudKeySet* bk) { bootsADDNbit(r, a, b, c, bits, bk); }

void bootsTwoSComplement(LweSample* result, LweSample* a, const int bitsize, const TFheGateBootstrappingCloudKeySet* bk) {
    RoutineScope scope(bk);
    Builder B{scope.e};
    store_word(result, B.twos_complement(B.load(a, bitsize), bitsize), bitsize);
}
void peba1_math_twos_complement(LweSample* r, LweSample* a, int bits, const TFheGateBootstrappingCloudKeySet* bk) { bootsTwoSComplement(r, a, bits, bk); }

void bootsABS(LweSample* result, LweSample* a, const int bitsize, const TFheGateBootstrappingCloudKeySet* bk) {
    RoutineScope scope(bk);
    Builder B{scope.e};
    store_word(result, B.abs(B.load(a, bitsize), bitsize), bitsize);
}
void peba1_math_abs(LweSample* r, LweSample* a, int bits, const TFheGateBootstrappingCloudKeySet* bk) { bootsABS(r, a, bits, bk); }

void bootsSUBNbit(LweSample* result, LweSample* a, LweSample* b, const int bitsize, const TFheGateBootstrappingCloudKeySet* bk) {
    RoutineScope scope(bk);
    Builder B{scope.e};
    store_word(result, B.sub(B.load(a, bitsize), B.load(b, bitsize), bitsize), bitsize + 1);
}
void peba1_math_sub(LweSample* r, LweSample* a, LweSample* b, int bits, const TFheGateBootstrappingCloudKeySet* bk) { bootsSUBNbit(r, a, b, bits, bk); }

void bootsShiftLeft(LweSample* result, LweSample* a, const int bitsize, const int n, const TFheGateBootstrappingCloudKeySet* bk) {
    RoutineScope scope(bk);
    Builder B{scope.e};
    const Word w = B.load(a, bitsize > n ? bitsize - n : 0);  // read everything first: result may be a itself
    for (int i = 0; i < bitsize; ++i) (i < n ? Bit::constant(false) : w[i - n]).store(&result[i]);
}
void bootsShiftRight(LweSample* result, LweSample* a, const int bitsize, const int n, const TFheGateBootstrappingCloudKeySet* bk) {
    RoutineScope scope(bk);
    Builder B{scope.e};
    const Word w = B.load(a, bitsize);
    for (int i = 0; i < bitsize; ++i) (i + n < bitsize ? w[i + n] : Bit::constant(false)).store(&result[i]);
}
void bootsShiftLeftNR(LweSample* a, const int bitsize, const int n, const TFheGateBootstrappingCloudKeySet* bk) {
    bootsShiftLeft(a, a, bitsize, n, bk);
}
void peba1_math_shift_left(LweSample* r, LweSample* a, int bits, int n, const TFheGateBootstrappingCloudKeySet* bk) { bootsShiftLeft(r, a, bits, n, bk); }
void peba1_math_shift_right(LweSample* r, LweSample* a, int bits, int n, const TFheGateBootstrappingCloudKeySet* bk) { bootsShiftRight(r, a, bits, n, bk); }
void peba1_math_shift_left_nr(LweSample* a, int bits, int n, const TFheGateBootstrappingCloudKeySet* bk) { bootsShiftLeftNR(a, bits, n, bk); }

void peba1_math_multiply(LweSample* result, LweSample* a, LweSample* b, int bitsize, int width, const TFheGateBootstrappingCloudKeySet* bk) {
    RoutineScope scope(bk);
    Builder B{scope.e};
    const int w = multiply_width(bitsize, width);
    store_word(result, B.multiply(B.load(a, bitsize), B.load(b, bitsize), bitsize, w), w);
}
void bootsMultiply(LweSample* result, LweSample* a, LweSample* b, const int bitsize, const TFheGateBootstrappingCloudKeySet* bk) {
    peba1_math_multiply(result, a, b, bitsize, 0, bk);
}

double approxEquals(Torus32 a, Torus32 b) { return labs((long)a - (long)b) < 10; }

void compare_bit(LweSample* result, const LweSample* a, const LweSample* b, const LweSample* lsb_carry, LweSample* tmp, const TFheGateBootstrappingCloudKeySet* bk) {
    RoutineScope scope(bk);
    Builder B{scope.e};
    const Bit eq = B.XNOR(Bit::of(B.e, a), Bit::of(B.e, b));
    const Bit r = B.MUX(eq, Bit::of(B.e, lsb_carry), Bit::of(B.e, a));
    eq.store(tmp);
    r.store(result);
}
void peba1_math_compare_bit(LweSample* r, const LweSample* a, const LweSample* b, const LweSample* lsb_carry, LweSample* tmp, const TFheGateBootstrappingCloudKeySet* bk) { compare_bit(r, a, b, lsb_carry, tmp, bk); }

void minimum(LweSample* result, LweSample* bit, const LweSample* a, const LweSample* b, const int nb_bits, const TFheGateBootstrappingCloudKeySet* bk) {
    RoutineScope scope(bk);
    Builder B{scope.e};
    Word mn;
    const Bit c = B.minimum(B.load(a, nb_bits), B.load(b, nb_bits), nb_bits, mn);
    store_word(result, mn, nb_bits);
    c.store(&bit[0]);
    for (int i = 1; i < nb_bits; ++i) Bit::constant(false).store(&bit[i]);
}
void peba1_math_minimum(LweSample* r, LweSample* bit, const LweSample* a, const LweSample* b, int n, const TFheGateBootstrappingCloudKeySet* bk) { minimum(r, bit, a, b, n, bk); }

void peba1_math_manhattan(LweSample* result, LweSample* const* a, LweSample* const* b, int slots, int bitsize, const TFheGateBootstrappingCloudKeySet* bk) {
    RoutineScope scope(bk);
    Builder B{scope.e};
    // reference HE_ManhattanDistance (Math.cpp:294-327): sum of |b_i - a_i| kept on `bitsize` bits (wraps)
    Word acc = B.load(result, bitsize + 1);
    for (int i = 0; i < slots; ++i) {
        const Word d = B.sub(B.load(b[i], bitsize), B.load(a[i], bitsize), bitsize);
        Bit carry;
        const Word s = B.add(d, acc, bitsize, carry);
        for (int j = 0; j < bitsize; ++j) acc[j] = s[j];
    }
    store_word(result, acc, bitsize);
}
void HE_ManhattanDistance(LweSample* result, std::vector<LweSample*> a, std::vector<LweSample*> b, const int bitsize, const TFheGateBootstrappingCloudKeySet* bk) {
    check_vectors("One of the ciphertext is empty in a Manhattan distance\n", a.size(), b.size());
    peba1_math_manhattan(result, a.data(), b.data(), (int)std::min(a.size(), b.size()), bitsize, bk);
}

void peba1_math_euclidean(LweSample* result, LweSample* const* a, LweSample* const* b, int slots, int bitsize, const TFheGateBootstrappingCloudKeySet* bk) {
    RoutineScope scope(bk);
    Builder B{scope.e};
    // reference HE_EuclideanDistance (Math.cpp:333-369): the caller's `result` is the running sum (read first)
    check_distance_bitsize("HE_EuclideanDistance", bitsize);
    Word acc = B.load(result, 24);
    euclid_into(B, acc, a, b, slots, bitsize);
    store_word(result, acc, 23);  // bit 23 is never written (Q5)
}
void HE_EuclideanDistance(LweSample* result, std::vector<LweSample*> a, std::vector<LweSample*> b, const int bitsize, const TFheGateBootstrappingCloudKeySet* bk) {
    check_vectors("One of the ciphertext is empty in a Euclidean distance\n", a.size(), b.size());
    peba1_math_euclidean(result, a.data(), b.data(), (int)std::min(a.size(), b.size()), bitsize, bk);
}

void peba1_math_function_f(LweSample* result_b, LweSample* const* a, LweSample* const* b, int slots, LweSample* bound_match, int bitsize, const TFheGateBootstrappingCloudKeySet* bk) {
    RoutineScope scope(bk);
    Builder B{scope.e};
    // reference Function_f (Math.cpp:379-387): the distance accumulator is a fresh allocation, defined here as
    // trivial false (Q2); comparator bit = 1 iff distance > bound (Q3)
    check_distance_bitsize("Function_f", bitsize);
    const int nb = bitsize * 3;
    Word ed((size_t)std::max(nb, 24), Bit::constant(false));
    euclid_into(B, ed, a, b, slots, bitsize);
    Word mn;
    const Bit c = B.minimum(ed, B.load(bound_match, nb), nb, mn);
    c.store(&result_b[0]);
    for (int i = 1; i < nb; ++i) Bit::constant(false).store(&result_b[i]);
}
void Function_f(LweSample* result_b, std::vector<LweSample*> a, std::vector<LweSample*> b, LweSample* bound_match, const int bitsize, const TFheGateBootstrappingCloudKeySet* bk) {
    check_vectors("One of the ciphertext is empty in Function_f\n", a.size(), b.size());
    peba1_math_function_f(result_b, a.data(), b.data(), (int)std::min(a.size(), b.size()), bound_match, bitsize, bk);
}

void Function_g(LweSample* result, LweSample* result_b, LweSample* r0, LweSample* r1, const int bitsize, const TFheGateBootstrappingCloudKeySet* bk) {
    RoutineScope scope(bk);
    Builder B{scope.e};
    // reference Function_g (Math.cpp:390-417): (1 - b) * r0 + b * r1 with SUB / Multiply / ADD; the reference's
    // bitsize-entry buffer for 1 - b is one short (Q4) — a Word has no such limit
    Word one;
    for (int i = 0; i < bitsize; ++i) one.push_back(Bit::constant(i == 0));
    const Word sel = B.load(result_b, bitsize);
    const Word not_b = B.sub(one, sel, bitsize);
    const Word p0 = B.multiply(not_b, B.load(r0, bitsize), bitsize, 23);
    const Word p1 = B.multiply(sel, B.load(r1, bitsize), bitsize, 23);
    Bit carry;
    const Word low0(p0.begin(), p0.begin() + bitsize);
    store_word(result, B.add(low0, p1, bitsize, carry), bitsize);
}
void peba1_math_function_g(LweSample* r, LweSample* rb, LweSample* r0, LweSample* r1, int bits, const TFheGateBootstrappingCloudKeySet* bk) { Function_g(r, rb, r0, r1, bits, bk); }

// ================================================================================================ plaintext mirror
// Same results as the reference's uint64_t routines (Math.cpp:427-587) on every input; written as closed forms where
// one exists.  Not accelerated: main.cpp uses them to check the encrypted results.
uint64_t ADDNbit(uint64_t a, uint64_t b, const int bitsize) {
    // carry-save iteration run bitsize+1 times: exact a+b whenever no carry chain is longer than that
    for (int round = bitsize; round >= 0; --round) {
        const uint64_t generated = a & b;
        a ^= b;
        b = generated << 1;
    }
    return a;
}
uint64_t TwoSComplement(uint64_t a, const int bitsize) {
    const uint64_t low = bitsize >= 64 ? ~UINT64_C(0) : ((UINT64_C(1) << bitsize) - 1);
    return ADDNbit(a ^ low, 1, bitsize);
}
uint64_t ABS(uint64_t a, const int bitsize) {
    const uint64_t top = a >> (bitsize - 1);
    uint64_t mask = 0;
    for (int i = 0; i < bitsize; ++i) mask ^= top << i;
    return ADDNbit(a, mask, bitsize) ^ mask;
}
uint64_t SUBNbit(uint64_t a, uint64_t b, const int bitsize) { return ABS(ADDNbit(a, TwoSComplement(b, bitsize), bitsize), bitsize); }
uint64_t Multiply(uint64_t a, uint64_t b, const int bitsize) {
    const uint64_t low = bitsize >= 64 ? ~UINT64_C(0) : ((UINT64_C(1) << bitsize) - 1);
    uint64_t total = 0;
    for (int i = 0; i < bitsize; ++i) {
        const uint64_t row = ((b >> i) & 1) ? (a & low) : 0;
        total = ADDNbit(row << i, total, 64);
    }
    return total;
}
template <class T>
static uint64_t l1_distance(const std::vector<T>& a, const std::vector<T>& b) {
    if (a.size() != b.size()) perror("Manhattan Distance between two vectors of different sizes.\n");
    long r = 0;
    for (size_t i = 0; i < a.size(); ++i) r += labs((long)a[i] - (long)b[i]);
    return (uint64_t)r;
}
template <class T>
static uint64_t l2sq_distance(const std::vector<T>& a, const std::vector<T>& b) {
    if (a.size() != b.size()) perror("Euclidean Distance between two vectors of different sizes.\n");
    long r = 0;
    for (size_t i = 0; i < a.size(); ++i) r += (long)powl((long double)((long)a[i] - (long)b[i]), 2);
    return (uint64_t)r;
}
uint64_t ManhattanDistance(std::vector<uint8_t> a, std::vector<uint8_t> b) { return l1_distance(a, b); }
uint64_t ManhattanDistance64(std::vector<uint64_t> a, std::vector<uint64_t> b) { return l1_distance(a, b); }
uint64_t EuclideanDistance(std::vector<uint8_t> a, std::vector<uint8_t> b) { return l2sq_distance(a, b); }
uint64_t EuclideanDistance64(std::vector<uint64_t> a, std::vector<uint64_t> b) { return l2sq_distance(a, b); }
uint64_t ManhattanDistanceBitwise(std::vector<uint64_t> a, std::vector<uint64_t> b, const int bitsize) {
    uint64_t r = 0;
    for (size_t i = 0; i < a.size(); ++i) r = ADDNbit(ABS(SUBNbit(b[i], a[i], bitsize), bitsize), r, bitsize);
    return r;
}
uint64_t EuclideanDistanceBitwise(std::vector<uint64_t> a, std::vector<uint64_t> b, const int bitsize) {
    uint64_t r = 0;
    for (size_t i = 0; i < a.size(); ++i) {
        const uint64_t d = SUBNbit(b[i], a[i], bitsize);
        r = ADDNbit(Multiply(d, d, 8), r, bitsize);
    }
    return r;
}
uint64_t Function_f_clear(std::vector<uint64_t> a, std::vector<uint64_t> b, uint64_t bound_match_clear, const int bitsize) {
    return EuclideanDistanceBitwise(a, b, bitsize) <= bound_match_clear ? 1 : 0;
}
uint64_t Function_g_clear(uint64_t result_b, uint64_t r0, uint64_t r1, const int) { return (1 - result_b) * r0 + result_b * r1; }
