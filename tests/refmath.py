"""The reference's own circuits as the circuit-level oracle — TEST INFRASTRUCTURE.

`oracle/_ref/libpeba1_refmath_cpu.so` is /root/reference/src/Math.cpp compiled UNCHANGED from where it lies (recipe:
`make -C oracle ref`), linked with oracle/tfhe_api_cpu.cpp (the TFHE gate API evaluated on the CPU oracle) and
oracle/refmath_shim.cpp (C-ABI views that only forward).  This module gives it and the product
(`libpeba1_tfhe.so`, include/peba1_math.h) ONE calling convention, so a test runs the same case on both sides and
compares raw ciphertext words, gate counts and level counts.

`CASES` is the list of routine invocations (one per Math.h routine, Math.cpp:27,54,71,97,123,182,192,201,214,259,265,
294,333,379,390) that tests/golden/make_refmath_golden.py ran through the reference to produce
tests/golden/refmath_seed42.npz; the tests replay them.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os

import numpy as np

from peba1_b200 import tfhe_host as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libpeba1_refmath_cpu.so")
GOLDEN = os.path.join(ROOT, "tests", "golden", "refmath_seed42.npz")
GOLDEN_FULL = os.path.join(ROOT, "tests", "golden", "refmath_function_f_full_seed42.npz")

LS = C.POINTER(H.LweSample)
PLS = C.POINTER(LS)
CK = C.POINTER(H.CloudKey)
i32p = C.POINTER(C.c_int32)


class RefStats(C.Structure):
    _fields_ = [("gates", C.c_uint64), ("pbs", C.c_uint64), ("linear", C.c_uint64), ("levels", C.c_uint64),
                ("flushes", C.c_uint64), ("max_level_width", C.c_uint64)]


def ref_available() -> bool:
    return os.path.exists(REF_LIB)


_ref = None


def load_ref():
    global _ref
    if _ref is not None:
        return _ref
    if not ref_available():
        raise ImportError(f"{REF_LIB} is missing: `make -C oracle ref` builds it where /root/reference exists")
    L = C.CDLL(REF_LIB)  # RTLD_LOCAL; the library was linked -Bsymbolic, its boots* are its own
    PS, SK = C.POINTER(H.ParamSet), C.POINTER(H.SecretKey)
    L.new_default_gate_bootstrapping_parameters.restype = PS
    L.new_default_gate_bootstrapping_parameters.argtypes = [C.c_int32]
    L.refapi_keyset_from_oracle.restype = SK
    L.refapi_keyset_from_oracle.argtypes = [PS, C.c_void_p]
    L.delete_gate_bootstrapping_secret_keyset.argtypes = [SK]
    L.new_gate_bootstrapping_ciphertext_array.restype = LS
    L.new_gate_bootstrapping_ciphertext_array.argtypes = [C.c_int32, PS]
    L.delete_gate_bootstrapping_ciphertext_array.argtypes = [C.c_int32, LS]
    L.refapi_sample_at.restype = LS
    L.refapi_sample_at.argtypes = [LS, C.c_int32]
    L.refapi_sample_set.argtypes = [LS, i32p, CK]
    L.refapi_sample_get.argtypes = [LS, CK, i32p]
    L.refapi_get_stats.argtypes = [C.POINTER(RefStats)]
    L.refapi_set_threads.argtypes = [C.c_int]
    L.refapi_set_eager.argtypes = [C.c_int]
    L.refapi_last_level_widths.restype = C.c_int64
    L.refapi_last_level_widths.argtypes = [C.POINTER(C.c_uint64), C.c_int64]
    L.refmath_add1bit.argtypes = [LS, LS, LS, LS, CK]
    L.refmath_add.argtypes = [LS, LS, LS, LS, C.c_int, CK]
    L.refmath_twos_complement.argtypes = [LS, LS, C.c_int, CK]
    L.refmath_abs.argtypes = [LS, LS, C.c_int, CK]
    L.refmath_sub.argtypes = [LS, LS, LS, C.c_int, CK]
    L.refmath_shift_left.argtypes = [LS, LS, C.c_int, C.c_int, CK]
    L.refmath_shift_right.argtypes = [LS, LS, C.c_int, C.c_int, CK]
    L.refmath_shift_left_nr.argtypes = [LS, C.c_int, C.c_int, CK]
    L.refmath_multiply.argtypes = [LS, LS, LS, C.c_int, CK]
    L.refmath_compare_bit.argtypes = [LS, LS, LS, LS, LS, CK]
    L.refmath_minimum.argtypes = [LS, LS, LS, LS, C.c_int, CK]
    L.refmath_manhattan.argtypes = [LS, PLS, PLS, C.c_int, C.c_int, CK]
    L.refmath_euclidean.argtypes = [LS, PLS, PLS, C.c_int, C.c_int, CK]
    L.refmath_function_f.argtypes = [LS, PLS, PLS, C.c_int, LS, C.c_int, CK]
    L.refmath_function_g.argtypes = [LS, LS, LS, LS, C.c_int, CK]
    u64 = C.c_uint64
    u64p = C.POINTER(C.c_uint64)
    for nm, args in (("add", [u64, u64, C.c_int]), ("twos", [u64, C.c_int]), ("abs", [u64, C.c_int]),
                     ("sub", [u64, u64, C.c_int]), ("multiply", [u64, u64, C.c_int]),
                     ("manhattan_bitwise", [u64p, u64p, C.c_int, C.c_int]),
                     ("euclidean_bitwise", [u64p, u64p, C.c_int, C.c_int]),
                     ("function_f", [u64p, u64p, C.c_int, u64, C.c_int]), ("function_g", [u64, u64, u64, C.c_int])):
        f = getattr(L, "refmath_clear_" + nm)
        f.restype, f.argtypes = u64, args
    _ref = L
    return L


def _p(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(i32p)


class RefSide:
    """The reference's Math.cpp on the CPU oracle, keyed by an `oracle.pyoracle.Keys`."""
    name = "reference Math.cpp on the CPU oracle"

    def __init__(self, oracle_keys, threads: int = 0):
        self.L = load_ref()
        self.K = oracle_keys
        self.words = oracle_keys.words
        self.params_p = self.L.new_default_gate_bootstrapping_parameters(128)
        self.sk = self.L.refapi_keyset_from_oracle(self.params_p, oracle_keys.h)
        self.cloud = C.pointer(self.sk.contents.cloud)
        self.L.refapi_set_threads(threads)

    def fn(self, routine):
        return getattr(self.L, "refmath_" + routine)

    def multiply(self, r, a, b, bits):
        self.L.refmath_multiply(r, a, b, bits, self.cloud)

    def new_array(self, n):
        return self.L.new_gate_bootstrapping_ciphertext_array(n, self.params_p)

    def free_array(self, n, arr):
        self.L.delete_gate_bootstrapping_ciphertext_array(n, arr)

    def at(self, arr, i):
        return self.L.refapi_sample_at(arr, i)

    def set_words(self, arr, words):
        words = np.ascontiguousarray(words, dtype=np.int32).reshape(-1, self.words)
        for i in range(words.shape[0]):
            self.L.refapi_sample_set(self.at(arr, i), _p(words[i]), self.cloud)

    def get_words(self, arr, count):
        out = np.empty((count, self.words), dtype=np.int32)
        for i in range(count):
            self.L.refapi_sample_get(self.at(arr, i), self.cloud, _p(out[i]))
        return out

    def begin(self):
        pass

    def end(self):
        self.L.refapi_flush()

    def counters(self):
        s = RefStats()
        self.L.refapi_get_stats(C.byref(s))
        return np.array([s.gates, s.pbs, s.levels], dtype=np.int64)

    def last_level_widths(self):
        buf = (C.c_uint64 * 4096)()
        n = self.L.refapi_last_level_widths(buf, 4096)
        return [int(buf[i]) for i in range(1, min(n, 4096))]


class ProductSide:
    """The product (libpeba1_tfhe.so): on the GPU, or on whatever test backend is installed behind `keyset`."""
    name = "peba1-b200"

    def __init__(self, keyset: H.KeySet):
        self.L = H.load_library()
        self.ks = keyset
        self.words = keyset.words
        self.cloud = keyset.cloud
        self.L.peba1_math_add1bit.argtypes = [LS, LS, LS, LS, CK]
        self.L.peba1_math_compare_bit.argtypes = [LS, LS, LS, LS, LS, CK]
        self.L.peba1_math_shift_left.argtypes = [LS, LS, C.c_int, C.c_int, CK]
        self.L.peba1_math_shift_right.argtypes = [LS, LS, C.c_int, C.c_int, CK]
        self.L.peba1_math_shift_left_nr.argtypes = [LS, C.c_int, C.c_int, CK]

    def fn(self, routine):
        return getattr(self.L, "peba1_math_" + routine)

    def multiply(self, r, a, b, bits):
        self.L.peba1_math_multiply(r, a, b, bits, 0, self.cloud)   # width 0 = the reference's 23

    def new_array(self, n):
        return self.ks.new_array(n)

    def free_array(self, n, arr):
        self.ks.free_array(n, arr)

    def at(self, arr, i):
        return self.ks.at(arr, i)

    def set_words(self, arr, words):
        self.ks.set_words(arr, words)

    def get_words(self, arr, count):
        return self.ks.get_words(arr, count)

    def begin(self):
        self.L.peba1_batch_begin()

    def end(self):
        self.L.peba1_batch_end()

    def counters(self):
        s = self.ks.stats()
        return np.array([s.gates, s.pbs, s.levels], dtype=np.int64)


# ------------------------------------------------------------------------------------------------------------------
# The cases.  Every case is a function(side, enc) -> (list of output word arrays); `enc(value, bits, tag)` returns the
# input ciphertext words [bits][n+1] of an integer (LSB first), the same on every side.

def _int_in(side, enc, value, bits, tag, alloc=None):
    arr = side.new_array(alloc or bits)
    side.set_words(arr, enc(value, bits, tag))
    return arr


def case_add1bit(side, enc):
    """Math.cpp:27 — the public entry, carry in and carry out, all 8 input triples."""
    outs = []
    for t in range(8):
        a, b, c = _int_in(side, enc, t & 1, 1, f"a{t}"), _int_in(side, enc, (t >> 1) & 1, 1, f"b{t}"), \
            _int_in(side, enc, (t >> 2) & 1, 1, f"c{t}")
        r = side.new_array(1)
        side.fn("add1bit")(r, a, b, c, side.cloud)
        outs += [side.get_words(r, 1), side.get_words(c, 1)]
        for x in (a, b, c, r):
            side.free_array(1, x)
    return outs


def case_add8(side, enc):
    """Math.cpp:54"""
    a, b = _int_in(side, enc, 0xB7, 8, "a"), _int_in(side, enc, 0x6D, 8, "b")
    r, c = side.new_array(8), side.new_array(1)
    side.fn("add")(r, a, b, c, 8, side.cloud)
    out = [side.get_words(r, 8), side.get_words(c, 1)]
    for x, n in ((a, 8), (b, 8), (r, 8), (c, 1)):
        side.free_array(n, x)
    return out


def case_twos8(side, enc):
    """Math.cpp:71"""
    a, r = _int_in(side, enc, 0x5A, 8, "a"), side.new_array(8)
    side.fn("twos_complement")(r, a, 8, side.cloud)
    out = [side.get_words(r, 8)]
    side.free_array(8, a); side.free_array(8, r)
    return out


def case_abs9(side, enc):
    """Math.cpp:97 — 9 bits, as bootsSUBNbit's callers use it; a negative and a positive operand"""
    outs = []
    for v in (0x1C9, 0x037):
        a, r = _int_in(side, enc, v, 9, f"a{v}"), side.new_array(9)
        side.fn("abs")(r, a, 9, side.cloud)
        outs.append(side.get_words(r, 9))
        side.free_array(9, a); side.free_array(9, r)
    return outs


def case_sub8(side, enc):
    """Math.cpp:123 — a > b, a < b, and b = 0 (quirk Q1: the result is 256 - a)"""
    outs = []
    for x, y in ((200, 77), (77, 200), (77, 0)):
        a, b, r = _int_in(side, enc, x, 8, f"a{x}"), _int_in(side, enc, y, 8, f"b{y}"), side.new_array(9)
        side.fn("sub")(r, a, b, 8, side.cloud)
        outs.append(side.get_words(r, 9))
        for arr, n in ((a, 8), (b, 8), (r, 9)):
            side.free_array(n, arr)
    return outs


def case_shifts(side, enc):
    """Math.cpp:182,192,201 — no bootstraps, only COPY / CONSTANT"""
    a = _int_in(side, enc, 0xC5, 8, "a")
    l, r = side.new_array(8), side.new_array(8)
    side.fn("shift_left")(l, a, 8, 3, side.cloud)
    side.fn("shift_right")(r, a, 8, 2, side.cloud)
    side.fn("shift_left_nr")(a, 8, 1, side.cloud)
    out = [side.get_words(l, 8), side.get_words(r, 8), side.get_words(a, 8)]
    for arr in (a, l, r):
        side.free_array(8, arr)
    return out


def case_multiply8(side, enc):
    """Math.cpp:214 — 8 x 8 bits into the reference's 23-sample product"""
    a, b, r = _int_in(side, enc, 0xB5, 8, "a"), _int_in(side, enc, 0xE9, 8, "b"), side.new_array(23)
    side.multiply(r, a, b, 8)
    out = [side.get_words(r, 23)]
    for arr, n in ((a, 8), (b, 8), (r, 23)):
        side.free_array(n, arr)
    return out


def case_compare_bit(side, enc):
    """Math.cpp:259 — the public entry, with result aliasing lsb_carry as Math.cpp:272 calls it"""
    outs = []
    for t in range(8):
        a, b = _int_in(side, enc, t & 1, 1, f"a{t}"), _int_in(side, enc, (t >> 1) & 1, 1, f"b{t}")
        carry, tmp = _int_in(side, enc, (t >> 2) & 1, 1, f"c{t}"), side.new_array(1)
        side.fn("compare_bit")(carry, a, b, carry, tmp, side.cloud)
        outs += [side.get_words(carry, 1), side.get_words(tmp, 1)]
        for x in (a, b, carry, tmp):
            side.free_array(1, x)
    return outs


def case_minimum8(side, enc):
    """Math.cpp:265"""
    outs = []
    for x, y in ((0x9C, 0x5B), (0x21, 0xE0)):
        a, b = _int_in(side, enc, x, 8, f"a{x}"), _int_in(side, enc, y, 8, f"b{y}")
        m, bit = side.new_array(8), side.new_array(8)
        side.fn("minimum")(m, bit, a, b, 8, side.cloud)
        outs += [side.get_words(m, 8), side.get_words(bit, 8)]
        for arr in (a, b, m, bit):
            side.free_array(8, arr)
    return outs


VEC_A = [0, 201, 17]     # a zero byte in the FIRST vector: it is the subtrahend of bootsSUBNbit(b[i], a[i]) -> quirk Q1
VEC_B = [93, 180, 17]


def _vectors(side, enc, slots):
    A = [_int_in(side, enc, VEC_A[i], 8, f"va{i}") for i in range(slots)]
    B = [_int_in(side, enc, VEC_B[i], 8, f"vb{i}") for i in range(slots)]
    return A, B, (LS * slots)(*A), (LS * slots)(*B)


def case_manhattan3(side, enc):
    """Math.cpp:294 — 3 slots, result read before it is written (quirk Q2), 9 result samples"""
    A, B, pa, pb = _vectors(side, enc, 3)
    r = side.new_array(9)
    side.fn("manhattan")(r, pa, pb, 3, 8, side.cloud)
    out = [side.get_words(r, 9)]
    for arr in A + B:
        side.free_array(8, arr)
    side.free_array(9, r)
    return out


def case_euclidean3(side, enc):
    """Math.cpp:333 — 3 slots incl. a zero byte in the first vector; all 24 result samples (23 carry the value)"""
    A, B, pa, pb = _vectors(side, enc, 3)
    r = side.new_array(24)
    side.fn("euclidean")(r, pa, pb, 3, 8, side.cloud)
    out = [side.get_words(r, 24)]
    for arr in A + B:
        side.free_array(8, arr)
    side.free_array(24, r)
    return out


def case_function_f2(side, enc):
    """Math.cpp:379 — 2 slots, two bounds around the distance"""
    outs = []
    for bound in (0x00FFFF, 0x000100):
        A, B, pa, pb = _vectors(side, enc, 2)
        eb, rb = _int_in(side, enc, bound, 24, f"bound{bound}"), side.new_array(24)
        side.fn("function_f")(rb, pa, pb, 2, eb, 8, side.cloud)
        outs.append(side.get_words(rb, 24))
        for arr in A + B:
            side.free_array(8, arr)
        side.free_array(24, eb); side.free_array(24, rb)
    return outs


def case_function_g8(side, enc):
    """Math.cpp:390 — result_b = 1 (quirks Q1/Q4 shape the low bits); operands allocated with the head-room the
    reference's own caller gives them (main.cpp:534-570 allocates 24 samples)"""
    rb = _int_in(side, enc, 1, 8, "rb", alloc=24)
    r0, r1 = _int_in(side, enc, 0x3C, 8, "r0", alloc=24), _int_in(side, enc, 0xA7, 8, "r1", alloc=24)
    r = side.new_array(24)
    side.fn("function_g")(r, rb, r0, r1, 8, side.cloud)
    out = [side.get_words(r, 8)]
    for arr in (rb, r0, r1, r):
        side.free_array(24, arr)
    return out


CASES = [("add1bit", case_add1bit), ("add8", case_add8), ("twos8", case_twos8), ("abs9", case_abs9),
         ("sub8", case_sub8), ("shifts", case_shifts), ("multiply8", case_multiply8),
         ("compare_bit", case_compare_bit), ("minimum8", case_minimum8), ("manhattan3", case_manhattan3),
         ("euclidean3", case_euclidean3), ("function_f2", case_function_f2), ("function_g8", case_function_g8)]
CASE = dict(CASES)
CHEAP = ("add1bit", "add8", "twos8", "abs9", "shifts", "compare_bit", "minimum8")   # < 400 gates together


def encoder(oracle_keys, case_name):
    """enc(value, bits, tag): deterministic input ciphertexts, a function of (key, case, tag) only."""
    def enc(value, bits, tag):
        seed = int.from_bytes(hashlib.sha256(f"{case_name}/{tag}".encode()).digest()[:8], "little") >> 1
        return oracle_keys.encrypt([(value >> i) & 1 for i in range(bits)], seed)
    return enc


def run_case(side, oracle_keys, case_name):
    """-> (words [total][n+1], counters [gates, pbs, levels] spent by the case)"""
    c0 = side.counters()
    outs = CASE[case_name](side, encoder(oracle_keys, case_name))
    return np.concatenate(outs), side.counters() - c0


def key_fingerprint(oracle_keys) -> str:
    h = hashlib.sha256()
    for arr in (oracle_keys.lwe_key, oracle_keys.tlwe_key, oracle_keys.bk, oracle_keys.ksk):
        h.update(np.ascontiguousarray(arr).tobytes())
    return h.hexdigest()


def bits_value(bits) -> int:
    return int(sum(int(v) << i for i, v in enumerate(bits)))
