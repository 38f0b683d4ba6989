"""ctypes view of the CPU oracle (oracle/libtfhe_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the product package.
See oracle/tfhe_oracle.h for provenance and parity status.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtfhe_oracle.so")

OP = {"NAND": 0, "AND": 1, "OR": 2, "XOR": 3, "XNOR": 4, "NOR": 5,
      "ANDNY": 6, "ANDYN": 7, "ORNY": 8, "ORYN": 9, "MUX": 10}
EXACT, FFT = 0, 1


class Params(C.Structure):
    _fields_ = [("n", C.c_int32), ("N", C.c_int32), ("k", C.c_int32), ("l", C.c_int32),
                ("Bgbit", C.c_int32), ("ks_t", C.c_int32), ("ks_basebit", C.c_int32),
                ("ks_stdev", C.c_double), ("bk_stdev", C.c_double)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tfhe_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        i32p = C.POINTER(C.c_int32)
        L.oracle_default_params.argtypes = [C.c_int32, C.POINTER(Params)]
        L.oracle_keygen.restype = C.c_void_p
        L.oracle_keygen.argtypes = [C.POINTER(Params), C.c_uint64]
        L.oracle_keys_from_arrays.restype = C.c_void_p
        L.oracle_keys_from_arrays.argtypes = [C.POINTER(Params), i32p, i32p, i32p, i32p]
        L.oracle_keys_free.argtypes = [C.c_void_p]
        for name in ("oracle_lwe_key", "oracle_tlwe_key", "oracle_bk", "oracle_ksk"):
            getattr(L, name).restype = i32p
            getattr(L, name).argtypes = [C.c_void_p]
        L.oracle_bk_words.restype = C.c_size_t
        L.oracle_bk_words.argtypes = [C.POINTER(Params)]
        L.oracle_ksk_words.restype = C.c_size_t
        L.oracle_ksk_words.argtypes = [C.POINTER(Params)]
        L.oracle_mod_switch_to_torus32.restype = C.c_int32
        L.oracle_mod_switch_to_torus32.argtypes = [C.c_int32, C.c_int32]
        L.oracle_mod_switch_from_torus32.restype = C.c_int32
        L.oracle_mod_switch_from_torus32.argtypes = [C.c_int32, C.c_int32]
        L.oracle_encrypt_bit.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_uint64), i32p]
        L.oracle_decrypt_bit.restype = C.c_int32
        L.oracle_decrypt_bit.argtypes = [C.c_void_p, i32p]
        L.oracle_phase.restype = C.c_int32
        L.oracle_phase.argtypes = [C.c_void_p, i32p]
        L.oracle_trivial.argtypes = [C.POINTER(Params), C.c_int32, i32p]
        L.oracle_not.argtypes = [C.POINTER(Params), i32p, i32p]
        L.oracle_gate_prologue.restype = C.c_int
        L.oracle_gate_prologue.argtypes = [C.POINTER(Params), C.c_int, i32p, i32p, i32p]
        L.oracle_bootstrap_woks.argtypes = [C.c_void_p, C.c_int, i32p, i32p]
        L.oracle_keyswitch.argtypes = [C.c_void_p, i32p, i32p]
        L.oracle_gate.restype = C.c_int
        L.oracle_gate.argtypes = [C.c_void_p, C.c_int, C.c_int, i32p, i32p, i32p, i32p]
        L.oracle_gate_batch.restype = C.c_int
        L.oracle_gate_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, i32p, i32p, i32p, i32p, C.c_int64, C.c_int]
        L.oracle_cmux_step.argtypes = [C.c_void_p, C.c_int, C.c_int32, C.c_int32, i32p]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def default_params(lam: int = 128) -> Params:
    p = Params()
    if lib().oracle_default_params(lam, C.byref(p)) != 0:
        raise ValueError("minimum_lambda > 128 is not supported (upstream aborts)")
    return p


class Keys:
    """Owning handle on an oracle key set; arrays are numpy views into the C buffers."""

    def __init__(self, handle: int, params: Params):
        self.h = handle
        self.params = params
        L = lib()
        n, N = params.n, params.N
        rows = (params.k + 1) * params.l
        base = 1 << params.ks_basebit
        self.lwe_key = np.ctypeslib.as_array(L.oracle_lwe_key(handle), shape=(n,))
        self.tlwe_key = np.ctypeslib.as_array(L.oracle_tlwe_key(handle), shape=(N,))
        self.bk = np.ctypeslib.as_array(L.oracle_bk(handle), shape=(n, rows, 2, N))
        self.ksk = np.ctypeslib.as_array(L.oracle_ksk(handle), shape=(N, params.ks_t, base, n + 1))

    @classmethod
    def generate(cls, params: Params, seed: int) -> "Keys":
        return cls(lib().oracle_keygen(C.byref(params), C.c_uint64(seed)), params)

    @classmethod
    def from_arrays(cls, params: Params, lwe_key, tlwe_key, bk, ksk) -> "Keys":
        lk = np.ascontiguousarray(lwe_key, dtype=np.int32)
        tk = np.ascontiguousarray(tlwe_key, dtype=np.int32)
        bk = np.ascontiguousarray(bk, dtype=np.int32)
        ksk = np.ascontiguousarray(ksk, dtype=np.int32)
        return cls(lib().oracle_keys_from_arrays(C.byref(params), _p(lk), _p(tk), _p(bk.reshape(-1)),
                                                 _p(ksk.reshape(-1))), params)

    def close(self):
        if self.h:
            lib().oracle_keys_free(self.h)
            self.h = None

    # ---- samples
    @property
    def words(self) -> int:
        return self.params.n + 1

    def encrypt(self, bits, seed: int) -> np.ndarray:
        bits = np.asarray(bits, dtype=np.int64).reshape(-1)
        out = np.empty((bits.size, self.words), dtype=np.int32)
        st = C.c_uint64(seed)
        L = lib()
        for i, b in enumerate(bits):
            L.oracle_encrypt_bit(self.h, int(b), C.byref(st), _p(out[i]))
        return out

    def decrypt(self, cts) -> np.ndarray:
        cts = np.ascontiguousarray(cts, dtype=np.int32).reshape(-1, self.words)
        L = lib()
        return np.array([L.oracle_decrypt_bit(self.h, _p(cts[i])) for i in range(cts.shape[0])], dtype=np.int64)

    def phase(self, cts) -> np.ndarray:
        cts = np.ascontiguousarray(cts, dtype=np.int32).reshape(-1, self.words)
        L = lib()
        return np.array([L.oracle_phase(self.h, _p(cts[i])) for i in range(cts.shape[0])], dtype=np.int64)

    def trivial(self, bit: int) -> np.ndarray:
        out = np.empty(self.words, dtype=np.int32)
        lib().oracle_trivial(C.byref(self.params), int(bit), _p(out))
        return out

    # ---- gates
    def gate(self, op: str, ca, cb, cc=None, mode: int = FFT) -> np.ndarray:
        out = np.empty(self.words, dtype=np.int32)
        ca = np.ascontiguousarray(ca, dtype=np.int32)
        cb = np.ascontiguousarray(cb, dtype=np.int32)
        ccp = _p(np.ascontiguousarray(cc, dtype=np.int32)) if cc is not None else None
        rc = lib().oracle_gate(self.h, mode, OP[op], _p(ca), _p(cb), ccp, _p(out))
        if rc:
            raise ValueError(f"bad op {op}")
        return out

    def gate_batch(self, op: str, ca, cb, cc=None, mode: int = FFT, threads: int = 1) -> np.ndarray:
        ca = np.ascontiguousarray(ca, dtype=np.int32).reshape(-1, self.words)
        cb = np.ascontiguousarray(cb, dtype=np.int32).reshape(-1, self.words)
        out = np.empty_like(ca)
        ccp = None
        if cc is not None:
            cc = np.ascontiguousarray(cc, dtype=np.int32).reshape(-1, self.words)
            ccp = _p(cc.reshape(-1))
        rc = lib().oracle_gate_batch(self.h, mode, OP[op], _p(ca.reshape(-1)), _p(cb.reshape(-1)), ccp,
                                     _p(out.reshape(-1)), ca.shape[0], threads)
        if rc:
            raise ValueError(f"bad op {op}")
        return out

    def bootstrap_woks(self, t, mode: int = FFT) -> np.ndarray:
        out = np.empty(self.params.N + 1, dtype=np.int32)
        lib().oracle_bootstrap_woks(self.h, mode, _p(np.ascontiguousarray(t, dtype=np.int32)), _p(out))
        return out

    def keyswitch(self, u) -> np.ndarray:
        out = np.empty(self.words, dtype=np.int32)
        lib().oracle_keyswitch(self.h, _p(np.ascontiguousarray(u, dtype=np.int32)), _p(out))
        return out

    def cmux_step(self, i: int, abar: int, acc, mode: int = FFT) -> np.ndarray:
        acc = np.array(acc, dtype=np.int32).reshape(2, self.params.N).copy()
        lib().oracle_cmux_step(self.h, mode, i, abar, _p(acc.reshape(-1)))
        return acc


def gate_prologue(params: Params, op: str, ca, cb) -> np.ndarray:
    out = np.empty(params.n + 1, dtype=np.int32)
    rc = lib().oracle_gate_prologue(C.byref(params), OP[op], _p(np.ascontiguousarray(ca, dtype=np.int32)),
                                    _p(np.ascontiguousarray(cb, dtype=np.int32)), _p(out))
    if rc:
        raise ValueError(op)
    return out
