"""ctypes view of libpeba1_tfhe.so — the TFHE-compatible front end (include/tfhe/*.h, include/peba1_host.h) and the
PEBA1 Math.h circuit library (include/peba1_math.h).  Used by tests and bench.py; computes nothing itself."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .cuda import load_library as _load_cuda

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class ParamSet(C.Structure):
    _fields_ = [("ks_t", C.c_int32), ("ks_basebit", C.c_int32), ("n", C.c_int32), ("N", C.c_int32),
                ("k", C.c_int32), ("l", C.c_int32), ("Bgbit", C.c_int32), ("ks_stdev", C.c_double),
                ("bk_stdev", C.c_double), ("max_stdev", C.c_double), ("minimum_lambda", C.c_int32)]


class CloudKey(C.Structure):
    _fields_ = [("params", C.POINTER(ParamSet)), ("material", C.c_void_p)]


class SecretKey(C.Structure):
    _fields_ = [("params", C.POINTER(ParamSet)), ("lwe_key", C.POINTER(C.c_int32)),
                ("tlwe_key", C.POINTER(C.c_int32)), ("cloud", CloudKey)]


class LweSample(C.Structure):
    _fields_ = [("a", C.POINTER(C.c_int32)), ("b", C.c_int32), ("current_variance", C.c_double),
                ("peba1_handle", C.c_uint64), ("peba1_n", C.c_int32), ("peba1_pad", C.c_int32)]


class HostStats(C.Structure):
    _fields_ = [("gates", C.c_uint64), ("pbs", C.c_uint64), ("linear", C.c_uint64), ("levels", C.c_uint64),
                ("flushes", C.c_uint64), ("max_level_width", C.c_uint64), ("uploads", C.c_uint64),
                ("downloads", C.c_uint64), ("live_slots", C.c_uint64), ("peak_slots", C.c_uint64),
                ("flush_seconds", C.c_double), ("executed_gates", C.c_uint64), ("executed_pbs", C.c_uint64),
                ("folded", C.c_uint64), ("dead", C.c_uint64), ("device_encryptions", C.c_uint64)]


class TestBackend(C.Structure):
    RESERVE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64)
    UPLOAD = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_int32))
    DOWNLOAD = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int32))
    GATES = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64)
    _fields_ = [("user", C.c_void_p), ("reserve", RESERVE), ("upload_slots", UPLOAD), ("download", DOWNLOAD),
                ("gates", GATES)]


def lib_path() -> str:
    # PEBA1_TFHE_LIB: an instrumented build of the same sources (scripts/sanitize_host.sh: ASan + UBSan)
    return os.environ.get("PEBA1_TFHE_LIB") or os.path.join(_HERE, "host", "libpeba1_tfhe.so")


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    _load_cuda()  # resolves libpeba1_cuda.so first (same directory tree)
    L = C.CDLL(path)
    PS, SK, CK, LS = C.POINTER(ParamSet), C.POINTER(SecretKey), C.POINTER(CloudKey), C.POINTER(LweSample)
    i32p = C.POINTER(C.c_int32)
    L.new_default_gate_bootstrapping_parameters.restype = PS
    L.new_default_gate_bootstrapping_parameters.argtypes = [C.c_int32]
    L.delete_gate_bootstrapping_parameters.argtypes = [PS]
    L.new_random_gate_bootstrapping_secret_keyset.restype = SK
    L.new_random_gate_bootstrapping_secret_keyset.argtypes = [PS]
    L.peba1_new_secret_keyset_seeded.restype = SK
    L.peba1_new_secret_keyset_seeded.argtypes = [PS, C.c_uint64]
    L.peba1_new_keyset_from_arrays.restype = SK
    L.peba1_new_keyset_from_arrays.argtypes = [PS, i32p, i32p, i32p, i32p]
    L.peba1_new_placeholder_keyset.restype = SK
    L.peba1_new_placeholder_keyset.argtypes = [PS]
    L.delete_gate_bootstrapping_secret_keyset.argtypes = [SK]
    L.new_gate_bootstrapping_ciphertext_array.restype = LS
    L.new_gate_bootstrapping_ciphertext_array.argtypes = [C.c_int32, PS]
    L.delete_gate_bootstrapping_ciphertext_array.argtypes = [C.c_int32, LS]
    L.bootsSymEncrypt.argtypes = [LS, C.c_int32, SK]
    L.bootsSymDecrypt.restype = C.c_int32
    L.bootsSymDecrypt.argtypes = [LS, SK]
    L.bootsCONSTANT.argtypes = [LS, C.c_int32, CK]
    for nm in ("bootsNOT", "bootsCOPY"):
        getattr(L, nm).argtypes = [LS, LS, CK]
    for nm in ("bootsNAND", "bootsOR", "bootsAND", "bootsXOR", "bootsXNOR", "bootsNOR", "bootsANDNY", "bootsANDYN",
               "bootsORNY", "bootsORYN"):
        getattr(L, nm).argtypes = [LS, LS, LS, CK]
    L.bootsMUX.argtypes = [LS, LS, LS, LS, CK]
    L.peba1_host_set_device.argtypes = [C.c_int]
    L.peba1_host_set_test_backend.argtypes = [C.POINTER(TestBackend)]
    L.peba1_set_eager.argtypes = [C.c_int]
    L.peba1_set_mode.argtypes = [C.c_int]
    L.peba1_get_mode.restype = C.c_int
    L.peba1_get_stats.argtypes = [C.POINTER(HostStats)]
    L.peba1_sample_get.argtypes = [LS, CK, i32p]
    L.peba1_sample_set.argtypes = [LS, i32p]
    L.peba1_sample_at.restype = LS
    L.peba1_sample_at.argtypes = [LS, C.c_int32]
    L.peba1_keyset_bk.restype = i32p
    L.peba1_keyset_bk.argtypes = [CK]
    L.peba1_keyset_ksk.restype = i32p
    L.peba1_keyset_ksk.argtypes = [CK]
    L.peba1_encrypt_bits.argtypes = [SK, i32p, C.c_int64, C.c_uint64, i32p]
    L.peba1_decrypt_words.argtypes = [SK, i32p, C.c_int64, i32p]
    L.peba1_encrypt_array_device.argtypes = [LS, C.c_int32, i32p, SK, C.c_uint64]
    L.peba1_encrypt_array_list_device.argtypes = [C.POINTER(LS), C.c_int32, C.c_int32, i32p, SK, C.c_uint64]
    # <tfhe/tfhe_io.h>: FILE* arguments are libc streams (see open_file below)
    FP = C.c_void_p
    L.export_tfheGateBootstrappingParameterSet_toFile.argtypes = [FP, PS]
    L.new_tfheGateBootstrappingParameterSet_fromFile.restype = PS
    L.new_tfheGateBootstrappingParameterSet_fromFile.argtypes = [FP]
    L.export_tfheGateBootstrappingCloudKeySet_toFile.argtypes = [FP, CK]
    L.new_tfheGateBootstrappingCloudKeySet_fromFile.restype = CK
    L.new_tfheGateBootstrappingCloudKeySet_fromFile.argtypes = [FP]
    L.export_tfheGateBootstrappingSecretKeySet_toFile.argtypes = [FP, SK]
    L.new_tfheGateBootstrappingSecretKeySet_fromFile.restype = SK
    L.new_tfheGateBootstrappingSecretKeySet_fromFile.argtypes = [FP]
    L.export_gate_bootstrapping_ciphertext_toFile.argtypes = [FP, LS, PS]
    L.import_gate_bootstrapping_ciphertext_fromFile.argtypes = [FP, LS, PS]
    L.delete_gate_bootstrapping_cloud_keyset.argtypes = [CK]
    L.peba1_sample_get_words.argtypes = [LS, i32p]
    L.peba1_sample_set_list_device.argtypes = [C.POINTER(LS), C.c_int32, C.c_int32, C.c_void_p, CK]
    L.peba1_sample_device_slot.restype = C.c_int64
    L.peba1_sample_device_slot.argtypes = [LS, CK]
    L.peba1_new_cloud_keyset_from_arrays.restype = CK
    L.peba1_new_cloud_keyset_from_arrays.argtypes = [PS, i32p, i32p]
    L.peba1_cuda_context_of.restype = C.c_void_p
    L.peba1_cuda_context_of.argtypes = [CK]
    L.peba1_math_add.argtypes = [LS, LS, LS, LS, C.c_int, CK]
    L.peba1_math_twos_complement.argtypes = [LS, LS, C.c_int, CK]
    L.peba1_math_abs.argtypes = [LS, LS, C.c_int, CK]
    L.peba1_math_sub.argtypes = [LS, LS, LS, C.c_int, CK]
    L.peba1_math_multiply.argtypes = [LS, LS, LS, C.c_int, C.c_int, CK]
    L.peba1_math_minimum.argtypes = [LS, LS, LS, LS, C.c_int, CK]
    PLS = C.POINTER(LS)
    L.peba1_math_manhattan.argtypes = [LS, PLS, PLS, C.c_int, C.c_int, CK]
    L.peba1_math_euclidean.argtypes = [LS, PLS, PLS, C.c_int, C.c_int, CK]
    L.peba1_math_function_f.argtypes = [LS, PLS, PLS, C.c_int, LS, C.c_int, CK]
    L.peba1_math_function_g.argtypes = [LS, LS, LS, LS, C.c_int, CK]
    _lib = L
    return L


def _i32(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class KeySet:
    """A secret keyset (params + LWE/TLWE keys + embedded cloud keyset), generated on the host."""

    def __init__(self, params, sk):
        self.L = load_library()
        self.params_p, self.sk = params, sk
        self.params = params.contents
        self.cloud = C.pointer(sk.contents.cloud)
        p = self.params
        self.words = p.n + 1
        rows = (p.k + 1) * p.l
        bkp, kskp = self.L.peba1_keyset_bk(self.cloud), self.L.peba1_keyset_ksk(self.cloud)
        self.bk = np.ctypeslib.as_array(bkp, shape=(p.n, rows, 2, p.N)) if bkp else None
        self.ksk = np.ctypeslib.as_array(kskp, shape=(p.N, p.ks_t, 1 << p.ks_basebit, p.n + 1)) if kskp else None
        self.lwe_key = np.ctypeslib.as_array(sk.contents.lwe_key, shape=(p.n,))
        self.tlwe_key = np.ctypeslib.as_array(sk.contents.tlwe_key, shape=(p.N,))

    @classmethod
    def generate(cls, minimum_lambda: int = 128, seed: int = 42) -> "KeySet":
        L = load_library()
        params = L.new_default_gate_bootstrapping_parameters(minimum_lambda)
        return cls(params, L.peba1_new_secret_keyset_seeded(params, seed))

    @classmethod
    def placeholder(cls, minimum_lambda: int = 128) -> "KeySet":
        """A replica's keyset: no key material on the host, device keys arrive by NCCL broadcast."""
        L = load_library()
        params = L.new_default_gate_bootstrapping_parameters(minimum_lambda)
        return cls(params, L.peba1_new_placeholder_keyset(params))

    def cuda_context(self) -> int:
        """Handle of the libpeba1_cuda context behind this keyset (created on first use)."""
        return self.L.peba1_cuda_context_of(self.cloud)

    def encrypt(self, bits, seed: int) -> np.ndarray:
        bits = np.ascontiguousarray(bits, dtype=np.int32).reshape(-1)
        out = np.empty((bits.size, self.words), dtype=np.int32)
        self.L.peba1_encrypt_bits(self.sk, _i32(bits), bits.size, seed, _i32(out.reshape(-1)))
        return out

    def decrypt(self, words) -> np.ndarray:
        words = np.ascontiguousarray(words, dtype=np.int32).reshape(-1, self.words)
        out = np.empty(words.shape[0], dtype=np.int32)
        self.L.peba1_decrypt_words(self.sk, _i32(words.reshape(-1)), words.shape[0], _i32(out))
        return out.astype(np.int64)

    # ---- ciphertext arrays (LweSample*) for the circuit API
    def new_array(self, count: int):
        return self.L.new_gate_bootstrapping_ciphertext_array(count, self.params_p)

    def free_array(self, count: int, arr):
        self.L.delete_gate_bootstrapping_ciphertext_array(count, arr)

    def at(self, arr, i: int):
        return self.L.peba1_sample_at(arr, i)

    def set_words(self, arr, words: np.ndarray):
        words = np.ascontiguousarray(words, dtype=np.int32).reshape(-1, self.words)
        for i in range(words.shape[0]):
            self.L.peba1_sample_set(self.at(arr, i), _i32(words[i]))

    def encrypt_array_device(self, arr, bits, seed: int):
        """arr[i] <- fresh encryption of bits[i], generated on the GPU (no host->device ciphertext traffic)."""
        bits = np.ascontiguousarray(bits, dtype=np.int32).reshape(-1)
        self.L.peba1_encrypt_array_device(arr, bits.size, _i32(bits), self.sk, seed)

    def encrypt_array_list_device(self, arrs, nbelems: int, bits, seed: int):
        """arrs[k][i] <- fresh encryption of bits[k*nbelems + i], one launch for the whole list."""
        bits = np.ascontiguousarray(bits, dtype=np.int32).reshape(-1)
        assert bits.size == len(arrs) * nbelems
        ptrs = (C.POINTER(LweSample) * len(arrs))(*arrs)
        self.L.peba1_encrypt_array_list_device(ptrs, len(arrs), nbelems, _i32(bits), self.sk, seed)

    def set_list_device(self, arrs, nbelems: int, device_ptr: int):
        """arrs[k][i] <- sample k*nbelems+i of a dense DEVICE buffer (one scatter on the GPU, no host bounce)."""
        ptrs = (C.POINTER(LweSample) * len(arrs))(*arrs)
        self.L.peba1_sample_set_list_device(ptrs, len(arrs), nbelems, C.c_void_p(device_ptr), self.cloud)

    def get_words(self, arr, count: int) -> np.ndarray:
        out = np.empty((count, self.words), dtype=np.int32)
        for i in range(count):
            self.L.peba1_sample_get(self.at(arr, i), self.cloud, _i32(out[i]))
        return out

    def encrypt_int(self, value: int, bits: int, seed: int):
        """`bits` fresh encryptions of the bits of `value`, LSB first, as a new ciphertext array."""
        arr = self.new_array(bits)
        self.set_words(arr, self.encrypt([(value >> i) & 1 for i in range(bits)], seed))
        return arr

    def decrypt_int(self, arr, bits: int) -> int:
        b = self.decrypt(self.get_words(arr, bits))
        return int(sum(int(v) << i for i, v in enumerate(b)))

    def stats(self) -> HostStats:
        s = HostStats()
        self.L.peba1_get_stats(C.byref(s))
        return s


class open_file:
    """libc FILE* for the <tfhe/tfhe_io.h> functions: `with open_file(path, "wb") as F: L.export_...(F, ...)`."""
    _libc = None

    def __init__(self, path: str, mode: str):
        if open_file._libc is None:
            open_file._libc = C.CDLL(None)
            open_file._libc.fopen.restype = C.c_void_p
            open_file._libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
            open_file._libc.fclose.argtypes = [C.c_void_p]
        self.fp = open_file._libc.fopen(os.fsencode(path), mode.encode())
        if not self.fp:
            raise OSError(f"cannot open {path!r} with mode {mode!r}")

    def __enter__(self):
        return self.fp

    def __exit__(self, *exc):
        open_file._libc.fclose(self.fp)
        return False
