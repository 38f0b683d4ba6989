"""ctypes binding of libpeba1_cuda (include/peba1_cuda.h).  No CPU fallback: a missing library or device raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

OP = {"NAND": 0, "AND": 1, "OR": 2, "XOR": 3, "XNOR": 4, "NOR": 5,
      "ANDNY": 6, "ANDYN": 7, "ORNY": 8, "ORYN": 9, "MUX": 10, "XOR3": 11, "MAJ3": 12}

IN_NEG = lambda i: 1 << i            # noqa: E731
IN_CONST = lambda i: 1 << (4 + i)    # noqa: E731

# every symbol include/peba1_cuda.h declares (tests check that the library exports all of them)
DECLARED_SYMBOLS = [
    "peba1_cuda_last_error", "peba1_cuda_device_count", "peba1_cuda_create", "peba1_cuda_destroy",
    "peba1_cuda_set_stream", "peba1_cuda_sync", "peba1_cuda_load_bootstrap_key", "peba1_cuda_load_keyswitch_key",
    "peba1_cuda_key_device_buffers", "peba1_cuda_keys_mark_loaded", "peba1_cuda_pool_reserve",
    "peba1_cuda_pool_info", "peba1_cuda_upload", "peba1_cuda_download", "peba1_cuda_upload_slots",
    "peba1_cuda_download_slots", "peba1_cuda_encrypt_bits", "peba1_cuda_gates", "peba1_cuda_gate_batch",
    "peba1_cuda_gate_batch_host",
    "peba1_cuda_linear", "peba1_cuda_bootstrap_woks_host", "peba1_cuda_keyswitch_host",
    "peba1_cuda_enable_timing", "peba1_cuda_get_stats", "peba1_cuda_reset_stats",
    "peba1_cuda_measure_fp64_peak", "peba1_cuda_set_option", "peba1_cuda_upload_slots_device",
    "peba1_cuda_level_log_begin", "peba1_cuda_level_log_end",
]


class PEBA1Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libpeba1_cuda error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    """peba1_cuda_params — mirrors TFheGateBootstrappingParameterSet (reference Math.cpp:28 `cloud_key->params`)."""
    _fields_ = [("n", C.c_int32), ("N", C.c_int32), ("k", C.c_int32), ("l", C.c_int32),
                ("Bgbit", C.c_int32), ("ks_t", C.c_int32), ("ks_basebit", C.c_int32)]


class Gate(C.Structure):
    _fields_ = [("op", C.c_int32), ("flags", C.c_uint32), ("in_", C.c_int64 * 3), ("out", C.c_int64)]


GATE_DTYPE = np.dtype([("op", "<i4"), ("flags", "<u4"), ("in", "<i8", (3,)), ("out", "<i8")], align=True)
assert GATE_DTYPE.itemsize == C.sizeof(Gate)


class Stats(C.Structure):
    _fields_ = [("launches", C.c_uint64), ("pbs", C.c_uint64), ("keyswitches", C.c_uint64),
                ("last_blind_rotate_ms", C.c_float), ("last_keyswitch_ms", C.c_float),
                ("sum_blind_rotate_ms", C.c_double), ("sum_keyswitch_ms", C.c_double)]


def default_params(minimum_lambda: int = 128) -> Params:
    """The two upstream gate-bootstrapping parameter sets (reference main.cpp:20-21 uses 128)."""
    if minimum_lambda > 128:
        raise ValueError("minimum_lambda > 128 is not supported")
    if minimum_lambda > 80:
        return Params(630, 1024, 1, 3, 7, 8, 2)
    return Params(500, 1024, 1, 2, 10, 8, 2)


def lib_path() -> str:
    # PEBA1_CUDA_LIB: tuning aid (scripts/ablate.sh builds what-if variants of the kernels)
    return os.environ.get("PEBA1_CUDA_LIB") or os.path.join(_HERE, "csrc", "libpeba1_cuda.so")


_lib = None


def load_library():
    """Loads csrc/libpeba1_cuda.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C peba1_b200/csrc`.  There is no CPU fallback.")
    L = C.CDLL(path)
    vp, i32p, i64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    L.peba1_cuda_last_error.restype = C.c_char_p
    L.peba1_cuda_device_count.restype = C.c_int
    L.peba1_cuda_create.argtypes = [C.c_int, C.POINTER(Params), C.POINTER(vp)]
    L.peba1_cuda_destroy.argtypes = [vp]
    L.peba1_cuda_set_stream.argtypes = [vp, vp]
    L.peba1_cuda_sync.argtypes = [vp]
    L.peba1_cuda_load_bootstrap_key.argtypes = [vp, i32p]
    L.peba1_cuda_load_keyswitch_key.argtypes = [vp, i32p]
    L.peba1_cuda_key_device_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64), C.POINTER(vp),
                                                C.POINTER(C.c_uint64)]
    L.peba1_cuda_keys_mark_loaded.argtypes = [vp]
    L.peba1_cuda_pool_reserve.argtypes = [vp, C.c_int64]
    L.peba1_cuda_pool_info.argtypes = [vp, C.POINTER(vp), i64p, i64p]
    L.peba1_cuda_upload.argtypes = [vp, C.c_int64, C.c_int64, i32p]
    L.peba1_cuda_download.argtypes = [vp, C.c_int64, C.c_int64, i32p]
    L.peba1_cuda_upload_slots.argtypes = [vp, i64p, C.c_int64, i32p]
    L.peba1_cuda_download_slots.argtypes = [vp, i64p, C.c_int64, i32p]
    L.peba1_cuda_gates.argtypes = [vp, vp, C.c_int64]
    L.peba1_cuda_encrypt_bits.argtypes = [vp, i32p, i32p, i64p, C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_double]
    L.peba1_cuda_gate_batch.argtypes = [vp, C.c_int32, i64p, i64p, i64p, i64p, C.c_int64]
    L.peba1_cuda_gate_batch_host.argtypes = [vp, C.c_int32, i32p, i32p, i32p, i32p, C.c_int64]
    L.peba1_cuda_linear.argtypes = [vp, i64p, i32p, i32p, i64p, C.c_int64]
    L.peba1_cuda_bootstrap_woks_host.argtypes = [vp, i32p, i32p, C.c_int64]
    L.peba1_cuda_keyswitch_host.argtypes = [vp, i32p, i32p, C.c_int64]
    L.peba1_cuda_enable_timing.argtypes = [vp, C.c_int]
    L.peba1_cuda_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.peba1_cuda_reset_stats.argtypes = [vp]
    L.peba1_cuda_measure_fp64_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.peba1_cuda_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.peba1_cuda_upload_slots_device.argtypes = [vp, i64p, C.c_int64, C.c_int64, vp]
    L.peba1_cuda_level_log_begin.argtypes = [vp]
    L.peba1_cuda_level_log_end.argtypes = [vp, i64p, i64p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int64, i64p]
    _lib = L
    return L


def _i32(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _i64(a: np.ndarray):
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int64))


class Engine:
    """One GPU context (peba1_cuda_ctx): keys + ciphertext pool + stream."""

    def __init__(self, params: Params | None = None, device: int = 0):
        self.L = load_library()
        self.params = params or default_params(128)
        h = C.c_void_p()
        self._chk(self.L.peba1_cuda_create(device, C.byref(self.params), C.byref(h)))
        self.h = h
        self.device = device
        self.words = self.params.n + 1

    @classmethod
    def from_handle(cls, handle: int, params: Params, device: int = 0) -> "Engine":
        """Wraps a context owned by someone else (e.g. the one behind a libpeba1_tfhe keyset); close() is a no-op."""
        self = cls.__new__(cls)
        self.L = load_library()
        self.params = params
        self.h = C.c_void_p(handle)
        self.device = device
        self.words = params.n + 1
        self._borrowed = True
        return self

    def _chk(self, rc: int):
        if rc != 0:
            raise PEBA1Error(rc, (self.L.peba1_cuda_last_error() or b"").decode())

    def close(self):
        if getattr(self, "_borrowed", False):
            self.h = None
            return
        if getattr(self, "h", None):
            self.L.peba1_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- keys
    def load_keys(self, bk: np.ndarray, ksk: np.ndarray):
        bk = np.ascontiguousarray(bk, dtype=np.int32).reshape(-1)
        ksk = np.ascontiguousarray(ksk, dtype=np.int32).reshape(-1)
        p = self.params
        assert bk.size == p.n * 2 * p.l * 2 * p.N, "bootstrapping key shape"
        assert ksk.size == p.N * p.ks_t * (1 << p.ks_basebit) * (p.n + 1), "keyswitch key shape"
        self._chk(self.L.peba1_cuda_load_bootstrap_key(self.h, _i32(bk)))
        self._chk(self.L.peba1_cuda_load_keyswitch_key(self.h, _i32(ksk)))

    def key_device_buffers(self):
        a, b = C.c_void_p(), C.c_void_p()
        na, nb = C.c_uint64(), C.c_uint64()
        self._chk(self.L.peba1_cuda_key_device_buffers(self.h, C.byref(a), C.byref(na), C.byref(b), C.byref(nb)))
        return (a.value, na.value), (b.value, nb.value)

    def keys_mark_loaded(self):
        self._chk(self.L.peba1_cuda_keys_mark_loaded(self.h))

    # ---- pool
    def reserve(self, slots: int):
        self._chk(self.L.peba1_cuda_pool_reserve(self.h, slots))

    def pool_info(self):
        p, s, c = C.c_void_p(), C.c_int64(), C.c_int64()
        self._chk(self.L.peba1_cuda_pool_info(self.h, C.byref(p), C.byref(s), C.byref(c)))
        return p.value, s.value, c.value

    def upload(self, first: int, cts: np.ndarray):
        cts = np.ascontiguousarray(cts, dtype=np.int32).reshape(-1, self.words)
        self._chk(self.L.peba1_cuda_upload(self.h, first, cts.shape[0], _i32(cts.reshape(-1))))

    def download(self, first: int, count: int) -> np.ndarray:
        out = np.empty((count, self.words), dtype=np.int32)
        self._chk(self.L.peba1_cuda_download(self.h, first, count, _i32(out.reshape(-1))))
        return out

    def upload_slots(self, slots, cts: np.ndarray):
        slots = np.ascontiguousarray(slots, dtype=np.int64)
        cts = np.ascontiguousarray(cts, dtype=np.int32).reshape(-1, self.words)
        assert cts.shape[0] == slots.size
        self._chk(self.L.peba1_cuda_upload_slots(self.h, _i64(slots), slots.size, _i32(cts.reshape(-1))))

    def upload_slots_device(self, device_ptr: int, count: int, first_slot: int = 0, slots=None):
        """pool rows <- `count` dense samples of a DEVICE buffer (no host bounce); see peba1_cuda_upload_slots_device."""
        sl = None if slots is None else np.ascontiguousarray(slots, dtype=np.int64)
        self._chk(self.L.peba1_cuda_upload_slots_device(self.h, None if sl is None else _i64(sl), first_slot, count,
                                                        C.c_void_p(device_ptr)))

    def download_slots(self, slots) -> np.ndarray:
        slots = np.ascontiguousarray(slots, dtype=np.int64)
        out = np.empty((slots.size, self.words), dtype=np.int32)
        self._chk(self.L.peba1_cuda_download_slots(self.h, _i64(slots), slots.size, _i32(out.reshape(-1))))
        return out

    def encrypt_bits(self, lwe_key, bits, first_slot: int, seed: int, sigma: float = 2.0 ** -15, slots=None, index0: int = 0):
        """Device-side encryption of `bits` into consecutive slots from `first_slot` (or into `slots`)."""
        key = np.ascontiguousarray(lwe_key, dtype=np.int32)
        bits = np.ascontiguousarray(bits, dtype=np.int32).reshape(-1)
        sl = None if slots is None else np.ascontiguousarray(slots, dtype=np.int64)
        self._chk(self.L.peba1_cuda_encrypt_bits(self.h, _i32(key), _i32(bits), None if sl is None else _i64(sl),
                                                 first_slot, bits.size, index0, seed, sigma))

    # ---- gates
    def gates(self, gates: np.ndarray):
        """gates: structured array of GATE_DTYPE (one circuit level)."""
        gates = np.ascontiguousarray(gates, dtype=GATE_DTYPE)
        self._chk(self.L.peba1_cuda_gates(self.h, gates.ctypes.data_as(C.c_void_p), gates.size))

    def gate_batch(self, op: str, in0, in1, out, in2=None):
        in0 = np.ascontiguousarray(in0, dtype=np.int64)
        in1 = np.ascontiguousarray(in1, dtype=np.int64)
        out = np.ascontiguousarray(out, dtype=np.int64)
        p2 = _i64(np.ascontiguousarray(in2, dtype=np.int64)) if in2 is not None else None
        self._chk(self.L.peba1_cuda_gate_batch(self.h, OP[op], _i64(in0), _i64(in1), p2, _i64(out), out.size))

    def gate_batch_host(self, op: str, ca, cb, cc=None, out=None) -> np.ndarray:
        """peba1_cuda_gate_batch_host: host buffers in, host buffer out (`out` may be a preallocated, e.g. pinned,
        int32 array of the same shape; a fresh pageable one costs page faults on every call)."""
        ca = np.ascontiguousarray(ca, dtype=np.int32).reshape(-1, self.words)
        cb = np.ascontiguousarray(cb, dtype=np.int32).reshape(-1, self.words)
        if out is None:
            out = np.empty_like(ca)
        assert out.dtype == np.int32 and out.flags["C_CONTIGUOUS"] and out.size == ca.size
        pc = None
        if cc is not None:
            cc = np.ascontiguousarray(cc, dtype=np.int32).reshape(-1, self.words)
            pc = _i32(cc.reshape(-1))
        self._chk(self.L.peba1_cuda_gate_batch_host(self.h, OP[op], _i32(ca.reshape(-1)), _i32(cb.reshape(-1)), pc,
                                                    _i32(out.reshape(-1)), ca.shape[0]))
        return out

    def linear(self, in_slots, sign, cst, out_slots):
        in_slots = np.ascontiguousarray(in_slots, dtype=np.int64)
        out_slots = np.ascontiguousarray(out_slots, dtype=np.int64)
        sign = np.ascontiguousarray(sign, dtype=np.int32)
        cst = np.ascontiguousarray(cst, dtype=np.int32)
        self._chk(self.L.peba1_cuda_linear(self.h, _i64(in_slots), _i32(sign), _i32(cst), _i64(out_slots),
                                           out_slots.size))

    def bootstrap_woks_host(self, t: np.ndarray) -> np.ndarray:
        t = np.ascontiguousarray(t, dtype=np.int32).reshape(-1, self.words)
        out = np.empty((t.shape[0], self.params.N + 1), dtype=np.int32)
        self._chk(self.L.peba1_cuda_bootstrap_woks_host(self.h, _i32(t.reshape(-1)), _i32(out.reshape(-1)),
                                                        t.shape[0]))
        return out

    def keyswitch_host(self, u: np.ndarray) -> np.ndarray:
        u = np.ascontiguousarray(u, dtype=np.int32).reshape(-1, self.params.N + 1)
        out = np.empty((u.shape[0], self.words), dtype=np.int32)
        self._chk(self.L.peba1_cuda_keyswitch_host(self.h, _i32(u.reshape(-1)), _i32(out.reshape(-1)), u.shape[0]))
        return out

    # ---- misc
    def sync(self):
        self._chk(self.L.peba1_cuda_sync(self.h))

    def set_stream(self, cuda_stream: int | None):
        self._chk(self.L.peba1_cuda_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def enable_timing(self, on: bool = True):
        self._chk(self.L.peba1_cuda_enable_timing(self.h, int(on)))

    def level_log_begin(self):
        self._chk(self.L.peba1_cuda_level_log_begin(self.h))

    def level_log_end(self, capacity: int = 1 << 16):
        """-> (gates, pbs, blind_rotate_ms, keyswitch_ms) arrays, one entry per launch pair since level_log_begin"""
        g, p = np.zeros(capacity, dtype=np.int64), np.zeros(capacity, dtype=np.int64)
        a, b = np.zeros(capacity, dtype=np.float32), np.zeros(capacity, dtype=np.float32)
        n = C.c_int64()
        f32p = C.POINTER(C.c_float)
        self._chk(self.L.peba1_cuda_level_log_end(self.h, _i64(g), _i64(p), a.ctypes.data_as(f32p), b.ctypes.data_as(f32p),
                                                  capacity, C.byref(n)))
        k = min(int(n.value), capacity)
        return g[:k], p[:k], a[:k], b[:k]

    def stats(self) -> Stats:
        s = Stats()
        self._chk(self.L.peba1_cuda_get_stats(self.h, C.byref(s)))
        return s

    def reset_stats(self):
        self._chk(self.L.peba1_cuda_reset_stats(self.h))

    def measure_fp64_peak(self) -> float:
        v = C.c_double()
        self._chk(self.L.peba1_cuda_measure_fp64_peak(self.h, C.byref(v)))
        return v.value

    def set_option(self, name: str, value: int):
        self._chk(self.L.peba1_cuda_set_option(self.h, name.encode(), value))

    def try_variant(self, variant: int) -> bool:
        """Selects a blind-rotation kernel variant; False when it is not compiled into this build (the measured
        alternatives that lost are only built with `make -C peba1_b200/csrc VARIANTS=all`)."""
        return self.L.peba1_cuda_set_option(self.h, b"kernel_variant", variant) == 0


def make_gates(op, in0, in1, out, in2=None, flags=0) -> np.ndarray:
    """Builds a GATE_DTYPE array; `op` may be a name, an int, or an array of ints."""
    out = np.asarray(out, dtype=np.int64)
    g = np.zeros(out.size, dtype=GATE_DTYPE)
    g["op"] = OP[op] if isinstance(op, str) else op
    g["flags"] = flags
    g["in"][:, 0] = in0
    g["in"][:, 1] = in1
    if in2 is not None:
        g["in"][:, 2] = in2
    g["out"] = out
    return g
