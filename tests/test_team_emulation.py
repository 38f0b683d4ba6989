"""CPU test of the very code the blind-rotation kernel runs: csrc/negacyclic_fft.cuh + blind_rotate.cuh are
__host__ __device__, and csrc/emu_team.cu executes one 64-thread team with CPU threads.  It must reproduce the
exact-integer oracle word for word (this container has no GPU; the -m gpu tests repeat this on the device)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "peba1_b200", "csrc")
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "gates_seed42.npz")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", CSRC, "libpeba1_emu.so"])
    return C.CDLL(os.path.join(CSRC, "libpeba1_emu.so"))


def test_team_bootstrap_matches_exact_oracle(emu, oracle, keys):
    p = keys.params
    # transform only the key rows is not possible (all n iterations are needed): full key transform, ~10 s
    bkfft = np.zeros((p.n, 6, 8, 2, 64, 2), dtype=np.float64)
    assert emu.peba1_emu_key_transform(keys.bk.ctypes.data_as(C.c_void_p), p.n, 6,
                                       bkfft.ctypes.data_as(C.c_void_p)) == 0
    g = np.load(GOLDEN)
    tin = np.ascontiguousarray(g["pbs_in"][None, :])
    u = np.zeros((1, 1025), dtype=np.int32)
    assert emu.peba1_emu_bootstrap_woks(3, 7, p.n, bkfft.ctypes.data_as(C.c_void_p),
                                        tin.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p),
                                        C.c_int64(1)) == 0
    assert np.array_equal(u[0], g["pbs_out"])
    # the tensor-memory variant (blind_rotate_tm.cuh) on the same input
    u2 = np.zeros((1, 1025), dtype=np.int32)
    assert emu.peba1_emu_bootstrap_woks_tm(3, 7, p.n, bkfft.ctypes.data_as(C.c_void_p),
                                           tin.ctypes.data_as(C.c_void_p), u2.ctypes.data_as(C.c_void_p),
                                           C.c_int64(1)) == 0
    assert np.array_equal(u2[0], g["pbs_out"])
