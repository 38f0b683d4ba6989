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


def test_pair_variant_matches_exact_oracle(emu, oracle, keys):
    """Two bootstraps per team (blind_rotate_pair.cuh): an odd batch (the last team carries a shadow job), and one input
    with zero mask words so that one ciphertext of a pair has a void step (abar == 0) while the other does not."""
    from oracle import pyoracle
    p = keys.params
    bkfft = np.zeros((p.n, 6, 8, 2, 64, 2), dtype=np.float64)
    assert emu.peba1_emu_key_transform(keys.bk.ctypes.data_as(C.c_void_p), p.n, 6,
                                       bkfft.ctypes.data_as(C.c_void_p)) == 0
    g = np.load(GOLDEN)
    rng = np.random.default_rng(7)
    tin = np.stack([g["pbs_in"], g["pbs_in"], g["pbs_in"]]).astype(np.int32)
    tin[1, :8] = 0                                # void steps for ciphertext 1 only
    tin[1, 8:16] = rng.integers(-2**19, 2**19, 8)  # rounds to abar = 0 as well
    tin[2] = np.roll(tin[2], 5)
    tin = np.ascontiguousarray(tin)
    u = np.zeros((3, 1025), dtype=np.int32)
    assert emu.peba1_emu_bootstrap_woks_pair(3, 7, p.n, bkfft.ctypes.data_as(C.c_void_p),
                                             tin.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p),
                                             C.c_int64(3)) == 0
    assert np.array_equal(u[0], g["pbs_out"])
    for k in (1, 2):
        assert np.array_equal(u[k], keys.bootstrap_woks(tin[k], mode=pyoracle.EXACT)), k
    # key rows through tensor memory (blind_rotate_tmk.cuh): chunk order and void steps (consumer side; the TMA /
    # tcgen05.cp producer only exists on the device and is covered by the -m gpu variant test)
    u4 = np.zeros((3, 1025), dtype=np.int32)
    assert emu.peba1_emu_bootstrap_woks_tmk(3, 7, p.n, bkfft.ctypes.data_as(C.c_void_p),
                                            tin.ctypes.data_as(C.c_void_p), u4.ctypes.data_as(C.c_void_p),
                                            C.c_int64(3)) == 0
    assert np.array_equal(u4, u)
    # one exchange buffer for both exchanges (TeamSmemTm1, the 10-teams-per-SM layout)
    u5 = np.zeros((3, 1025), dtype=np.int32)
    assert emu.peba1_emu_bootstrap_woks_tm1(3, 7, p.n, bkfft.ctypes.data_as(C.c_void_p),
                                            tin.ctypes.data_as(C.c_void_p), u5.ctypes.data_as(C.c_void_p),
                                            C.c_int64(3)) == 0
    assert np.array_equal(u5, u)
    # two alternating exchange buffers, no "buffer free" barrier (TeamSmemTm2, the default kernel's layout): here the
    # emulation is a real check of the hazard argument, the only synchronisation left is the transforms' own barriers
    u6 = np.zeros((3, 1025), dtype=np.int32)
    assert emu.peba1_emu_bootstrap_woks_tm2(3, 7, p.n, bkfft.ctypes.data_as(C.c_void_p),
                                            tin.ctypes.data_as(C.c_void_p), u6.ctypes.data_as(C.c_void_p),
                                            C.c_int64(3)) == 0
    assert np.array_equal(u6, u)
