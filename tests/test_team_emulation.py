"""CPU test of the very code the blind-rotation kernel runs: csrc/negacyclic_fft.cuh + blind_rotate.cuh are
__host__ __device__, and tests/emu/emu_team.cu executes one 64-thread team with CPU threads.  It must reproduce the
exact-integer oracle word for word (this container has no GPU; the -m gpu tests repeat this on the device)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "gates_seed42.npz")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", EMU, "libpeba1_emu.so"])
    return C.CDLL(os.path.join(EMU, "libpeba1_emu.so"))


@pytest.fixture(scope="module")
def bkfft(emu, keys):
    """the bootstrapping key transformed by the emulated key_transform_team (all n rows are needed): once per module"""
    p = keys.params
    out = np.zeros((p.n, 6, 8, 2, 64, 2), dtype=np.float64)
    assert emu.peba1_emu_key_transform(keys.bk.ctypes.data_as(C.c_void_p), p.n, 6, out.ctypes.data_as(C.c_void_p)) == 0
    return out


def test_team_bootstrap_matches_exact_oracle(emu, oracle, keys, bkfft):
    p = keys.params
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
    # the layout the default wide-batch kernel runs on: two alternating exchange buffers (TeamSmemTm2)
    u3 = np.zeros((1, 1025), dtype=np.int32)
    assert emu.peba1_emu_bootstrap_woks_tm2(3, 7, p.n, bkfft.ctypes.data_as(C.c_void_p),
                                            tin.ctypes.data_as(C.c_void_p), u3.ctypes.data_as(C.c_void_p),
                                            C.c_int64(1)) == 0
    assert np.array_equal(u3[0], g["pbs_out"])
    # the low-latency kernel code (blind_rotate_wide.cuh): 2l teams = 384 CPU threads on the one bootstrap
    u4 = np.zeros((1, 1025), dtype=np.int32)
    assert emu.peba1_emu_bootstrap_woks_wide(3, 7, p.n, bkfft.ctypes.data_as(C.c_void_p),
                                             tin.ctypes.data_as(C.c_void_p), u4.ctypes.data_as(C.c_void_p),
                                             C.c_int64(1)) == 0
    assert np.array_equal(u4[0], g["pbs_out"])


def test_other_variants_match_exact_oracle_on_a_short_key(emu, oracle):
    """The other team functions on a SHORT key (n = 48 CMux steps: the emulation's 64 CPU threads meet at some 30
    barriers per step), against the exact-integer oracle with the same key:
    two bootstraps per team (blind_rotate_pair.cuh) on an odd batch (the last team carries a shadow job) with one input
    that has void steps (abar == 0) while its partner has none; the key-ring consumer (blind_rotate_tmk.cuh); the
    single-exchange-buffer layouts (TeamSmemTm1, TeamSmemTm2)."""
    po = oracle.default_params(128)
    po.n = 48
    K = oracle.Keys.generate(po, 11)
    n = po.n
    bkfft = np.zeros((n, 6, 8, 2, 64, 2), dtype=np.float64)
    assert emu.peba1_emu_key_transform(K.bk.ctypes.data_as(C.c_void_p), n, 6, bkfft.ctypes.data_as(C.c_void_p)) == 0
    rng = np.random.default_rng(7)
    tin = rng.integers(-2**31, 2**31, size=(3, n + 1), dtype=np.int64).astype(np.int32)
    tin[1, :8] = 0                                 # void steps for ciphertext 1 only
    tin[1, 8:16] = rng.integers(-2**19, 2**19, 8)  # rounds to abar = 0 as well
    tin = np.ascontiguousarray(tin)
    want = np.stack([K.bootstrap_woks(t, mode=oracle.EXACT) for t in tin])

    def run(fn, x):
        x = np.ascontiguousarray(x)
        u = np.zeros((x.shape[0], 1025), dtype=np.int32)
        assert fn(3, 7, n, bkfft.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p),
                  C.c_int64(x.shape[0])) == 0
        return u
    try:
        assert np.array_equal(run(emu.peba1_emu_bootstrap_woks_pair, tin), want)
        # consumer side of the key ring (the TMA / tcgen05.cp producer only exists on the device: -m gpu variant tests)
        assert np.array_equal(run(emu.peba1_emu_bootstrap_woks_tmk, tin[1:]), want[1:])
        assert np.array_equal(run(emu.peba1_emu_bootstrap_woks_tm1, tin[1:]), want[1:])
        # two alternating exchange buffers, no "buffer free" barrier: here the emulation really checks the hazard
        # argument, the only synchronisation left is the transforms' own barriers
        assert np.array_equal(run(emu.peba1_emu_bootstrap_woks_tm2, tin[1:]), want[1:])
        assert np.array_equal(run(emu.peba1_emu_bootstrap_woks_tm, tin), want)
        # the inner exchange by warp shuffles (lane8_transpose; kernel_variant 10)
        assert np.array_equal(run(emu.peba1_emu_bootstrap_woks_tm2s, tin[1:]), want[1:])
        # the cross-team exchange through tensor memory (TeamSmemTmX; kernel_variant 11): warps (w, w + 4) as a team, the
        # 16x256b fragment layout, the permuted logical thread positions and the re-swizzled inner exchange
        assert np.array_equal(run(emu.peba1_emu_bootstrap_woks_tmx, tin), want)
        # the low-latency kernel does not skip void steps (all-zero digits instead): ciphertext 1 exercises that, and
        # three jobs in a row the re-use of the key-row pipeline across jobs
        assert np.array_equal(run(emu.peba1_emu_bootstrap_woks_wide, tin), want)
    finally:
        K.close()


def test_second_parameter_set_on_a_short_key(emu, oracle):
    """l = 2, Bgbit = 10 (upstream's minimum_lambda <= 80 set: 4 key rows per CMux step, 10-bit digits) through the
    register kernel code and both tensor-memory layouts, n = 40 CMux steps."""
    po = oracle.default_params(80)
    assert (po.l, po.Bgbit) == (2, 10)
    po.n = 40
    K = oracle.Keys.generate(po, 13)
    n, rows = po.n, 2 * po.l
    bkfft = np.zeros((n, rows, 8, 2, 64, 2), dtype=np.float64)
    assert emu.peba1_emu_key_transform(K.bk.ctypes.data_as(C.c_void_p), n, rows, bkfft.ctypes.data_as(C.c_void_p)) == 0
    rng = np.random.default_rng(17)
    tin = np.ascontiguousarray(rng.integers(-2**31, 2**31, size=(2, n + 1), dtype=np.int64).astype(np.int32))
    try:
        want = np.stack([K.bootstrap_woks(t, mode=oracle.EXACT) for t in tin])
        for fn in (emu.peba1_emu_bootstrap_woks, emu.peba1_emu_bootstrap_woks_tm, emu.peba1_emu_bootstrap_woks_tm2,
                   emu.peba1_emu_bootstrap_woks_tm2s, emu.peba1_emu_bootstrap_woks_tmx, emu.peba1_emu_bootstrap_woks_wide):
            u = np.zeros((2, 1025), dtype=np.int32)
            assert fn(2, 10, n, bkfft.ctypes.data_as(C.c_void_p), tin.ctypes.data_as(C.c_void_p),
                      u.ctypes.data_as(C.c_void_p), C.c_int64(2)) == 0
            assert np.array_equal(u, want)
    finally:
        K.close()
