"""CPU tests of the oracle (oracle/tfhe_oracle.c) — the checker itself must be pinned before it checks anything.

Pins: gate truth tables (the decrypted-bit identities main.cpp relies on), FFT form == exact
integer form, the committed golden vectors, the mod-switch constants, and the trivial/NOT helpers.
"""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "gates_seed42.npz")

TRUTH = {
    "NAND": lambda a, b: 1 - (a & b), "AND": lambda a, b: a & b, "OR": lambda a, b: a | b,
    "XOR": lambda a, b: a ^ b, "XNOR": lambda a, b: 1 - (a ^ b), "NOR": lambda a, b: 1 - (a | b),
    "ANDNY": lambda a, b: (1 - a) & b, "ANDYN": lambda a, b: a & (1 - b),
    "ORNY": lambda a, b: (1 - a) | b, "ORYN": lambda a, b: a | (1 - b),
}


def test_params(oracle):
    p = oracle.default_params(128)
    assert (p.n, p.N, p.k, p.l, p.Bgbit, p.ks_t, p.ks_basebit) == (630, 1024, 1, 3, 7, 8, 2)
    p80 = oracle.default_params(80)
    assert (p80.n, p80.l, p80.Bgbit) == (500, 2, 10)
    with pytest.raises(ValueError):
        oracle.default_params(129)


def test_mod_switch(oracle):
    L = oracle.lib()
    assert L.oracle_mod_switch_to_torus32(1, 8) == 1 << 29
    assert L.oracle_mod_switch_to_torus32(-1, 8) == -(1 << 29)
    assert L.oracle_mod_switch_to_torus32(2, 8) == 1 << 30
    # round to nearest multiple of 2^-11
    assert L.oracle_mod_switch_from_torus32(0, 2048) == 0
    assert L.oracle_mod_switch_from_torus32((1 << 20) - 1, 2048) == 0
    assert L.oracle_mod_switch_from_torus32(1 << 20, 2048) == 1
    assert L.oracle_mod_switch_from_torus32(-1, 2048) == 0        # wraps to 2048 == 0 (mod 2N)
    assert L.oracle_mod_switch_from_torus32(-(1 << 21), 2048) == 2047


def test_encrypt_decrypt_roundtrip(keys):
    """reference main.cpp:58-94 — the demo's only hard assertion."""
    bits = np.random.default_rng(7).integers(0, 2, size=256)
    cts = keys.encrypt(bits, 99)
    assert np.array_equal(keys.decrypt(cts), bits)
    # phases sit near +-1/8 with sigma = 2^-15
    ph = keys.phase(cts).astype(np.float64) / 2.0**32
    assert np.all(np.abs(np.abs(ph) - 0.125) < 6 * 2.0**-15)


def test_trivial_and_not(oracle, keys):
    t1, t0 = keys.trivial(1), keys.trivial(0)
    assert t1[-1] == 1 << 29 and t0[-1] == -(1 << 29) and not t1[:-1].any()
    assert keys.decrypt(np.stack([t1, t0])).tolist() == [1, 0]
    ct = keys.encrypt([1], 5)[0]
    out = np.empty_like(ct)
    oracle.lib().oracle_not(__import__("ctypes").byref(keys.params), oracle._p(ct), oracle._p(out))
    assert keys.decrypt(out[None])[0] == 0


def test_golden_key_checksums(keys):
    g = np.load(GOLDEN)
    assert int(g["key_seed"]) == 42
    assert np.bitwise_xor.reduce(keys.bk.reshape(-1).view(np.uint32)) == g["bk_xor"]
    assert np.bitwise_xor.reduce(keys.ksk.reshape(-1).view(np.uint32)) == g["ksk_xor"]


def test_fft_form_matches_golden_exact(oracle, keys):
    """FP64-FFT oracle with round-to-nearest == exact-integer oracle, word for word."""
    g = np.load(GOLDEN)
    for op in TRUTH:
        out = keys.gate_batch(op, g["ca"], g["cb"], mode=oracle.FFT, threads=4)
        assert np.array_equal(out, g["out_" + op]), op
        assert keys.decrypt(out).tolist() == [TRUTH[op](int(a), int(b)) for a, b in zip(g["a_bits"], g["b_bits"])]
    out = keys.gate_batch("MUX", g["cs"], g["cx"], g["cy"], mode=oracle.FFT, threads=4)
    assert np.array_equal(out, g["out_MUX"])
    assert keys.decrypt(out).tolist() == [int(x if s else y) for s, x, y in zip(g["s_bits"], g["x_bits"], g["y_bits"])]


def test_pieces_match_golden(oracle, keys):
    g = np.load(GOLDEN)
    u = keys.bootstrap_woks(g["pbs_in"], mode=oracle.FFT)
    assert np.array_equal(u, g["pbs_out"])
    assert np.array_equal(keys.keyswitch(u), g["ks_out"])


def test_exact_one_gate_recomputed(oracle, keys):
    """One exact-integer gate recomputed live (1.3 s) so the golden file cannot drift from the C source."""
    g = np.load(GOLDEN)
    out = keys.gate("XOR", g["ca"][1], g["cb"][1], mode=oracle.EXACT)
    assert np.array_equal(out, g["out_XOR"][1])


def test_cmux_step_fft_vs_exact(oracle, keys):
    rng = np.random.default_rng(3)
    acc = rng.integers(-2**31, 2**31, size=(2, 1024), dtype=np.int64).astype(np.int32)
    for i, abar in [(0, 1), (17, 1023), (629, 1024), (300, 2047)]:
        assert np.array_equal(keys.cmux_step(i, abar, acc, mode=oracle.EXACT),
                              keys.cmux_step(i, abar, acc, mode=oracle.FFT))


def test_chained_gates_on_constants(oracle, keys):
    """bootsCONSTANT inputs still bootstrap correctly (reference Math.cpp:58,84 — quirk Q6)."""
    t0, t1 = keys.trivial(0), keys.trivial(1)
    x = keys.gate("XOR", t1, t0)            # 1
    y = keys.gate("AND", x, t1)             # 1
    z = keys.gate("NAND", y, x)             # 0
    m = keys.gate("MUX", z, t0, y)          # z ? 0 : y = 1
    assert keys.decrypt(np.stack([x, y, z, m])).tolist() == [1, 1, 0, 1]


def test_output_phase_noise_margin(oracle, keys):
    """bootstrapped outputs sit within 2^-5 of +-1/8 (SURVEY.md §8c spec check: max 0.008 observed)."""
    bits_a = np.random.default_rng(11).integers(0, 2, size=16)
    bits_b = np.random.default_rng(12).integers(0, 2, size=16)
    out = keys.gate_batch("NAND", keys.encrypt(bits_a, 21), keys.encrypt(bits_b, 22), threads=4)
    assert keys.decrypt(out).tolist() == [1 - (int(a) & int(b)) for a, b in zip(bits_a, bits_b)]
    ph = keys.phase(out).astype(np.float64) / 2.0**32
    assert np.all(np.abs(np.abs(ph) - 0.125) < 2.0**-5)


def test_lambda80_parameter_set_fft_equals_exact(oracle):
    """upstream's 80-bit set (n=500, l=2, Bgbit=10): larger digits, still exact after rounding on these samples"""
    K = oracle.Keys.generate(oracle.default_params(80), 7)
    a, b = np.array([0, 1, 0, 1]), np.array([0, 0, 1, 1])
    ca, cb = K.encrypt(a, 1), K.encrypt(b, 2)
    f = K.gate_batch("NAND", ca, cb, mode=oracle.FFT, threads=4)
    assert K.decrypt(f).tolist() == (1 - (a & b)).tolist()
    for i in (0, 3):
        assert np.array_equal(K.gate("NAND", ca[i], cb[i], mode=oracle.EXACT), f[i])
    K.close()
