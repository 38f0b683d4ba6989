"""The oracle and the golden vectors against an INDEPENDENT numpy restatement of the spec (oracle/np_spec.py, SURVEY.md
§8c).  The reference holds no golden vectors for this path and its TFHE dependency is absent, so this is the strongest
pin available here: two implementations that share no code (C with Karatsuba-free schoolbook / FFT forms vs numpy
convolutions with explicit rotation loops) must agree word for word, on the pieces and on whole gates."""
import os

import numpy as np
import pytest

from oracle import np_spec as S

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "gates_seed42.npz")


def test_pieces_against_the_c_oracle(oracle, keys):
    p = keys.params
    rng = np.random.default_rng(12)
    # mod-switch and prologue
    ca, cb = keys.encrypt([1], 5)[0], keys.encrypt([0], 6)[0]
    for op in ("NAND", "XOR", "ORNY"):
        t = oracle.gate_prologue(p, op, ca, cb)
        assert np.array_equal(S.as_i32(S.prologue(op, ca, cb)), t)
    words = rng.integers(-2**31, 2**31, size=64)
    want = [oracle.lib().oracle_mod_switch_from_torus32(int(w), 2048) % 2048 for w in words]
    assert S.mod_switch(words).tolist() == want
    # CMux steps on a random accumulator: rotations below and above N, and the void step
    acc = rng.integers(-2**31, 2**31, size=(2, 1024)).astype(np.int32)
    for i, abar in ((0, 1), (7, 1023), (300, 1024), (629, 2047), (5, 0)):
        got = S.cmux([acc[0], acc[1]], keys.bk[i], abar, p.l, p.Bgbit)
        want = keys.cmux_step(i, abar, acc, mode=oracle.EXACT)
        assert np.array_equal(S.as_i32(np.stack(got)), want), (i, abar)
    # keyswitch of a random extracted sample
    u = rng.integers(-2**31, 2**31, size=1025).astype(np.int32)
    assert np.array_equal(S.as_i32(S.keyswitch(u, keys.ksk)), keys.keyswitch(u))


def test_golden_bootstrap_and_keyswitch_recomputed_independently(keys):
    """the committed pbs_in -> pbs_out -> ks_out chain (one full n = 630 bootstrap), word for word (~15 s)"""
    g = np.load(GOLDEN)
    u = S.bootstrap_woks(g["pbs_in"], keys.bk, keys.params.l, keys.params.Bgbit)
    assert np.array_equal(S.as_i32(u), g["pbs_out"])
    assert np.array_equal(S.as_i32(S.keyswitch(u, keys.ksk)), g["ks_out"])
    # pbs_in itself is the NAND prologue of the golden pair 3
    assert np.array_equal(S.as_i32(S.prologue("NAND", g["ca"][3], g["cb"][3])), g["pbs_in"])


def test_keyswitch_through_row_sums_returns_the_same_words(keys):
    """the table of key-row sums the CUDA keyswitch reads (one entry per group of 4 or 2 digits) is the per-digit
    keyswitch in another order of additions: same words on the golden extracted sample"""
    g = np.load(GOLDEN)
    u = S.u32(g["pbs_out"])
    for group in (4, 2):
        assert np.array_equal(S.as_i32(S.keyswitch_row_sums(u, keys.ksk, group)), g["ks_out"]), group


def test_golden_mux_gate_recomputed_independently(keys):
    """one whole bootsMUX (two bootstraps, sum, keyswitch) of the golden set (~30 s)"""
    g = np.load(GOLDEN)
    k = 5
    out = S.mux(g["cs"][k], g["cx"][k], g["cy"][k], keys.bk, keys.ksk, keys.params.l, keys.params.Bgbit)
    assert np.array_equal(S.as_i32(out), g["out_MUX"][k])
    assert int(keys.decrypt(S.as_i32(out)[None, :])[0]) == int(g["x_bits"][k] if g["s_bits"][k] else g["y_bits"][k])
