"""CPU tests of the host side: the TFHE-compatible recorder and the Math.h circuit builder of libpeba1_tfhe.so,
executed on the bit-level simulator (tests/sim_backend.py) instead of the GPU.

Checks (a) decrypted results against native integer arithmetic — the comparisons the reference's own self-test makes
(src/main.cpp:320 ADD, :352 two's complement, :383 ABS, :422 SUB, :456 MUL, :578 protocol verdict) including the
reference's quirks, and (b) the schedule: gate counts and level counts per routine equal the
reference's structure (7 gates per full adder etc.; SURVEY.md §8a [MODEL])."""
import ctypes as C

import numpy as np
import pytest

from peba1_b200 import tfhe_host as H
from sim_backend import SimBackend


@pytest.fixture(scope="module")
def env():
    L = H.load_library()
    sim = SimBackend()
    sim.install()
    ks = H.KeySet.generate(128, seed=7)
    sim.keyset = ks
    yield L, ks, sim
    L.peba1_host_set_test_backend(None)


def run(env, fn):
    """resets the schedule, runs fn (which calls Math routines; eager mode flushes at routine exit)"""
    L, ks, sim = env
    sim.reset_schedule()
    s0 = ks.stats()
    fn()
    s1 = ks.stats()
    return dict(gates=s1.gates - s0.gates, pbs=s1.pbs - s0.pbs, levels=len(sim.level_widths),
                widths=list(sim.level_widths))


def test_keygen_and_roundtrip(env):
    L, ks, sim = env
    assert ks.bk.shape == (630, 6, 2, 1024) and ks.ksk.shape == (1024, 8, 4, 631)
    assert not ks.ksk[:, :, 0, :].any()                      # digit-0 rows are noiseless zeros
    bits = np.random.default_rng(1).integers(0, 2, size=64)
    assert np.array_equal(ks.decrypt(ks.encrypt(bits, 5)), bits)
    # deterministic for a given seed
    ks2 = H.KeySet.generate(128, seed=7)
    assert np.array_equal(ks2.bk[:3], ks.bk[:3]) and np.array_equal(ks2.lwe_key, ks.lwe_key)


def test_host_keys_are_valid_tfhe_keys(env, oracle):
    """keys from the product's own host keygen drive the oracle correctly (independent check of §3.3 keygen)."""
    L, ks, sim = env
    K = oracle.Keys.from_arrays(oracle.default_params(128), ks.lwe_key, ks.tlwe_key, ks.bk, ks.ksk)
    a, b = np.array([0, 1, 0, 1]), np.array([0, 0, 1, 1])
    out = K.gate_batch("NAND", ks.encrypt(a, 1), ks.encrypt(b, 2), threads=4)
    assert ks.decrypt(out).tolist() == (1 - (a & b)).tolist()
    ph = K.phase(out).astype(np.float64) / 2.0**32
    assert np.all(np.abs(np.abs(ph) - 0.125) < 2.0**-5)
    K.close()


def test_single_gates_and_linear_ops(env):
    L, ks, sim = env
    for a in (0, 1):
        for b in (0, 1):
            x = ks.encrypt_int(a | (b << 1), 2, seed=11)
            r = ks.new_array(8)
            L.bootsAND(ks.at(r, 0), ks.at(x, 0), ks.at(x, 1), ks.cloud)
            L.bootsXOR(ks.at(r, 1), ks.at(x, 0), ks.at(x, 1), ks.cloud)
            L.bootsNOT(ks.at(r, 2), ks.at(x, 0), ks.cloud)
            L.bootsOR(ks.at(r, 3), ks.at(r, 2), ks.at(x, 1), ks.cloud)          # gate on a negated alias
            L.bootsCONSTANT(ks.at(r, 4), 1, ks.cloud)
            L.bootsMUX(ks.at(r, 5), ks.at(x, 0), ks.at(r, 4), ks.at(x, 1), ks.cloud)  # a ? 1 : b
            L.bootsCOPY(ks.at(r, 6), ks.at(r, 5), ks.cloud)
            L.bootsXNOR(ks.at(r, 6), ks.at(r, 6), ks.at(x, 1), ks.cloud)        # output aliases an input
            got = ks.decrypt(ks.get_words(r, 8)).tolist()
            want = [a & b, a ^ b, 1 - a, (1 - a) | b, 1, 1 if a else b, 1 - ((1 if a else b) ^ b), 0]
            assert got == want, (a, b)
            ks.free_array(8, r)
            ks.free_array(2, x)


@pytest.mark.parametrize("bits,gates,levels", [(8, 56, 24), (16, 112, 48)])
def test_add_counts_and_values(env, bits, gates, levels):
    L, ks, sim = env
    rng = np.random.default_rng(bits)
    for trial in range(4):
        x, y = (int(v) for v in rng.integers(0, 1 << bits, size=2))
        a, b = ks.encrypt_int(x, bits, 100 + trial), ks.encrypt_int(y, bits, 200 + trial)
        r, c = ks.new_array(bits), ks.new_array(1)
        info = run(env, lambda: L.peba1_math_add(r, a, b, c, bits, ks.cloud))
        assert (info["gates"], info["pbs"], info["levels"]) == (gates, gates, levels)
        assert ks.decrypt_int(r, bits) == (x + y) % (1 << bits)          # main.cpp:320
        assert ks.decrypt_int(c, 1) == (x + y) >> bits
        for arr, n in ((a, bits), (b, bits), (r, bits), (c, 1)):
            ks.free_array(n, arr)


def test_twos_complement_and_abs(env):
    L, ks, sim = env
    for x in (0, 1, 77, 128, 200, 255):
        a = ks.encrypt_int(x, 8, 300 + x)
        r = ks.new_array(8)
        info = run(env, lambda: L.peba1_math_twos_complement(r, a, 8, ks.cloud))
        assert (info["gates"], info["levels"]) == (64, 25)
        assert ks.decrypt_int(r, 8) == (-x) % 256                         # main.cpp:352
        info = run(env, lambda: L.peba1_math_abs(r, a, 8, ks.cloud))
        assert (info["gates"], info["levels"]) == (64, 24)
        sx = x - 256 if x >= 128 else x
        assert ks.decrypt_int(r, 8) == abs(sx) % 256                      # main.cpp:383
        ks.free_array(8, a)
        ks.free_array(8, r)


def ref_sub(a, b):
    """|a - b| with the reference's defect Q1: b == 0 gives (256 - a) on 9 bits."""
    return abs(a - b) if b != 0 else (256 - a) % 512


def test_sub_counts_values_and_quirk_q1(env):
    L, ks, sim = env
    for x, y in [(200, 13), (13, 200), (5, 5), (255, 1), (1, 255), (77, 0), (0, 0), (128, 0)]:
        a, b = ks.encrypt_int(x, 8, 400 + x), ks.encrypt_int(y, 8, 500 + y)
        r = ks.new_array(9)
        info = run(env, lambda: L.peba1_math_sub(r, a, b, 8, ks.cloud))
        assert (info["gates"], info["levels"]) == (226, 59)
        assert ks.decrypt_int(r, 9) == ref_sub(x, y), (x, y)              # main.cpp:422 (+ Q1)
        for arr, n in ((a, 8), (b, 8), (r, 9)):
            ks.free_array(n, arr)


def test_multiply_counts_and_values(env):
    L, ks, sim = env
    rng = np.random.default_rng(9)
    for trial in range(3):
        x, y = (int(v) for v in rng.integers(0, 256, size=2))
        a, b = ks.encrypt_int(x, 8, 600 + trial), ks.encrypt_int(y, 8, 700 + trial)
        r = ks.new_array(23)
        info = run(env, lambda: L.peba1_math_multiply(r, a, b, 8, 0, ks.cloud))
        assert (info["gates"], info["levels"]) == (1296, 81)
        assert max(info["widths"]) == 100
        assert ks.decrypt_int(r, 23) == x * y                             # main.cpp:456
        for arr, n in ((a, 8), (b, 8), (r, 23)):
            ks.free_array(n, arr)


def test_multiply_16bit_parametrised_width(env):
    """BASELINE config 3: 16-bit multiplier needs the reference's hard-coded 23 parametrised (3*16-1 = 47)."""
    L, ks, sim = env
    x, y = 54321, 65535
    a, b = ks.encrypt_int(x, 16, 801), ks.encrypt_int(y, 16, 802)
    r = ks.new_array(47)
    info = run(env, lambda: L.peba1_math_multiply(r, a, b, 16, 47, ks.cloud))
    assert (info["gates"], info["levels"]) == (5408, 169)
    assert ks.decrypt_int(r, 47) == x * y
    for arr, n in ((a, 16), (b, 16), (r, 47)):
        ks.free_array(n, arr)


def test_minimum_and_comparator_polarity(env):
    L, ks, sim = env
    for x, y in [(5, 9), (9, 5), (7, 7), (0, 1 << 23), ((1 << 24) - 1, 0)]:
        a, b = ks.encrypt_int(x, 24, 900), ks.encrypt_int(y, 24, 901)
        r, bit = ks.new_array(24), ks.new_array(24)
        info = run(env, lambda: L.peba1_math_minimum(r, bit, a, b, 24, ks.cloud))
        assert (info["gates"], info["pbs"], info["levels"]) == (72, 120, 26)
        assert ks.decrypt_int(r, 24) == min(x, y)
        assert ks.decrypt_int(bit, 24) == (1 if x > y else 0)             # Q3: 1 iff a > b
        for arr in (a, b, r, bit):
            ks.free_array(24, arr)


def _vec(ks, values, seed):
    arrs = [ks.encrypt_int(int(v), 8, seed + i) for i, v in enumerate(values)]
    ptrs = (C.POINTER(H.LweSample) * len(arrs))(*arrs)
    return arrs, ptrs


def test_distances_small(env):
    L, ks, sim = env
    rng = np.random.default_rng(21)
    slots = 4
    t = rng.integers(1, 255, size=slots)          # no zero byte in the subtrahend: Q1 does not fire
    s = rng.integers(0, 256, size=slots)
    ta, tp = _vec(ks, t, 1000)
    sa, sp = _vec(ks, s, 2000)
    r = ks.new_array(24)
    info = run(env, lambda: L.peba1_math_manhattan(r, tp, sp, slots, 8, ks.cloud))
    assert info["gates"] == slots * (226 + 56)
    assert ks.decrypt_int(r, 8) == int(np.abs(s.astype(int) - t.astype(int)).sum()) % 256
    ks.free_array(24, r)
    r = ks.new_array(24)
    info = run(env, lambda: L.peba1_math_euclidean(r, tp, sp, slots, 8, ks.cloud))
    assert info["gates"] == slots * (226 + 1296 + 161)                     # 215,424 / 128 per slot
    assert ks.decrypt_int(r, 24) == int(((s.astype(int) - t.astype(int)) ** 2).sum())
    for arr in ta + sa:
        ks.free_array(8, arr)
    ks.free_array(24, r)


def test_function_f_and_g_protocol(env):
    """The protocol path of main.cpp:513-581 on 4 slots: f gives 1 iff distance > bound (Q3), g selects r1 when b=1
    and (-r0 mod 256) when b=0 (Q1 inside 1-b)."""
    L, ks, sim = env
    t = np.array([10, 20, 30, 40])
    for probe, bound, expect in [(t + 1, 0, 1), (t + 1, 4, 0), (t + 3, 35, 1), (t + 3, 36, 0)]:
        ta, tp = _vec(ks, t, 3000)
        pa, pp = _vec(ks, probe, 4000)
        eb = ks.encrypt_int(bound, 24, 5000)
        rb = ks.new_array(24)
        info = run(env, lambda: L.peba1_math_function_f(rb, pp, tp, 4, eb, 8, ks.cloud))
        assert info["gates"] == 4 * 1683 + 72 and info["pbs"] == 4 * 1683 + 120
        assert ks.decrypt_int(rb, 24) == expect
        r0, r1 = 0x5A, 0xC3
        e0, e1 = ks.encrypt_int(r0, 8, 6000), ks.encrypt_int(r1, 8, 6001)
        y = ks.new_array(9)
        info = run(env, lambda: L.peba1_math_function_g(y, rb, e0, e1, 8, ks.cloud))
        assert info["gates"] == 2874
        assert ks.decrypt_int(y, 8) == (r1 if expect else (-r0) % 256)
        for arr in ta + pa + [e0, e1]:
            ks.free_array(8, arr)
        ks.free_array(24, eb); ks.free_array(24, rb); ks.free_array(9, y)


def test_function_f_full_size_schedule(env):
    """One face match at the reference's size (128 slots x 8 bits): 215,496 gates / 215,544 PBS / 377 levels, first
    level 6,276 PBS wide (SURVEY.md §8a level profile), and live ciphertexts stay in the low thousands."""
    L, ks, sim = env
    rng = np.random.default_rng(33)
    t = rng.integers(0, 255, size=128)
    probe = (t + 1) % 256                          # main.cpp:39 — never zero, so Q1 does not fire
    ta, tp = _vec(ks, t, 7000)
    pa, pp = _vec(ks, probe, 8000)
    eb = ks.encrypt_int(100, 24, 9000)
    rb = ks.new_array(24)
    info = run(env, lambda: L.peba1_math_function_f(rb, pp, tp, 128, eb, 8, ks.cloud))
    assert (info["gates"], info["pbs"], info["levels"]) == (215496, 215544, 377)
    assert sim.level_pbs[0] == 6276 and sim.level_pbs[1] == 2176
    assert ks.decrypt_int(rb, 24) == 1             # distance 128 > 100
    assert ks.stats().peak_slots < 25000   # ~19.6k live ciphertexts = 50 MB of pool per template in flight
    for arr in ta + pa:
        ks.free_array(8, arr)
    ks.free_array(24, eb); ks.free_array(24, rb)


def test_batched_instances_share_levels(env):
    """BASELINE config 3 shape: many independent adder instances recorded in one batch execute in the SAME 24 levels."""
    L, ks, sim = env
    n = 64
    rng = np.random.default_rng(4)
    xs, ys = rng.integers(0, 256, size=n), rng.integers(0, 256, size=n)
    A = [ks.encrypt_int(int(x), 8, 10000 + i) for i, x in enumerate(xs)]
    B = [ks.encrypt_int(int(y), 8, 20000 + i) for i, y in enumerate(ys)]
    R = [ks.new_array(8) for _ in range(n)]
    Cc = [ks.new_array(1) for _ in range(n)]

    def body():
        L.peba1_batch_begin()
        for i in range(n):
            L.peba1_math_add(R[i], A[i], B[i], Cc[i], 8, ks.cloud)
        L.peba1_batch_end()
    info = run(env, body)
    assert info["gates"] == 56 * n and info["levels"] == 24
    assert sum(info["widths"]) == 56 * n
    for i in range(n):
        assert ks.decrypt_int(R[i], 8) == (int(xs[i]) + int(ys[i])) % 256
    for i in range(n):
        ks.free_array(8, A[i]); ks.free_array(8, B[i]); ks.free_array(8, R[i]); ks.free_array(1, Cc[i])


def test_slots_are_recycled(env):
    """after everything above was freed the pool holds (almost) nothing live"""
    L, ks, sim = env
    assert ks.stats().live_slots < 64


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_fold_modes_keep_decrypted_results(env, mode):
    """constant folding + dead-gate elimination (mode 1) and MUX-carry adders (mode 2) change the
    number of bootstraps, never a decrypted result — including the reference's quirks (Q1, Q3)."""
    L, ks, sim = env
    L.peba1_set_mode(mode)
    try:
        rng = np.random.default_rng(40 + mode)
        for x, y in [(200, 13), (13, 200), (77, 0), (255, 255)] + [tuple(int(v) for v in rng.integers(0, 256, 2)) for _ in range(4)]:
            a, b = ks.encrypt_int(x, 8, 50), ks.encrypt_int(y, 8, 51)
            r, c, d, m = ks.new_array(8), ks.new_array(1), ks.new_array(9), ks.new_array(23)
            s0 = ks.stats()
            L.peba1_batch_begin()
            L.peba1_math_add(r, a, b, c, 8, ks.cloud)
            L.peba1_math_sub(d, a, b, 8, ks.cloud)
            L.peba1_math_multiply(m, a, b, 8, 0, ks.cloud)
            L.peba1_batch_end()
            s1 = ks.stats()
            assert ks.decrypt_int(r, 8) == (x + y) % 256 and ks.decrypt_int(c, 1) == (x + y) >> 8
            assert ks.decrypt_int(d, 9) == ref_sub(x, y)
            assert ks.decrypt_int(m, 23) == x * y
            assert s1.executed_pbs - s0.executed_pbs < (56 + 226 + 1296) * 0.62      # at least 38 % fewer bootstraps
            for arr, n in ((a, 8), (b, 8), (r, 8), (c, 1), (d, 9), (m, 23)):
                ks.free_array(n, arr)
        # protocol path on 4 slots
        t = np.array([10, 20, 30, 40])
        for probe, bound, expect in [(t + 1, 0, 1), (t + 1, 4, 0), (t + 3, 35, 1), (t + 3, 36, 0)]:
            ta, tp = _vec(ks, t, 3000)
            pa, pp = _vec(ks, probe, 4000)
            eb = ks.encrypt_int(bound, 24, 5000)
            rb = ks.new_array(24)
            L.peba1_math_function_f(rb, pp, tp, 4, eb, 8, ks.cloud)
            assert ks.decrypt_int(rb, 24) == expect
            e0, e1 = ks.encrypt_int(0x5A, 8, 6000), ks.encrypt_int(0xC3, 8, 6001)
            y = ks.new_array(9)
            L.peba1_math_function_g(y, rb, e0, e1, 8, ks.cloud)
            assert ks.decrypt_int(y, 8) == (0xC3 if expect else (-0x5A) % 256)
            for arr in ta + pa + [e0, e1]:
                ks.free_array(8, arr)
            ks.free_array(24, eb); ks.free_array(24, rb); ks.free_array(9, y)
    finally:
        L.peba1_set_mode(0)


def test_fold_mode_face_match_size(env):
    """bootstraps per full-size face match in each mode (reported in DESIGN.md)"""
    L, ks, sim = env
    rng = np.random.default_rng(33)
    t = rng.integers(0, 255, size=128)
    probe = (t + 1) % 256
    counts = {}
    for mode in (0, 1, 2, 3):
        L.peba1_set_mode(mode)
        ta, tp = _vec(ks, t, 7000)
        pa, pp = _vec(ks, probe, 8000)
        eb = ks.encrypt_int(100, 24, 9000)
        rb = ks.new_array(24)
        s0 = ks.stats()
        sim.reset_schedule()
        L.peba1_math_function_f(rb, pp, tp, 128, eb, 8, ks.cloud)
        s1 = ks.stats()
        counts[mode] = (s1.executed_pbs - s0.executed_pbs, len(sim.level_widths))
        assert ks.decrypt_int(rb, 24) == 1
        for arr in ta + pa:
            ks.free_array(8, arr)
        ks.free_array(24, eb); ks.free_array(24, rb)
    L.peba1_set_mode(0)
    print("PBS and levels per face match by mode:", counts)
    assert counts[0][0] == 215544 and counts[1][0] < 0.6 * counts[0][0] and counts[2][0] < counts[1][0]
    assert counts[3][0] < 0.5 * counts[2][0] and counts[3][1] < counts[2][1]


def test_encrypt_array_on_device_path(env):
    """peba1_encrypt_array_device: the samples become pool-resident wires holding fresh encryptions of
    the given bits; usable as gate operands.  (On the simulator the words are made on the host; the kernel itself is
    checked by the -m gpu test.)"""
    L, ks, sim = env
    bits = np.random.default_rng(5).integers(0, 2, size=24)
    arr = ks.new_array(24)
    s0 = ks.stats()
    ks.encrypt_array_device(arr, bits, seed=99)
    assert ks.stats().device_encryptions - s0.device_encryptions == 24
    assert ks.decrypt(ks.get_words(arr, 24)).tolist() == bits.tolist()
    out = ks.new_array(1)
    L.bootsXOR(out, ks.at(arr, 0), ks.at(arr, 1), ks.cloud)
    assert int(ks.decrypt(ks.get_words(out, 1))[0]) == int(bits[0] ^ bits[1])
    ks.free_array(24, arr)
    ks.free_array(1, out)


def test_gallery_shards_and_device_generated_templates(env):
    """BASELINE configs[4] plumbing: every rank derives its own shard of the synthetic gallery (template t is a function
    of (seed, t) only), encrypts it where it is used, and Function_f on those arrays gives the expected match bits."""
    from peba1_b200 import gallery as G
    L, ks, sim = env
    total = 37
    whole = G.synthetic_gallery_shard(total, 0, total, slots=4)
    parts = [G.synthetic_gallery_shard(total, *G.shard_bounds(total, 3, r), slots=4) for r in range(3)]
    assert np.array_equal(np.concatenate([p[1] for p in parts]), whole[1])
    assert np.array_equal(np.concatenate([p[2] for p in parts]), whole[2])
    assert all(np.array_equal(p[0], whole[0]) for p in parts)
    assert 0 < int((whole[2] == 0).sum()) < total                 # planted near-copies match, random templates do not
    probe, gal, expect = G.synthetic_gallery_shard(total, 0, 3, slots=4)
    tv = G.encrypt_templates_on_device(ks, gal, bits=8, seed=5)
    out = G.match_template_arrays(ks, ks.encrypt(G.bits_of(probe, 8), 1), ks.encrypt(G.bits_of([512], 24), 2), tv,
                                  slots=4, bits=8)
    assert ks.decrypt(out).tolist() == expect.tolist()


def test_batched_authentication_service(env):
    """the server side of the protocol (main.cpp:534-570) for several clients in one batch.  y = r1
    iff the client's probe is FARTHER than the bound from its template (the reference's polarity, Q3), else -r0 mod 256
    (Q1); the batch shares every circuit level: as many launches as ONE client needs."""
    from peba1_b200 import gallery as G
    L, ks, sim = env
    rng = np.random.default_rng(8)
    K, slots, bound = 3, 4, 35
    tmpl = rng.integers(1, 200, size=(K, slots))
    probes = tmpl + np.array([[1], [3], [4]])            # squared distances 4, 36, 64
    r0, r1 = rng.integers(1, 256, size=K), rng.integers(0, 256, size=K)
    pv = [G._host_vector(ks, ks.encrypt(G.bits_of(probes[k], 8), 100 + k), slots, 8) for k in range(K)]
    tv = G.encrypt_templates_on_device(ks, tmpl, bits=8, seed=9)
    r0w = np.stack([ks.encrypt(G.bits_of([r0[k]], 8), 200 + k) for k in range(K)])
    r1w = np.stack([ks.encrypt(G.bits_of([r1[k]], 8), 300 + k) for k in range(K)])
    sim.reset_schedule()
    y = G.authenticate_batch(ks, pv, tv, ks.encrypt(G.bits_of([bound], 24), 1), r0w, r1w, slots=slots, bits=8)
    got = [int(sum(int(b) << i for i, b in enumerate(ks.decrypt(y[k])))) for k in range(K)]
    far = ((probes - tmpl) ** 2).sum(axis=1) > bound
    assert got == [int(r1[k]) if far[k] else int((-r0[k]) % 256) for k in range(K)]
    levels_batch = len(sim.level_widths)
    # one client alone needs the same number of levels
    pv1 = [G._host_vector(ks, ks.encrypt(G.bits_of(probes[0], 8), 100), slots, 8)]
    tv1 = G.encrypt_templates_on_device(ks, tmpl[:1], bits=8, seed=9)
    sim.reset_schedule()
    G.authenticate_batch(ks, pv1, tv1, ks.encrypt(G.bits_of([bound], 24), 1), r0w[:1], r1w[:1], slots=slots, bits=8)
    assert len(sim.level_widths) == levels_batch


def test_mode3_distance_equals_faithful_on_random_vectors(env):
    """Circuit mode 3 (3-input adders, column-compressed squared distance) against the faithful circuit on the
    simulator: random byte vectors with zero bytes sprinkled in (quirk Q1 fires in the subtraction), several slot
    counts, an incoming non-zero accumulator; and every full-adder input combination."""
    L, ks, sim = env
    rng = np.random.default_rng(2027)

    def run_mode(mode, x, y, acc0):
        L.peba1_set_mode(mode)
        xa, xp = _vec(ks, x, 100)
        ya, yp = _vec(ks, y, 200)
        res = ks.encrypt_int(int(acc0), 24, 300)
        s0 = ks.stats()
        L.peba1_math_euclidean(res, xp, yp, len(x), 8, ks.cloud)
        pbs = ks.stats().executed_pbs - s0.executed_pbs
        v = ks.decrypt_int(res, 24)
        for arr in xa + ya:
            ks.free_array(8, arr)
        ks.free_array(24, res)
        return v, pbs
    try:
        for slots in (1, 3, 8):
            for trial in range(4):
                x, y = rng.integers(0, 256, size=slots), rng.integers(0, 256, size=slots)
                x[rng.integers(0, slots)] = 0
                if trial % 2:
                    y[rng.integers(0, slots)] = 0
                acc0 = int(rng.integers(0, 1 << 12)) if trial >= 2 else 0
                v0, p0 = run_mode(0, x, y, acc0)
                v3, p3 = run_mode(3, x, y, acc0)
                assert v3 == v0, (slots, trial, x, y, acc0)
                assert p3 < 0.2 * p0
        L.peba1_set_mode(3)
        for x in range(8):                 # 3-bit adds: every full-adder input combination, carry-in included
            for y in range(8):
                ea, eb = ks.encrypt_int(x, 3, 400), ks.encrypt_int(y, 3, 401)
                r, c = ks.new_array(3), ks.new_array(1)
                L.peba1_math_add(r, ea, eb, c, 3, ks.cloud)
                assert ks.decrypt_int(r, 3) == (x + y) % 8 and ks.decrypt_int(c, 1) == (x + y) >> 3
                for arr, n in ((ea, 3), (eb, 3), (r, 3), (c, 1)):
                    ks.free_array(n, arr)
    finally:
        L.peba1_set_mode(0)


def test_mode3_absdiff_matches_the_reference_subtraction_low_bits(env):
    """Mode 3's squared distance uses its own |a - b| (borrow chain + conditional negate) instead of bootsSUBNbit.  One
    slot of HE_EuclideanDistance is (low 8 bits of SUBNbit(probe, template))^2: checked against the reference formula
    (|x - y| for y != 0, (-x) mod 256 for y = 0: quirk Q1) for every (x, 0), every (0, y) and random pairs."""
    L, ks, sim = env
    rng = np.random.default_rng(99)
    pairs = [(x, 0) for x in range(256)] + [(0, y) for y in range(256)] + \
            [tuple(int(v) for v in rng.integers(0, 256, 2)) for _ in range(600)] + [(255, 255), (255, 1), (1, 255), (128, 128)]
    L.peba1_set_mode(3)
    try:
        for lo in range(0, len(pairs), 64):                      # 64 one-slot distances per batch
            chunk = pairs[lo:lo + 64]
            res, keep = [], []
            L.peba1_batch_begin()
            for x, y in chunk:
                xa, xp = _vec(ks, [y], 100)                      # first vector = subtrahend (Math.cpp:352: SUB(b[i], a[i]))
                ya, yp = _vec(ks, [x], 200)
                r = ks.new_array(24)
                L.peba1_math_euclidean(r, xp, yp, 1, 8, ks.cloud)
                res.append(r); keep += xa + ya
            L.peba1_batch_end()
            for (x, y), r in zip(chunk, res):
                d = abs(x - y) if y else (-x) % 256
                assert ks.decrypt_int(r, 24) == d * d, (x, y)
                ks.free_array(24, r)
            for arr in keep:
                ks.free_array(8, arr)
    finally:
        L.peba1_set_mode(0)


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_random_gate_dags_decrypt_the_same_in_every_mode(env, mode):
    """Random circuits of boots* gates (constants, NOTs, copies, repeated and complementary operands, MUXes, dead gates)
    recorded in a fold mode decrypt to the same bits as in the faithful mode — the recorder's constant folding,
    aliasing and dead-gate elimination are semantics-preserving."""
    L, ks, sim = env
    two = {"AND": lambda a, b: a & b, "OR": lambda a, b: a | b, "XOR": lambda a, b: a ^ b, "XNOR": lambda a, b: 1 - (a ^ b),
           "NAND": lambda a, b: 1 - (a & b), "NOR": lambda a, b: 1 - (a | b), "ANDNY": lambda a, b: (1 - a) & b,
           "ANDYN": lambda a, b: a & (1 - b), "ORNY": lambda a, b: (1 - a) | b, "ORYN": lambda a, b: a | (1 - b)}
    names = sorted(two)
    rng = np.random.default_rng(500 + mode)
    for trial in range(6):
        nin, nnodes = 5, 40
        inputs = rng.integers(0, 2, size=nin)
        prog = []                                                  # (kind, operands...)
        for k in range(nnodes):
            kind = rng.choice(["g2", "g2", "g2", "mux", "not", "const", "copy"])
            ops = [int(v) for v in rng.integers(0, nin + k, size=3)]
            if rng.random() < 0.2:
                ops[1] = ops[0]                                    # repeated operand
            prog.append((kind, names[int(rng.integers(0, len(names)))], ops, int(rng.integers(0, 2))))
        # plaintext evaluation
        val = list(int(v) for v in inputs)
        for kind, name, ops, cst in prog:
            a, b, c = (val[o] for o in ops)
            val.append({"g2": two[name](a, b), "mux": b if a else c, "not": 1 - a, "const": cst, "copy": a}[kind])
        for m in (0, mode):
            L.peba1_set_mode(m)
            arr = ks.new_array(nin + nnodes)
            ks.set_words(arr, ks.encrypt(inputs, 900 + trial))
            at = lambda i: ks.at(arr, i)
            L.peba1_batch_begin()
            for k, (kind, name, ops, cst) in enumerate(prog):
                out = at(nin + k)
                if kind == "g2":
                    getattr(L, "boots" + name)(out, at(ops[0]), at(ops[1]), ks.cloud)
                elif kind == "mux":
                    L.bootsMUX(out, at(ops[0]), at(ops[1]), at(ops[2]), ks.cloud)
                elif kind == "not":
                    L.bootsNOT(out, at(ops[0]), ks.cloud)
                elif kind == "const":
                    L.bootsCONSTANT(out, cst, ks.cloud)
                else:
                    L.bootsCOPY(out, at(ops[0]), ks.cloud)
            L.peba1_batch_end()
            got = ks.decrypt(ks.get_words(arr, nin + nnodes)).tolist()
            assert got == val, (m, trial)
            ks.free_array(nin + nnodes, arr)
    L.peba1_set_mode(0)


def test_out_of_range_sizes_are_refused_not_overrun():
    """The reference overruns its 23-sample product / 24-sample accumulator arrays for larger operands (Math.cpp:216, :337);
    the drop-in aborts with a message instead (upstream's own convention for fatal errors)."""
    import subprocess
    import sys
    code = ("import ctypes as C, sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from peba1_b200 import tfhe_host as H\nfrom sim_backend import SimBackend\n"
            "L = H.load_library(); sim = SimBackend(); sim.install(); ks = H.KeySet.generate(128, seed=7); sim.keyset = ks\n"
            "a, b, r = ks.new_array(16), ks.new_array(16), ks.new_array(48)\n"
            "L.peba1_math_multiply(r, a, b, 13, 0, ks.cloud)\n") % (
        __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))),
        __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode != 0 and "needs 2 * bitsize - 1 samples" in out.stderr


def test_two_key_sets_driven_by_two_threads():
    """One engine per key set, each with its own wire table and lock, handles that name their engine, batch scope per
    thread: two host threads (one GPU each in production; two simulators here) record and flush concurrently."""
    import threading
    L = H.load_library()
    sims, sets = [], []
    for seed in (21, 22):
        sim = SimBackend()
        sim.install()
        ks = H.KeySet.generate(128, seed=seed)
        sim.keyset = ks
        tmp = ks.new_array(2)                       # the engine is created on first use, while this simulator is installed
        L.bootsAND(ks.at(tmp, 0), ks.at(tmp, 0), ks.at(tmp, 1), ks.cloud)
        ks.get_words(tmp, 1)
        ks.free_array(2, tmp)
        sims.append(sim); sets.append(ks)
    L.peba1_host_set_test_backend(None)
    results, errors = {}, []

    def work(k):
        try:
            ks = sets[k]
            rng = np.random.default_rng(100 + k)
            out = []
            for trial in range(6):
                x, y = (int(v) for v in rng.integers(0, 256, size=2))
                a, b = ks.encrypt_int(x, 8, 10 * trial + k), ks.encrypt_int(y, 8, 10 * trial + 5 + k)
                r = ks.new_array(23)
                L.peba1_batch_begin()               # this thread's scope only
                L.peba1_math_multiply(r, a, b, 8, 0, ks.cloud)
                L.peba1_batch_end()
                out.append((x * y, ks.decrypt_int(r, 23)))
                for arr, n in ((a, 8), (b, 8), (r, 23)):
                    ks.free_array(n, arr)
            results[k] = out
        except Exception as e:                      # noqa: BLE001
            errors.append(e)
    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in th: t.start()
    for t in th: t.join()
    assert not errors, errors
    for k in range(2):
        assert all(want == got for want, got in results[k]), results[k]
