"""Host randomness (peba1_b200/host/chacha_rng.hpp): ChaCha20 against the RFC 8439 §2.3.2 block, OS-keyed by default,
deterministic only where a seed is named."""
import ctypes as C

import numpy as np

from peba1_b200 import tfhe_host as H


def test_chacha20_block_known_answer():
    L = H.load_library()
    L.peba1_host_chacha_block.argtypes = [C.POINTER(C.c_uint32), C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    key = (C.c_uint32 * 8)(*[int.from_bytes(bytes(range(4 * i, 4 * i + 4)), "little") for i in range(8)])
    out = (C.c_uint64 * 8)()
    # RFC 8439 §2.3.2: block counter 1, nonce 00:00:00:09 00:00:00:4a 00:00:00:00 -> state words 12..15 = 1, 0x09000000,
    # 0x4a000000, 0
    L.peba1_host_chacha_block(key, 1 | (0x09000000 << 32), 0x4A000000, out)
    words = [w for v in out for w in (v & 0xFFFFFFFF, v >> 32)]
    assert words == [0xe4e7f110, 0x15593bd1, 0x1fdd0f50, 0xc47120a3, 0xc7f4d1c7, 0x0368c033, 0x9aaa2204, 0x4e6cd4c3,
                     0x466482d2, 0x09aa9f07, 0x05d7c214, 0xa2028bd9, 0xd19c12b5, 0xb94e16de, 0xe883d0cb, 0x4e3c50a2]


def test_unseeded_keysets_differ_and_seeded_ones_repeat():
    L = H.load_library()
    p = L.new_default_gate_bootstrapping_parameters(128)
    a = H.KeySet(p, L.new_random_gate_bootstrapping_secret_keyset(p))   # main.cpp:22, never seeded by the reference
    b = H.KeySet(p, L.new_random_gate_bootstrapping_secret_keyset(p))
    assert (a.lwe_key != b.lwe_key).sum() > 200 and (a.tlwe_key != b.tlwe_key).sum() > 350
    assert 0.4 < a.lwe_key.mean() < 0.6
    s1, s2 = H.KeySet.generate(128, seed=9), H.KeySet.generate(128, seed=9)
    assert np.array_equal(s1.lwe_key, s2.lwe_key) and np.array_equal(s1.bk[:2], s2.bk[:2])
    bits = np.random.default_rng(0).integers(0, 2, size=32)
    assert np.array_equal(a.decrypt(a.encrypt(bits, 1)), bits)
    assert np.array_equal(a.encrypt(bits, 1), a.encrypt(bits, 1)) and not np.array_equal(a.encrypt(bits, 1), a.encrypt(bits, 2))


def test_setseed_makes_the_process_generator_deterministic():
    L = H.load_library()
    L.tfhe_random_generator_setSeed.argtypes = [C.POINTER(C.c_uint32), C.c_int32]
    p = L.new_default_gate_bootstrapping_parameters(128)
    seed = (C.c_uint32 * 3)(1, 2, 3)
    L.tfhe_random_generator_setSeed(seed, 3)
    a = H.KeySet(p, L.new_random_gate_bootstrapping_secret_keyset(p))
    L.tfhe_random_generator_setSeed(seed, 3)
    b = H.KeySet(p, L.new_random_gate_bootstrapping_secret_keyset(p))
    assert np.array_equal(a.lwe_key, b.lwe_key) and np.array_equal(a.ksk[:4], b.ksk[:4])
