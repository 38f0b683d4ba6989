"""Circuit parity against the REFERENCE ITSELF: /root/reference/src/Math.cpp, compiled unchanged into
oracle/_ref/libpeba1_refmath_cpu.so (`make -C oracle ref`), evaluated gate by gate on the CPU oracle.

tests/golden/refmath_seed42.npz holds what the reference's routines returned for the cases of tests/refmath.py
(raw ciphertext words, gate / bootstrap / level counts, decrypted bits); tests/golden/make_refmath_golden.py made it in
the build container, where /root/reference exists.  Here:

  CPU (-m "not gpu")
    * the prebuilt reference library reproduces the golden words for the cheap cases (so the file is what the
      reference computes, not a stale artefact) and evaluates the same words call by call as deferred;
    * the product's circuit builder (math_circuits.cpp) on the bit-level simulator issues the reference's gate,
      bootstrap and level counts and decrypts to the reference's bits for EVERY case;
    * the product's schedule run on the CPU oracle returns the reference's raw words (cheap cases; all with
      PEBA1_SLOW=1).
  GPU (-m gpu)
    * every Math.h routine through libpeba1_tfhe + libpeba1_cuda returns the reference's raw ciphertext words
      (tolerance: 0 — the gate path is deterministic integer arithmetic), with equal gate, bootstrap
      and level counts: Math.cpp:27 (bootsADD1bit, carry in/out), :54, :71, :97 (bootsABS), :123 (incl. quirk Q1),
      :182-201, :214, :259 (compare_bit with result == lsb_carry), :265, :294 (HE_ManhattanDistance), :333
      (HE_EuclideanDistance with a zero byte, all 24 samples), :379 (Function_f), :390 (Function_g);
    * a live run of the reference library on the box for the cheap cases;
    * the full-size face match (128 slots, 215,424 + 72 gates) against the reference's words
      (tests/golden/refmath_function_f_full_seed42.npz).
"""
import ctypes as C
import os

import numpy as np
import pytest

from peba1_b200 import tfhe_host as H
import refmath as R

SLOW = os.environ.get("PEBA1_SLOW") == "1"


@pytest.fixture(scope="module")
def golden(keys):
    g = np.load(R.GOLDEN)
    assert str(g["key_fingerprint"]) == R.key_fingerprint(keys), \
        "the oracle's keygen(seed 42) no longer produces the keys the golden file was made with"
    return g


def product_keyset(keys):
    """a product keyset around the oracle's key arrays (so every side holds the same keys)"""
    L = H.load_library()
    p = L.new_default_gate_bootstrapping_parameters(128)
    i32p = C.POINTER(C.c_int32)
    arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (keys.lwe_key, keys.tlwe_key, keys.bk, keys.ksk)]
    sk = L.peba1_new_keyset_from_arrays(p, *[a.reshape(-1).ctypes.data_as(i32p) for a in arrs])
    return H.KeySet(p, sk)


# ------------------------------------------------------------------------------------------------------------ CPU
@pytest.mark.skipif(not R.ref_available(), reason="oracle/_ref not built (needs /root/reference; `make -C oracle ref`)")
def test_reference_library_reproduces_the_golden_file(keys, golden):
    side = R.RefSide(keys)
    for name in R.CHEAP:
        words, counters = R.run_case(side, keys, name)
        assert np.array_equal(words, golden[name + "/words"]), name
        assert counters.tolist() == golden[name + "/counters"].tolist(), name


@pytest.mark.skipif(not R.ref_available(), reason="oracle/_ref not built")
def test_reference_library_deferred_equals_call_by_call(keys, golden):
    """the CPU gate API under the reference defers evaluation to use all cores; gate by gate gives the same words"""
    side = R.RefSide(keys)
    side.L.refapi_set_eager(1)
    try:
        for name in ("add1bit", "compare_bit", "minimum8"):
            words, _ = R.run_case(side, keys, name)
            assert np.array_equal(words, golden[name + "/words"]), name
    finally:
        side.L.refapi_set_eager(0)


@pytest.mark.skipif(not R.ref_available(), reason="oracle/_ref not built")
def test_reference_plaintext_mirror_matches_the_decrypted_golden(keys, golden):
    """the reference's own self-test compares decryptions with its plaintext routines (main.cpp:320,352,383,422,456);
    the golden bits satisfy those identities"""
    L = R.load_ref()
    val = R.bits_value
    assert val(golden["add8/bits"][:8]) == L.refmath_clear_add(0xB7, 0x6D, 8) & 0xFF
    assert val(golden["twos8/bits"]) == L.refmath_clear_twos(0x5A, 8) & 0xFF
    assert val(golden["multiply8/bits"]) == L.refmath_clear_multiply(0xB5, 0xE9, 8) == 0xB5 * 0xE9
    sub = golden["sub8/bits"].reshape(3, 9)
    assert [val(s) for s in sub] == [123, 123, 256 - 77]          # |200-77|, |77-200|, quirk Q1 for b = 0
    ed = val(golden["euclidean3/bits"][:23])
    # slot 0 has a zero subtrahend: the reference's circuit squares the LOW 8 BITS of (256 - 93) = 163, not 93
    assert ed == 163 * 163 + 21 * 21 + 0


def test_product_circuits_on_the_simulator_match_the_reference_counts_and_bits(keys, golden):
    """every case: the product's builder issues exactly the reference's number of gates, bootstraps and levels, and the
    decrypted outputs are the reference's"""
    from sim_backend import SimBackend
    L = H.load_library()
    sim = SimBackend()
    sim.install()
    try:
        ks = product_keyset(keys)
        sim.keyset = ks
        side = R.ProductSide(ks)
        for name, _ in R.CASES:
            words, counters = R.run_case(side, keys, name)
            assert counters.tolist() == golden[name + "/counters"].tolist(), name
            assert np.array_equal(ks.decrypt(words), golden[name + "/bits"]), name
    finally:
        L.peba1_host_set_test_backend(None)


def test_product_schedule_on_the_oracle_returns_the_reference_words(keys, golden):
    """the product's recorder + level scheduler, its gates executed by the CPU oracle: raw words == the reference's"""
    from oracle_backend import OracleBackend
    be = OracleBackend(keys)
    be.install()
    try:
        side = R.ProductSide(product_keyset(keys))
        for name in ([n for n, _ in R.CASES] if SLOW else R.CHEAP):
            words, counters = R.run_case(side, keys, name)
            assert np.array_equal(words, golden[name + "/words"]), name
            assert counters.tolist() == golden[name + "/counters"].tolist(), name
    finally:
        OracleBackend.uninstall()


# ------------------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def gpu_side(keys):
    H.load_library().peba1_host_set_test_backend(None)
    ks = product_keyset(keys)
    assert ks.cuda_context(), "no CUDA context: the product has no CPU path"
    return R.ProductSide(ks)


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n, _ in R.CASES])
def test_gpu_returns_the_reference_words(gpu_side, keys, golden, name):
    words, counters = R.run_case(gpu_side, keys, name)
    assert np.array_equal(words, golden[name + "/words"])               # bit-exact, every word of every sample
    assert counters.tolist() == golden[name + "/counters"].tolist()     # gates, bootstraps, levels


@pytest.mark.gpu
@pytest.mark.skipif(not R.ref_available(), reason="oracle/_ref did not travel to this box")
def test_gpu_equals_a_live_run_of_the_reference(gpu_side, keys):
    side = R.RefSide(keys)
    for name in R.CHEAP:
        want, wc = R.run_case(side, keys, name)
        got, gc = R.run_case(gpu_side, keys, name)
        assert np.array_equal(got, want), name
        assert gc.tolist() == wc.tolist(), name


@pytest.mark.gpu
def test_gpu_full_size_face_match_equals_the_reference(gpu_side, keys):
    """128 slots x 8 bits: HE_EuclideanDistance (Math.cpp:333) word for word — all 24 samples, 23-bit value decrypted —
    then Function_f (Math.cpp:379) on the same inputs for a bound at and just below the distance"""
    if not os.path.exists(R.GOLDEN_FULL):
        pytest.fail("tests/golden/refmath_function_f_full_seed42.npz is missing")
    g = np.load(R.GOLDEN_FULL)
    assert str(g["key_fingerprint"]) == R.key_fingerprint(keys)
    side, slots = gpu_side, len(g["template"])
    enc = R.encoder(keys, "full")
    A = [R._int_in(side, enc, int(v), 8, f"t{i}") for i, v in enumerate(g["template"])]
    B = [R._int_in(side, enc, int(v), 8, f"p{i}") for i, v in enumerate(g["probe"])]
    pa, pb = (R.LS * slots)(*A), (R.LS * slots)(*B)
    ed = side.new_array(24)
    c0 = side.counters()
    side.fn("euclidean")(ed, pa, pb, slots, 8, side.cloud)
    got = side.get_words(ed, 24)
    assert (side.counters() - c0).tolist() == g["ed/counters"].tolist()
    assert np.array_equal(got, g["ed/words"])
    assert R.bits_value(keys.decrypt(got)[:23]) == int(g["distance"])
    # the bit minimum() hands Function_f is "ed > bound" (compare_bit keeps a's bit at the highest differing position,
    # Math.cpp:259-261): 0 for the bound AT the distance, 1 for the bound just below it — the reference's own words say so
    for tag, verdict in (("ge", 0), ("lt", 1)):
        eb = R._int_in(side, enc, int(g[f"result_b/{tag}/bound"]), 24, f"bound{int(g[f'result_b/{tag}/bound'])}")
        rb = side.new_array(24)
        side.fn("function_f")(rb, pa, pb, slots, eb, 8, side.cloud)
        words = side.get_words(rb, 24)
        assert np.array_equal(words, g[f"result_b/{tag}/words"])
        assert int(keys.decrypt(words[:1])[0]) == verdict
        side.free_array(24, eb); side.free_array(24, rb)
    for arr in A + B:
        side.free_array(8, arr)
    side.free_array(24, ed)
