"""Generates tests/golden/refmath_seed42.npz (and, with --full, refmath_function_f_full_seed42.npz): outputs of the
REFERENCE's own circuits — /root/reference/src/Math.cpp compiled unchanged into oracle/_ref/libpeba1_refmath_cpu.so by
`make -C oracle ref` — evaluated gate by gate on the CPU oracle, for the cases listed in tests/refmath.py.

Run in the build container (where /root/reference exists):   python tests/golden/make_refmath_golden.py [--full]
The GPU box has no /root/reference; the tests there replay the cases against these files (and against the prebuilt
oracle/_ref library, which travels with the repository snapshot)."""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle as O   # noqa: E402
import refmath as R                # noqa: E402

KEY_SEED = 42
FULL_SLOTS = 128


def full_vectors():
    """the full-size face match of tests: 128 bytes per vector, zero bytes in the first vector (quirk Q1)"""
    rng = np.random.default_rng(4242)
    t = rng.integers(0, 256, size=FULL_SLOTS)
    p = (t + rng.integers(-3, 4, size=FULL_SLOTS)) % 256
    t[[5, 77]] = 0
    return t.astype(np.int64), p.astype(np.int64)


def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all", "ref"])
    K = O.Keys.generate(O.default_params(128), KEY_SEED)
    side = R.RefSide(K)
    if "--full" not in sys.argv:
        out = {"key_fingerprint": np.array(R.key_fingerprint(K))}
        for name, _ in R.CASES:
            t0 = time.time()
            words, counters = R.run_case(side, K, name)
            out[name + "/words"] = words
            out[name + "/counters"] = counters
            out[name + "/bits"] = K.decrypt(words).astype(np.int8)
            print(f"{name}: {words.shape[0]} samples, gates/pbs/levels {counters.tolist()}, {time.time() - t0:.1f} s")
        np.savez_compressed(R.GOLDEN, **out)
        print("wrote", R.GOLDEN, os.path.getsize(R.GOLDEN), "bytes")
        return
    # full size: HE_EuclideanDistance over 128 slots (Math.cpp:333), then Function_f's own second step on its result
    # (Math.cpp:384: minimum(tmp_min, result_b, ed, bound_match, 24)) for two bounds
    import ctypes as C
    t, p = full_vectors()
    enc = R.encoder(K, "full")
    A = [R._int_in(side, enc, int(v), 8, f"t{i}") for i, v in enumerate(t)]
    B = [R._int_in(side, enc, int(v), 8, f"p{i}") for i, v in enumerate(p)]
    pa, pb = (R.LS * FULL_SLOTS)(*A), (R.LS * FULL_SLOTS)(*B)
    ed = side.new_array(24)
    c0 = side.counters()
    t0 = time.time()
    side.fn("euclidean")(ed, pa, pb, FULL_SLOTS, 8, side.cloud)
    ed_words = side.get_words(ed, 24)
    c1 = side.counters()
    print(f"euclidean over {FULL_SLOTS} slots: gates/pbs/levels {(c1 - c0).tolist()}, {time.time() - t0:.0f} s")
    dist = R.bits_value(K.decrypt(ed_words)[:23])
    clear = R.load_ref().refmath_clear_euclidean_bitwise(
        (C.c_uint64 * FULL_SLOTS)(*[int(v) for v in t]), (C.c_uint64 * FULL_SLOTS)(*[int(v) for v in p]), FULL_SLOTS, 8)
    print("decrypted distance", dist, "reference's plaintext mirror (EuclideanDistanceBitwise)", clear)
    out = {"key_fingerprint": np.array(R.key_fingerprint(K)), "template": t, "probe": p, "ed/words": ed_words,
           "ed/counters": c1 - c0, "distance": np.array(dist), "ed/level_widths": np.array(side.last_level_widths())}
    for bound in (dist, dist - 1):
        eb = R._int_in(side, enc, int(bound), 24, f"bound{bound}")
        tmp_min, rb = side.new_array(24), side.new_array(24)
        side.fn("minimum")(tmp_min, rb, ed, eb, 24, side.cloud)
        out[f"result_b/{'ge' if bound == dist else 'lt'}/words"] = side.get_words(rb, 24)
        out[f"result_b/{'ge' if bound == dist else 'lt'}/bound"] = np.array(bound)
        print("bound", bound, "-> verdict bit", int(K.decrypt(side.get_words(rb, 1))[0]))
    np.savez_compressed(R.GOLDEN_FULL, **out)
    print("wrote", R.GOLDEN_FULL, os.path.getsize(R.GOLDEN_FULL), "bytes")


if __name__ == "__main__":
    main()
