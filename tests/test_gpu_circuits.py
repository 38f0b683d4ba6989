"""GPU tests of the whole stack: TFHE-compatible front end + Math.h circuits (libpeba1_tfhe.so) on libpeba1_cuda.

Decrypted results are checked against native integer arithmetic (the reference's own self-test comparisons,
src/main.cpp:320,352,383,422,456,578) and raw output ciphertexts against a faithful per-gate evaluation of the same
DAG on the CPU oracle (tests/oracle_backend.py) — bit-exact."""
import ctypes as C

import numpy as np
import pytest

from peba1_b200 import tfhe_host as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu_keys():
    H.load_library().peba1_host_set_test_backend(None)
    ks = H.KeySet.generate(128, seed=42)
    yield ks


@pytest.fixture(scope="module")
def cpu_twin(gpu_keys, oracle):
    """the same key material behind an oracle-backed engine"""
    from oracle_backend import OracleBackend
    K = oracle.Keys.from_arrays(oracle.default_params(128), gpu_keys.lwe_key, gpu_keys.tlwe_key, gpu_keys.bk,
                                gpu_keys.ksk)
    be = OracleBackend(K)
    be.install()
    L = H.load_library()
    p = L.new_default_gate_bootstrapping_parameters(128)
    i32p = C.POINTER(C.c_int32)
    sk = L.peba1_new_keyset_from_arrays(p, gpu_keys.lwe_key.ctypes.data_as(i32p), gpu_keys.tlwe_key.ctypes.data_as(i32p),
                                        gpu_keys.bk.ctypes.data_as(i32p), gpu_keys.ksk.ctypes.data_as(i32p))
    twin = H.KeySet(p, sk)
    twin._oracle_keys = K
    # force the engine creation while the oracle backend is installed, then restore the product default
    tmp = twin.new_array(2)
    L.bootsAND(twin.at(tmp, 0), twin.at(tmp, 0), twin.at(tmp, 1), twin.cloud)
    twin.get_words(tmp, 1)
    OracleBackend.uninstall()
    yield twin
    K.close()


def test_uses_the_gpu(gpu_keys):
    L = H.load_library()
    assert L.peba1_cuda_context_of(gpu_keys.cloud)
    maps = open("/proc/self/maps").read()
    assert "libpeba1_cuda.so" in maps and "libpeba1_tfhe.so" in maps


def test_add8_values_and_raw_parity(gpu_keys, cpu_twin):
    L = H.load_library()
    rng = np.random.default_rng(1)
    for trial in range(2):
        x, y = (int(v) for v in rng.integers(0, 256, size=2))
        wa = gpu_keys.encrypt([(x >> i) & 1 for i in range(8)], 100 + trial)
        wb = gpu_keys.encrypt([(y >> i) & 1 for i in range(8)], 200 + trial)
        outs = []
        for ks in (gpu_keys, cpu_twin):
            a, b, r, c = ks.new_array(8), ks.new_array(8), ks.new_array(8), ks.new_array(1)
            ks.set_words(a, wa); ks.set_words(b, wb)
            L.peba1_math_add(r, a, b, c, 8, ks.cloud)
            outs.append(np.concatenate([ks.get_words(r, 8), ks.get_words(c, 1)]))
            for arr, n in ((a, 8), (b, 8), (r, 8), (c, 1)):
                ks.free_array(n, arr)
        assert np.array_equal(outs[0], outs[1])                      # raw ciphertext words, GPU == per-gate oracle
        bits = gpu_keys.decrypt(outs[0])
        assert sum(int(v) << i for i, v in enumerate(bits)) == x + y  # 9 bits incl. carry


def test_sub_and_minimum_raw_parity(gpu_keys, cpu_twin):
    L = H.load_library()
    x, y = 77, 0                                                      # exercises quirk Q1 (b = 0)
    wa = gpu_keys.encrypt([(x >> i) & 1 for i in range(8)], 301)
    wb = gpu_keys.encrypt([(y >> i) & 1 for i in range(8)], 302)
    outs = []
    for ks in (gpu_keys, cpu_twin):
        a, b, r = ks.new_array(8), ks.new_array(8), ks.new_array(9)
        m, bit = ks.new_array(8), ks.new_array(8)
        ks.set_words(a, wa); ks.set_words(b, wb)
        L.peba1_batch_begin()
        L.peba1_math_sub(r, a, b, 8, ks.cloud)
        L.peba1_math_minimum(m, bit, a, b, 8, ks.cloud)               # MUX path: 2 bootstraps + 1 keyswitch
        L.peba1_batch_end()
        outs.append(np.concatenate([ks.get_words(r, 9), ks.get_words(m, 8), ks.get_words(bit, 1)]))
        for arr, n in ((a, 8), (b, 8), (r, 9), (m, 8), (bit, 8)):
            ks.free_array(n, arr)
    assert np.array_equal(outs[0], outs[1])
    bits = gpu_keys.decrypt(outs[0])
    val = lambda bs: sum(int(v) << i for i, v in enumerate(bs))
    assert val(bits[:9]) == 256 - 77 and val(bits[9:17]) == 0 and bits[17] == 1


def test_multiply8_batch_of_instances(gpu_keys):
    """BASELINE config 3 shape at reduced instance count: 32 multipliers batched per level (81 launches pairs)."""
    L = H.load_library()
    n = 32
    rng = np.random.default_rng(2)
    xs, ys = rng.integers(0, 256, size=n), rng.integers(0, 256, size=n)
    A = [gpu_keys.encrypt_int(int(x), 8, 1000 + i) for i, x in enumerate(xs)]
    B = [gpu_keys.encrypt_int(int(y), 8, 2000 + i) for i, y in enumerate(ys)]
    R = [gpu_keys.new_array(23) for _ in range(n)]
    s0 = gpu_keys.stats()
    L.peba1_batch_begin()
    for i in range(n):
        L.peba1_math_multiply(R[i], A[i], B[i], 8, 0, gpu_keys.cloud)
    L.peba1_batch_end()
    s1 = gpu_keys.stats()
    assert s1.gates - s0.gates == 1296 * n and s1.levels - s0.levels == 81
    for i in range(n):
        assert gpu_keys.decrypt_int(R[i], 23) == int(xs[i]) * int(ys[i])          # main.cpp:456
    for i in range(n):
        gpu_keys.free_array(8, A[i]); gpu_keys.free_array(8, B[i]); gpu_keys.free_array(23, R[i])


def test_face_match_function_f_and_g(gpu_keys):
    """One full-size face match (128 slots x 8 bits, 215,544 bootstraps) + Function_g, as main.cpp:513-581 does."""
    L = H.load_library()
    rng = np.random.default_rng(3)
    t = rng.integers(0, 255, size=128)
    probe = (t + 1) % 256
    def vec(values, seed):
        arrs = [gpu_keys.encrypt_int(int(v), 8, seed + i) for i, v in enumerate(values)]
        return arrs, (C.POINTER(H.LweSample) * len(arrs))(*arrs)
    ta, tp = vec(t, 5000)
    pa, pp = vec(probe, 6000)
    for bound, expect in ((100, 1), (128, 0)):                        # distance is exactly 128
        eb = gpu_keys.encrypt_int(bound, 24, 7000)
        rb = gpu_keys.new_array(24)
        s0 = gpu_keys.stats()
        L.peba1_math_function_f(rb, pp, tp, 128, eb, 8, gpu_keys.cloud)
        s1 = gpu_keys.stats()
        assert (s1.gates - s0.gates, s1.pbs - s0.pbs, s1.levels - s0.levels) == (215496, 215544, 377)
        assert gpu_keys.decrypt_int(rb, 24) == expect
        r0, r1 = 0x3C, 0xA7
        e0, e1 = gpu_keys.encrypt_int(r0, 8, 8000), gpu_keys.encrypt_int(r1, 8, 8001)
        y = gpu_keys.new_array(9)
        L.peba1_math_function_g(y, rb, e0, e1, 8, gpu_keys.cloud)
        assert gpu_keys.decrypt_int(y, 8) == (r1 if expect else (-r0) % 256)   # main.cpp:578 (+ Q1)
        gpu_keys.free_array(24, eb); gpu_keys.free_array(24, rb); gpu_keys.free_array(9, y)
        gpu_keys.free_array(8, e0); gpu_keys.free_array(8, e1)
    for arr in ta + pa:
        gpu_keys.free_array(8, arr)


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_fold_modes_on_gpu(gpu_keys, mode):
    """SURVEY.md §8f item 1 on real ciphertexts: fewer bootstraps, identical decrypted results (noise included)."""
    L = H.load_library()
    L.peba1_set_mode(mode)
    try:
        rng = np.random.default_rng(60 + mode)
        n = 24
        xs, ys = rng.integers(0, 256, size=n), rng.integers(0, 256, size=n)
        ys[0] = 0                                                      # quirk Q1 must survive folding
        A = [gpu_keys.encrypt_int(int(x), 8, 3000 + i) for i, x in enumerate(xs)]
        B = [gpu_keys.encrypt_int(int(y), 8, 4000 + i) for i, y in enumerate(ys)]
        D = [gpu_keys.new_array(9) for _ in range(n)]
        M = [gpu_keys.new_array(23) for _ in range(n)]
        s0 = gpu_keys.stats()
        L.peba1_batch_begin()
        for i in range(n):
            L.peba1_math_sub(D[i], A[i], B[i], 8, gpu_keys.cloud)
            L.peba1_math_multiply(M[i], A[i], B[i], 8, 0, gpu_keys.cloud)
        L.peba1_batch_end()
        s1 = gpu_keys.stats()
        assert s1.executed_pbs - s0.executed_pbs < 0.62 * n * (226 + 1296)
        for i in range(n):
            x, y = int(xs[i]), int(ys[i])
            assert gpu_keys.decrypt_int(D[i], 9) == (abs(x - y) if y else (256 - x) % 512)
            assert gpu_keys.decrypt_int(M[i], 23) == x * y
        for i in range(n):
            gpu_keys.free_array(8, A[i]); gpu_keys.free_array(8, B[i])
            gpu_keys.free_array(9, D[i]); gpu_keys.free_array(23, M[i])
    finally:
        L.peba1_set_mode(0)


def test_config3_adders_4096_instances(gpu_keys):
    """BASELINE configs[2]: 8- and 16-bit ripple-carry adders, 4,096 instances batched per level (24 / 48 launches pairs);
    every instance checked against native integer arithmetic (main.cpp:320)."""
    L = H.load_library()
    for bits, levels in ((8, 24), (16, 48)):
        n = 4096
        rng = np.random.default_rng(bits)
        xs, ys = rng.integers(0, 1 << bits, size=n), rng.integers(0, 1 << bits, size=n)
        wa = gpu_keys.encrypt(((xs[:, None] >> np.arange(bits)) & 1).reshape(-1), 11).reshape(n, bits, -1)
        wb = gpu_keys.encrypt(((ys[:, None] >> np.arange(bits)) & 1).reshape(-1), 12).reshape(n, bits, -1)
        A, B = [gpu_keys.new_array(bits) for _ in range(n)], [gpu_keys.new_array(bits) for _ in range(n)]
        R, Cy = [gpu_keys.new_array(bits) for _ in range(n)], [gpu_keys.new_array(1) for _ in range(n)]
        for i in range(n):
            gpu_keys.set_words(A[i], wa[i]); gpu_keys.set_words(B[i], wb[i])
        s0 = gpu_keys.stats()
        L.peba1_batch_begin()
        for i in range(n):
            L.peba1_math_add(R[i], A[i], B[i], Cy[i], bits, gpu_keys.cloud)
        L.peba1_batch_end()
        s1 = gpu_keys.stats()
        assert s1.executed_pbs - s0.executed_pbs == 7 * bits * n and s1.levels - s0.levels == levels
        out = np.stack([gpu_keys.get_words(R[i], bits) for i in range(n)])
        got = (gpu_keys.decrypt(out.reshape(n * bits, -1)).reshape(n, bits) << np.arange(bits)).sum(axis=1)
        assert np.array_equal(got, (xs + ys) % (1 << bits))
        for i in range(n):
            for arr, k in ((A[i], bits), (B[i], bits), (R[i], bits), (Cy[i], 1)):
                gpu_keys.free_array(k, arr)


def test_config3_multiplier16_instances(gpu_keys):
    """BASELINE configs[2]: the 16-bit multiplier (width 47 — the reference's hard-coded 23 parametrised), 64 instances
    batched per level: 169 launch pairs, 346 k bootstraps."""
    L = H.load_library()
    n = 64
    rng = np.random.default_rng(16)
    xs, ys = rng.integers(0, 1 << 16, size=n), rng.integers(0, 1 << 16, size=n)
    A = [gpu_keys.encrypt_int(int(x), 16, 100 + i) for i, x in enumerate(xs)]
    B = [gpu_keys.encrypt_int(int(y), 16, 900 + i) for i, y in enumerate(ys)]
    R = [gpu_keys.new_array(47) for _ in range(n)]
    s0 = gpu_keys.stats()
    L.peba1_batch_begin()
    for i in range(n):
        L.peba1_math_multiply(R[i], A[i], B[i], 16, 47, gpu_keys.cloud)
    L.peba1_batch_end()
    s1 = gpu_keys.stats()
    assert s1.executed_pbs - s0.executed_pbs == 5408 * n and s1.levels - s0.levels == 169
    for i in range(n):
        assert gpu_keys.decrypt_int(R[i], 47) == int(xs[i]) * int(ys[i])
    for i in range(n):
        gpu_keys.free_array(16, A[i]); gpu_keys.free_array(16, B[i]); gpu_keys.free_array(47, R[i])


def test_batched_authentication_with_device_generated_templates(gpu_keys):
    """SURVEY §8f items 2 + 4 on the GPU: enrolled templates encrypted on the device, three clients authenticated in one
    batch (Function_f then Function_g per client, levels shared); decrypted verdicts follow main.cpp:578 with the
    reference's polarity and quirks."""
    from peba1_b200 import gallery as G
    ks = gpu_keys
    rng = np.random.default_rng(8)
    K, slots, bound = 3, 8, 120
    tmpl = rng.integers(1, 200, size=(K, slots))
    probes = tmpl + np.array([[1], [4], [5]])            # squared distances 8, 128, 200
    r0, r1 = rng.integers(1, 256, size=K), rng.integers(0, 256, size=K)
    pv = [G._host_vector(ks, ks.encrypt(G.bits_of(probes[k], 8), 100 + k), slots, 8) for k in range(K)]
    s0 = ks.stats()
    tv = G.encrypt_templates_on_device(ks, tmpl, bits=8, seed=9)
    assert ks.stats().device_encryptions - s0.device_encryptions == K * slots * 8
    r0w = np.stack([ks.encrypt(G.bits_of([r0[k]], 8), 200 + k) for k in range(K)])
    r1w = np.stack([ks.encrypt(G.bits_of([r1[k]], 8), 300 + k) for k in range(K)])
    y = G.authenticate_batch(ks, pv, tv, ks.encrypt(G.bits_of([bound], 24), 1), r0w, r1w, slots=slots, bits=8)
    got = [int(sum(int(b) << i for i, b in enumerate(ks.decrypt(y[k])))) for k in range(K)]
    far = ((probes - tmpl) ** 2).sum(axis=1) > bound
    assert far.tolist() == [False, True, True]
    assert got == [int(r1[k]) if far[k] else int((-r0[k]) % 256) for k in range(K)]


def test_three_input_gates_and_mode3_face_match(gpu_keys, cpu_twin):
    """XOR3 / MAJ3 (one bootstrap each; the full adder of circuit mode 3) on real ciphertexts: truth tables on all 8
    triples, phase margins, raw words equal to the oracle's bootstrap of the same linear combination; then a full-size
    face match in mode 3 (about 19 k bootstraps in 49 levels instead of 215 k in 377) with the decrypted verdicts of the faithful circuit."""
    from oracle_backend import OracleBackend
    import peba1_b200 as P
    L = H.load_library()
    ks = gpu_keys
    eng = P.Engine.from_handle(ks.cuda_context(), P.default_params(128), device=0)
    bits = np.array([[a, b, c] for a in (0, 1) for b in (0, 1) for c in (0, 1)] * 8)       # 64 triples
    ca, cb, cc = (ks.encrypt(bits[:, q], 700 + q) for q in range(3))
    for op, truth in (("XOR3", bits.sum(axis=1) & 1), ("MAJ3", (bits.sum(axis=1) >= 2).astype(int))):
        got = eng.gate_batch_host(op, ca, cb, cc)
        assert np.array_equal(ks.decrypt(got), truth), op
        # raw parity: the oracle bootstraps the same prologue
        s = 2 if op == "XOR3" else 1
        t = (s * (ca.astype(np.int64) + cb + cc)) & 0xFFFFFFFF
        if op == "XOR3":
            t[:, -1] = (t[:, -1] + (1 << 31)) & 0xFFFFFFFF
        t = t.astype(np.uint32).view(np.int32)
        K = cpu_twin_oracle(cpu_twin)
        want = np.stack([K.keyswitch(K.bootstrap_woks(row)) for row in t[:6]])
        assert np.array_equal(got[:6], want), op
        # inputs of the bootstrap sit at least 1/16 away from the decision boundary (nominal 1/8 resp. 1/4)
        ph = K.phase(t).astype(np.float64) / 2.0**32
        assert np.all(np.abs(ph) > 1 / 16) and np.all(np.abs(ph) < 0.5 - 1 / 16), op
    # full-size face match, mode 3
    rng = np.random.default_rng(3)
    tmpl = rng.integers(0, 255, size=128)
    probe = (tmpl + 1) % 256
    def vec(values, seed):
        arrs = [ks.encrypt_int(int(v), 8, seed + i) for i, v in enumerate(values)]
        return arrs, (C.POINTER(H.LweSample) * len(arrs))(*arrs)
    ta, tp = vec(tmpl, 5000)
    pa, pp = vec(probe, 6000)
    L.peba1_set_mode(3)
    try:
        for bound, expect in ((100, 1), (128, 0)):                        # distance is exactly 128
            eb = ks.encrypt_int(bound, 24, 7000)
            rb = ks.new_array(24)
            s0 = ks.stats()
            L.peba1_math_function_f(rb, pp, tp, 128, eb, 8, ks.cloud)
            s1 = ks.stats()
            assert s1.executed_pbs - s0.executed_pbs < 19000 and s1.levels - s0.levels < 60
            assert ks.decrypt_int(rb, 24) == expect
            ks.free_array(24, eb); ks.free_array(24, rb)
    finally:
        L.peba1_set_mode(0)
    for arr in ta + pa:
        ks.free_array(8, arr)


def cpu_twin_oracle(twin):
    """the oracle key object behind the cpu_twin fixture (same key material as the GPU keys)"""
    return twin._oracle_keys
