"""GPU parity tests: the CUDA path through the C ABI (include/peba1_cuda.h) against the oracle, bit-exact.

Everything here runs on the B200 box (`-m gpu`); nothing reads /root/reference.
"""
import os

import numpy as np
import pytest

from test_oracle import TRUTH

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "gates_seed42.npz")


def test_native_library_is_loaded(engine):
    import peba1_b200 as P
    maps = open("/proc/self/maps").read()
    assert "libpeba1_cuda.so" in maps
    assert P.load_library().peba1_cuda_device_count() >= 1


def test_golden_pieces(engine):
    """bootstrap-without-keyswitch and keyswitch alone reproduce the exact-integer golden vectors."""
    g = np.load(GOLDEN)
    u = engine.bootstrap_woks_host(g["pbs_in"][None, :])
    assert np.array_equal(u[0], g["pbs_out"])
    out = engine.keyswitch_host(g["pbs_out"][None, :])
    assert np.array_equal(out[0], g["ks_out"])


def test_golden_gates_all_ops(engine, keys):
    g = np.load(GOLDEN)
    for op in TRUTH:
        out = engine.gate_batch_host(op, g["ca"], g["cb"])
        assert np.array_equal(out, g["out_" + op]), op
    out = engine.gate_batch_host("MUX", g["cs"], g["cx"], g["cy"])
    assert np.array_equal(out, g["out_MUX"])


@pytest.mark.parametrize("op", ["NAND", "XOR", "MUX"])
def test_random_batch_matches_oracle(engine, oracle, keys, op):
    """256 random gates per op vs the FFT-form oracle (itself pinned to the exact form in test_oracle.py)."""
    n = 256
    rng = np.random.default_rng(100 + len(op))
    a, b, c = (rng.integers(0, 2, size=n) for _ in range(3))
    ca, cb, cc = keys.encrypt(a, 31), keys.encrypt(b, 32), keys.encrypt(c, 33)
    if op == "MUX":
        got = engine.gate_batch_host(op, ca, cb, cc)
        want = keys.gate_batch(op, ca, cb, cc, threads=os.cpu_count() or 1)
        bits = np.where(a == 1, b, c)
    else:
        got = engine.gate_batch_host(op, ca, cb)
        want = keys.gate_batch(op, ca, cb, threads=os.cpu_count() or 1)
        bits = np.array([TRUTH[op](int(x), int(y)) for x, y in zip(a, b)])
    assert np.array_equal(got, want)
    assert np.array_equal(keys.decrypt(got), bits)


def test_ragged_and_empty_batches(engine, keys):
    assert engine.gate_batch_host("AND", np.zeros((0, 631), np.int32), np.zeros((0, 631), np.int32)).shape == (0, 631)
    for n in (1, 3, 5, 7):   # not a multiple of the teams per CTA
        a = np.arange(n) % 2
        b = (np.arange(n) // 2) % 2
        ca, cb = keys.encrypt(a, 41 + n), keys.encrypt(b, 51 + n)
        got = engine.gate_batch_host("OR", ca, cb)
        assert np.array_equal(got, keys.gate_batch("OR", ca, cb, threads=4))


def test_operand_flags_and_aliasing(engine, oracle, keys):
    """NEG / CONST operand modifiers (bootsNOT / bootsCONSTANT folded into the gate) and out == in aliasing."""
    from peba1_b200.cuda import make_gates, IN_NEG, IN_CONST
    a = np.array([0, 1, 0, 1]); b = np.array([0, 0, 1, 1])
    ca, cb = keys.encrypt(a, 61), keys.encrypt(b, 62)
    engine.reserve(64)
    engine.upload(0, ca); engine.upload(4, cb)
    # gate i: AND(NOT a_i, b_i) -> slot 8+i
    g1 = make_gates("AND", np.arange(4), 4 + np.arange(4), 8 + np.arange(4), flags=IN_NEG(0))
    # gate i: XOR(a_i, const 1) -> written over a_i itself (aliasing, reference Math.cpp:84,272)
    g2 = make_gates("XOR", np.arange(4), np.ones(4, dtype=np.int64), np.arange(4), flags=IN_CONST(1))
    engine.gates(np.concatenate([g1, g2]))
    got1, got2 = engine.download(8, 4), engine.download(0, 4)
    nota = np.stack([-ca[i] for i in range(4)]).astype(np.int32)
    want1 = keys.gate_batch("AND", nota, cb, threads=4)
    t1 = np.stack([keys.trivial(1)] * 4)
    want2 = keys.gate_batch("XOR", ca, t1, threads=4)
    assert np.array_equal(got1, want1) and np.array_equal(got2, want2)
    assert keys.decrypt(got1).tolist() == ((1 - a) & b).tolist()
    assert keys.decrypt(got2).tolist() == (1 - a).tolist()


def test_never_written_slot_is_trivial_false(engine, keys):
    """fresh pool slots read as (a=0, b=-1/8)."""
    _, _, cap = engine.pool_info()
    engine.reserve(cap + 16)
    fresh = engine.download(cap, 16)
    assert not fresh[:, :-1].any() and np.all(fresh[:, -1] == -(1 << 29))
    assert keys.decrypt(fresh).tolist() == [0] * 16


def test_linear_ops(engine, keys):
    """bootsNOT / bootsCOPY / bootsCONSTANT on device slots."""
    ct = keys.encrypt([1, 0], 71)
    engine.reserve(32)
    engine.upload(0, ct)
    mu = 1 << 29
    engine.linear([0, 1, -1, -1], [-1, 1, 0, 0], [0, 0, mu, -mu], [10, 11, 12, 13])
    got = engine.download(10, 4)
    assert np.array_equal(got[0], (-ct[0]).astype(np.int32)) and np.array_equal(got[1], ct[1])
    assert np.array_equal(got[2], keys.trivial(1)) and np.array_equal(got[3], keys.trivial(0))


def test_errors(engine):
    import peba1_b200 as P
    from peba1_b200.cuda import make_gates
    _, _, cap = engine.pool_info()
    with pytest.raises(P.PEBA1Error):
        engine.gates(make_gates("AND", [0], [cap + 5], [0]))
    with pytest.raises(P.PEBA1Error):
        engine.gates(make_gates(99, [0], [0], [0]))
    with pytest.raises(P.PEBA1Error):
        engine.download(cap, 1)


def test_host_buffer_calls_leave_the_pool_alone(engine, keys):
    """peba1_cuda_gate_batch_host / _keyswitch_host work on private rows: device-resident slots a caller keeps in the pool
    (the recorder allocates from slot 0) are not clobbered."""
    rng = np.random.default_rng(5)
    n = 12
    resident = keys.encrypt(rng.integers(0, 2, size=4 * n), 61)
    engine.reserve(4 * n)
    engine.upload(0, resident)
    a, b, c = (rng.integers(0, 2, size=n) for _ in range(3))
    ca, cb, cc = keys.encrypt(a, 62), keys.encrypt(b, 63), keys.encrypt(c, 64)
    assert np.array_equal(keys.decrypt(engine.gate_batch_host("XOR", ca, cb)), a ^ b)
    assert np.array_equal(keys.decrypt(engine.gate_batch_host("MUX", ca, cb, cc)), np.where(a == 1, b, c))
    u = engine.bootstrap_woks_host(ca)
    assert np.array_equal(engine.keyswitch_host(u), np.stack([keys.keyswitch(row) for row in u]))
    assert np.array_equal(engine.download(0, 4 * n), resident)


def test_in_place_level_larger_than_one_launch_is_refused(engine, keys):
    """Level semantics hold inside one launch pair; a chunked level whose later chunk reads what an earlier chunk writes
    would depend on the chunk size: PEBA1_EINVAL instead.  Disjoint chunks run fine."""
    import peba1_b200 as P
    from peba1_b200.cuda import make_gates
    rng = np.random.default_rng(6)
    n = 10
    a, b = rng.integers(0, 2, size=n), rng.integers(0, 2, size=n)
    engine.reserve(3 * n)
    engine.upload(0, keys.encrypt(a, 71))
    engine.upload(n, keys.encrypt(b, 72))
    idx = np.arange(n)
    engine.set_option("max_pbs_per_launch", 4)
    try:
        engine.gates(make_gates("AND", idx, n + idx, 2 * n + idx))                 # 3 chunks, outputs elsewhere
        assert np.array_equal(keys.decrypt(engine.download(2 * n, n)), a & b)
        with pytest.raises(P.PEBA1Error, match="in-place levels must fit one launch"):
            engine.gates(make_gates("AND", idx, n + idx, (idx + 5) % n))            # later chunks read earlier outputs
        assert np.array_equal(keys.decrypt(engine.download(0, n)), a)              # nothing was launched
    finally:
        engine.set_option("max_pbs_per_launch", 0)
    engine.gates(make_gates("XOR", idx, n + idx, idx))                              # in place within ONE launch: fine
    assert np.array_equal(keys.decrypt(engine.download(0, n)), a ^ b)


def test_level_log_and_device_source_upload(engine, keys):
    from peba1_b200.cuda import make_gates
    import torch
    rng = np.random.default_rng(7)
    n = 40
    a, b = rng.integers(0, 2, size=n), rng.integers(0, 2, size=n)
    words = torch.from_numpy(np.concatenate([keys.encrypt(a, 81), keys.encrypt(b, 82)])).cuda()
    engine.reserve(3 * n)
    engine.upload_slots_device(words.data_ptr(), 2 * n, first_slot=0)              # contiguous rows, no host bounce
    engine.upload_slots_device(words[:n].data_ptr(), n, slots=2 * n + np.arange(n)[::-1])   # scattered
    engine.sync()
    assert np.array_equal(keys.decrypt(engine.download(2 * n, n)), a[::-1])
    idx = np.arange(n)
    engine.level_log_begin()
    engine.gates(make_gates("NOR", idx, n + idx, 2 * n + idx))
    engine.gates(make_gates("MUX", idx[:5], n + idx[:5], 2 * n + idx[:5], in2=idx[:5]))
    g, p, br, ks = engine.level_log_end()
    assert g.tolist() == [n, 5] and p.tolist() == [n, 10] and (br > 0).all() and (ks > 0).all()
    assert np.array_equal(keys.decrypt(engine.download(2 * n + 5, n - 5)), (1 - (a | b))[5:])


def test_size_independent_properties_full_batch(engine, keys):
    """BASELINE config 2 size (16,384 gates): properties that need no oracle run.
    NAND(x, x) == NOT x after decryption; XOR(x, y) then XOR(., y) gives x back; all phases within 2^-5 of +-1/8."""
    n = 16384
    rng = np.random.default_rng(5)
    x, y = rng.integers(0, 2, size=n), rng.integers(0, 2, size=n)
    cx, cy = keys.encrypt(x, 81), keys.encrypt(y, 82)
    nx = engine.gate_batch_host("NAND", cx, cx)
    assert np.array_equal(keys.decrypt(nx), 1 - x)
    z = engine.gate_batch_host("XOR", cx, cy)
    back = engine.gate_batch_host("XOR", z, cy)
    assert np.array_equal(keys.decrypt(back), x)
    ph = keys.phase(back).astype(np.float64) / 2.0**32
    assert np.all(np.abs(np.abs(ph) - 0.125) < 2.0**-5)
    # determinism: the same launch twice gives identical words
    assert np.array_equal(z, engine.gate_batch_host("XOR", cx, cy))


def test_lambda80_parameter_set(oracle):
    """The second upstream parameter set (minimum_lambda <= 80: n=500, l=2, Bgbit=10) through the same kernels."""
    import peba1_b200 as P
    K = oracle.Keys.generate(oracle.default_params(80), 7)
    eng = P.Engine(P.default_params(80), device=0)
    eng.load_keys(K.bk, K.ksk)
    rng = np.random.default_rng(8)
    n = 37
    a, b, c = (rng.integers(0, 2, size=n) for _ in range(3))
    ca, cb, cc = K.encrypt(a, 1), K.encrypt(b, 2), K.encrypt(c, 3)
    for op in ("NAND", "XOR"):
        got = eng.gate_batch_host(op, ca, cb)
        assert np.array_equal(got, K.gate_batch(op, ca, cb, threads=os.cpu_count() or 1)), op
    got = eng.gate_batch_host("MUX", ca, cb, cc)
    assert np.array_equal(got, K.gate_batch("MUX", ca, cb, cc, threads=os.cpu_count() or 1))
    assert np.array_equal(K.decrypt(got), np.where(a == 1, b, c))
    for variant in (0, 2, 7, 8, 11):  # every compiled alternative gives the same words (0, 2, 7: VARIANTS=all builds only)
        if eng.try_variant(variant):
            assert np.array_equal(eng.gate_batch_host("MUX", ca, cb, cc), got), variant
    eng.close()
    K.close()


def test_kernel_variants_agree(engine, keys):
    """register accumulators (0), tensor-memory accumulators (1, default), shared-key (2): identical words; also the
    1/2/4/8 teams-per-CTA shapes the launcher picks for narrow levels."""
    rng = np.random.default_rng(77)
    n = 45
    a, b = rng.integers(0, 2, size=n), rng.integers(0, 2, size=n)
    ca, cb = keys.encrypt(a, 91), keys.encrypt(b, 92)
    want = keys.gate_batch("ANDYN", ca, cb, threads=os.cpu_count() or 1)
    try:
        # the product build holds variants 1 (default), 7, 8 and 11; a VARIANTS=all build also the ones that lost
        for variant in (0, 1, 2):
            if engine.try_variant(variant):
                assert np.array_equal(engine.gate_batch_host("ANDYN", ca, cb), want), variant
        engine.set_option("pair_min_batch", 1)   # also below one full wave; n is odd: a pair team carries a shadow
        for variant in (3, 4, 5, 6, 7, 8, 11):    # pair, key ring, 10 teams, persistent CTAs (1 / 2 buffers), tensor-memory exchange
            if engine.try_variant(variant):
                assert np.array_equal(engine.gate_batch_host("ANDYN", ca, cb), want), variant
                assert np.array_equal(engine.gate_batch_host("ANDYN", ca[:1], cb[:1]), want[:1]), (variant, "single")
        engine.set_option("kernel_variant", 1)
        for teams in (1, 2, 4, 8):
            engine.set_option("tm_teams", teams)
            assert np.array_equal(engine.gate_batch_host("ANDYN", ca, cb), want), teams
        from peba1_b200.cuda import PEBA1Error
        try:                                     # keyswitch forms that lost: compiled only with `make VARIANTS=all`
            for grp in (2, 4, 6, 8):
                engine.set_option("ks_group", grp)
                assert np.array_equal(engine.gate_batch_host("ANDYN", ca, cb), want), grp
            engine.set_option("ks_group", 1)
            engine.set_option("ks_tile_min", 1)  # the tiled keyswitch (key streamed once per tile of ciphertexts), also split
            for tile, parts in ((8, 1), (12, 1), (16, 1), (8, 4), (16, 2)):
                engine.set_option("ks_tile", tile)
                engine.set_option("ks_tile_parts", parts)
                assert np.array_equal(engine.gate_batch_host("ANDYN", ca, cb), want), (tile, parts)
        except PEBA1Error as e:
            assert "VARIANTS=all" in str(e)
        engine.set_option("ks_split", 4)         # narrow-level keyswitch split over 4 CTAs per ciphertext
        assert np.array_equal(engine.gate_batch_host("ANDYN", ca, cb), want), "ks_split"
        engine.set_option("ks_split", -1)
        # the keyswitch through a table of row sums (default: 4 digits per row) against one row per digit (0) and the
        # 2-digit table; whole ciphertexts per CTA (ks_split 0), several of them per CTA (the batch is not a multiple)
        for lut in (0, 2, 4):
            engine.set_option("ks_lut", lut)
            for split in (-1, 0, 8):
                engine.set_option("ks_split", split)
                assert np.array_equal(engine.gate_batch_host("ANDYN", ca, cb), want), ("ks_lut", lut, split)
        engine.set_option("ks_split", 0)
        engine.set_option("ks_lut_group_min", 1)
        for grp in (4, 1):
            engine.set_option("ks_lut_group", grp)
            assert np.array_equal(engine.gate_batch_host("ANDYN", ca, cb), want), ("ks_lut_group", grp)
            assert np.array_equal(engine.gate_batch_host("ANDYN", ca[:3], cb[:3]), want[:3]), ("ks_lut_group", grp, 3)
    finally:
        engine.set_option("ks_lut", 4)
        engine.set_option("ks_lut_group", 4)
        engine.set_option("ks_lut_group_min", -1)
        engine.set_option("ks_split", -1)
        engine.set_option("kernel_variant", 1)
        engine.set_option("tm_teams", 0)
        engine.set_option("ks_group", 1)
        engine.set_option("ks_tile", 0)
        engine.set_option("ks_tile_parts", 1)
        engine.set_option("ks_tile_min", 4096)
        engine.set_option("pair_min_batch", -1)


@pytest.mark.gpu
def test_device_side_encryption(engine, keys):
    """peba1_cuda_encrypt_bits: decrypts to the bits, noise of the stated width, uniform masks,
    a function of (seed, index) only."""
    n = 4096
    bits = np.random.default_rng(3).integers(0, 2, size=n)
    engine.reserve(3 * n)
    engine.encrypt_bits(keys.lwe_key, bits, 0, seed=1234)
    ct = engine.download(0, n)
    assert np.array_equal(keys.decrypt(ct), bits)
    ph = keys.phase(ct).astype(np.int64)                         # torus32 phase = +-2^29 + noise
    noise = (ph - np.where(bits == 1, 1 << 29, -(1 << 29))).astype(np.float64) / 2.0**32
    assert abs(noise.std() / 2.0**-15 - 1.0) < 0.1 and abs(noise.mean()) < 2.0**-15 * 0.1
    mask = ct[:, :-1].astype(np.uint32)
    assert abs((mask >> 31).mean() - 0.5) < 0.002 and abs((mask & 1).mean() - 0.5) < 0.002
    assert len(np.unique(mask[:, 0])) > n - 4
    # same (seed, index) -> same ciphertext, whatever the slot and the launch shape
    engine.encrypt_bits(keys.lwe_key, bits[100:200], 2 * n, seed=1234, index0=100)
    assert np.array_equal(engine.download(2 * n, 100), ct[100:200])
    slots = n + np.arange(50)[::-1]
    engine.encrypt_bits(keys.lwe_key, bits[:50], 0, seed=1234, slots=slots)
    assert np.array_equal(engine.download_slots(slots), ct[:50])
    engine.encrypt_bits(keys.lwe_key, bits[:8], 2 * n, seed=1235)
    assert not np.array_equal(engine.download(2 * n, 8), ct[:8])
    # fresh encryptions are valid gate inputs
    out = engine.gate_batch_host("NAND", ct[:64], ct[64:128])
    assert np.array_equal(keys.decrypt(out), 1 - (bits[:64] & bits[64:128]))


@pytest.mark.gpu
def test_short_key_every_kernel_variant_against_the_oracle(oracle):
    """n = 8 mask words (8 CMux steps): cheap enough to run EVERY blind-rotation variant and launch shape — fixed grids,
    persistent CTAs, 8 and 10 teams, pairs, the TMA/tcgen05.cp key ring — on ragged batches, word for word against the
    exact-integer oracle."""
    import peba1_b200 as P
    from peba1_b200.cuda import Params
    po = oracle.default_params(128)
    po.n = 8
    K = oracle.Keys.generate(po, 5)
    eng = P.Engine(Params(8, 1024, 1, 3, 7, 8, 2), device=0)
    eng.load_keys(K.bk, K.ksk)
    rng = np.random.default_rng(4)
    try:
        for n in (1, 37, 301):
            a, b, c = (rng.integers(0, 2, size=n) for _ in range(3))
            ca, cb, cc = K.encrypt(a, 1), K.encrypt(b, 2), K.encrypt(c, 3)
            want = K.gate_batch("MUX", ca, cb, cc, mode=oracle.EXACT, threads=os.cpu_count() or 1)
            assert np.array_equal(K.decrypt(want), np.where(a == 1, b, c))
            for min_batch in (-1, 1):
                eng.set_option("pair_min_batch", min_batch)
                for variant in range(12):
                    if eng.try_variant(variant):
                        assert np.array_equal(eng.gate_batch_host("MUX", ca, cb, cc), want), (n, min_batch, variant)
    finally:
        eng.close()
        K.close()
