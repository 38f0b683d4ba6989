"""<tfhe/tfhe_io.h>: the client / server split the reference's single-process demo never makes.
Client: keygen, exports the cloud keyset and the encrypted operands.  Server: loads both, evaluates an 8-bit adder with
Math.h, exports the result ciphertexts.  Client: loads them and decrypts with the reloaded SECRET keyset.
Runs on the bit-level simulator (no GPU); the file formats are exercised for real."""
import os
import subprocess
import sys

import numpy as np
import pytest

from peba1_b200 import tfhe_host as H
from sim_backend import SimBackend


@pytest.fixture(scope="module")
def env():
    L = H.load_library()
    sim = SimBackend()
    sim.install()
    ks = H.KeySet.generate(128, seed=11)
    sim.keyset = ks
    yield L, ks, sim
    L.peba1_host_set_test_backend(None)


def test_ciphertext_file_layout(env, tmp_path):
    """int32 type id 42, n mask words, body, double variance: 2,536 bytes at n = 630 (upstream's write_lweSample layout,
    [UPSTREAM-RECALL])."""
    L, ks, sim = env
    w = ks.encrypt([1], 3)
    a = ks.new_array(1)
    ks.set_words(a, w)
    path = str(tmp_path / "ct.bin")
    with H.open_file(path, "wb") as F:
        L.export_gate_bootstrapping_ciphertext_toFile(F, a, ks.params_p)
    raw = open(path, "rb").read()
    assert len(raw) == 4 + 4 * 631 + 8
    assert int.from_bytes(raw[:4], "little") == 42
    assert np.array_equal(np.frombuffer(raw[4:4 + 4 * 631], dtype="<i4"), w[0])
    assert np.frombuffer(raw[-8:], dtype="<f8")[0] == 0.0
    b = ks.new_array(1)
    with H.open_file(path, "rb") as F:
        L.import_gate_bootstrapping_ciphertext_fromFile(F, b, ks.params_p)
    assert np.array_equal(ks.get_words(b, 1), w)
    ks.free_array(1, a)
    ks.free_array(1, b)


def test_client_server_split_through_files(env, tmp_path):
    L, ks, sim = env
    x, y = 201, 77
    secret, cloud, req, resp = (str(tmp_path / n) for n in ("secret.key", "cloud.key", "request.ct", "response.ct"))
    # ---- client
    with H.open_file(secret, "wb") as F:
        L.export_tfheGateBootstrappingSecretKeySet_toFile(F, ks.sk)
    with H.open_file(cloud, "wb") as F:
        L.export_tfheGateBootstrappingCloudKeySet_toFile(F, ks.cloud)
    ca, cb = ks.encrypt_int(x, 8, 21), ks.encrypt_int(y, 8, 22)
    with H.open_file(req, "wb") as F:
        for arr in (ca, cb):
            for i in range(8):
                L.export_gate_bootstrapping_ciphertext_toFile(F, ks.at(arr, i), ks.params_p)
    assert os.path.getsize(cloud) > 113_000_000 and os.path.getsize(req) == 16 * 2536
    # ---- server: only the cloud keyset and the request
    with H.open_file(cloud, "rb") as F:
        ck = L.new_tfheGateBootstrappingCloudKeySet_fromFile(F)
    assert ck.contents.params.contents.n == 630
    sa, sb, sr, sc = (L.new_gate_bootstrapping_ciphertext_array(k, ck.contents.params) for k in (8, 8, 8, 1))
    with H.open_file(req, "rb") as F:
        for arr in (sa, sb):
            for i in range(8):
                L.import_gate_bootstrapping_ciphertext_fromFile(F, L.peba1_sample_at(arr, i), ck.contents.params)
    L.peba1_math_add(sr, sa, sb, sc, 8, ck)   # bootsADDNbit(result, a, b, carry, bitsize, bk), src/Math.cpp:54
    with H.open_file(resp, "wb") as F:
        for i in range(8):   # the results are still device-resident handles: export fetches them
            L.export_gate_bootstrapping_ciphertext_toFile(F, L.peba1_sample_at(sr, i), ck.contents.params)
        L.export_gate_bootstrapping_ciphertext_toFile(F, sc, ck.contents.params)
    for k, arr in ((8, sa), (8, sb), (8, sr), (1, sc)):
        L.delete_gate_bootstrapping_ciphertext_array(k, arr)
    L.delete_gate_bootstrapping_cloud_keyset(ck)
    # ---- client again: a fresh process would start from the secret key file
    with H.open_file(secret, "rb") as F:
        sk2 = L.new_tfheGateBootstrappingSecretKeySet_fromFile(F)
    p2 = sk2.contents.params
    out = L.new_gate_bootstrapping_ciphertext_array(9, p2)
    with H.open_file(resp, "rb") as F:
        for i in range(9):
            L.import_gate_bootstrapping_ciphertext_fromFile(F, L.peba1_sample_at(out, i), p2)
    bits = [L.bootsSymDecrypt(L.peba1_sample_at(out, i), sk2) for i in range(9)]
    assert sum(b << i for i, b in enumerate(bits[:8])) == (x + y) % 256 and bits[8] == (x + y) >> 8
    L.delete_gate_bootstrapping_ciphertext_array(9, out)
    L.delete_gate_bootstrapping_secret_keyset(sk2)
    ks.free_array(8, ca)
    ks.free_array(8, cb)


def test_parameter_file_round_trip_and_rejections(env, tmp_path):
    L, ks, sim = env
    path = str(tmp_path / "params.bin")
    with H.open_file(path, "wb") as F:
        L.export_tfheGateBootstrappingParameterSet_toFile(F, ks.params_p)
    with H.open_file(path, "rb") as F:
        p = L.new_tfheGateBootstrappingParameterSet_fromFile(F)
    for f in ("n", "N", "k", "l", "Bgbit", "ks_t", "ks_basebit", "ks_stdev", "bk_stdev", "minimum_lambda"):
        assert getattr(p.contents, f) == getattr(ks.params_p.contents, f), f
    L.delete_gate_bootstrapping_parameters(p)
    # fatal errors abort with a message, like every error of this API: checked in a child process
    raw = bytearray(open(path, "rb").read())
    cases = {"bad magic": bytes([0x51]) + bytes(raw[1:]), "checksum mismatch": bytes(raw[:20]) + bytes([raw[20] ^ 1]) + bytes(raw[21:]),
             "unexpected end of file": bytes(raw[:-3]), "different kind": None}
    for msg, blob in cases.items():
        bad = str(tmp_path / "bad.bin")
        open(bad, "wb").write(blob if blob is not None else bytes(raw))
        fn = "new_tfheGateBootstrappingCloudKeySet_fromFile" if blob is None else "new_tfheGateBootstrappingParameterSet_fromFile"
        code = (f"import sys; sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r});"
                f"from peba1_b200 import tfhe_host as H; L = H.load_library();\n"
                f"with H.open_file({bad!r}, 'rb') as F: L.{fn}(F)")
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
        assert r.returncode != 0 and "peba1 tfhe_io" in r.stderr and msg.split()[0] in r.stderr, (msg, r.stderr)
