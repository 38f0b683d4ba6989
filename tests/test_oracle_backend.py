"""CPU check of tests/oracle_backend.py itself (the GPU circuit tests rely on it): a 2-bit adder recorded by the
front end and executed gate by gate on the oracle decrypts correctly and honours NEG / CONST operand flags."""
import ctypes as C

import numpy as np

from peba1_b200 import tfhe_host as H
from oracle_backend import OracleBackend


def test_two_bit_adder_on_oracle_backend(oracle, keys):
    L = H.load_library()
    be = OracleBackend(keys)
    be.install()
    try:
        p = L.new_default_gate_bootstrapping_parameters(128)
        i32p = C.POINTER(C.c_int32)
        sk = L.peba1_new_keyset_from_arrays(p, keys.lwe_key.ctypes.data_as(i32p), keys.tlwe_key.ctypes.data_as(i32p),
                                            keys.bk.ctypes.data_as(i32p), keys.ksk.ctypes.data_as(i32p))
        ks = H.KeySet(p, sk)
        a, b = ks.encrypt_int(3, 2, 1), ks.encrypt_int(2, 2, 2)
        r, c, t = ks.new_array(2), ks.new_array(1), ks.new_array(2)
        L.peba1_math_add(r, a, b, c, 2, ks.cloud)                      # 14 gates, 6 levels
        L.bootsNOT(ks.at(t, 0), ks.at(a, 0), ks.cloud)
        L.bootsOR(ks.at(t, 1), ks.at(t, 0), ks.at(b, 0), ks.cloud)     # NOT a0 | b0 = 0
        assert ks.decrypt_int(r, 2) == (3 + 2) % 4 and ks.decrypt_int(c, 1) == 1
        assert ks.decrypt(ks.get_words(t, 2)).tolist() == [0, 0]
        assert be.levels >= 6
        # raw words: the adder's bit 0 is XOR(XOR(a0, b0), const 0) evaluated by the oracle directly
        wa, wb = ks.get_words(a, 2), ks.get_words(b, 2)
        x = keys.gate("XOR", wa[0], wb[0])
        assert np.array_equal(ks.get_words(r, 1)[0], keys.gate("XOR", x, keys.trivial(0)))
    finally:
        OracleBackend.uninstall()
