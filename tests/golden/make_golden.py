"""Generates tests/golden/gates_seed42.npz with the EXACT-integer oracle (oracle/tfhe_oracle.c, ORACLE_EXACT).

Run from the repo root:  python tests/golden/make_golden.py
The reference (lab-incert/peba1) holds no golden vectors for this path and its TFHE
dependency is absent, so these vectors pin the repo's own exact-integer restatement: the FFT-form oracle and the
CUDA kernels must reproduce every word.  Keys are NOT stored (114 MB); they are regenerated from the seed by
oracle_keygen, which is deterministic.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402

KEY_SEED = 42


def main():
    p = O.default_params(128)
    K = O.Keys.generate(p, KEY_SEED)
    a_bits = np.array([0, 0, 1, 1], dtype=np.int64)
    b_bits = np.array([0, 1, 0, 1], dtype=np.int64)
    ca = K.encrypt(a_bits, 1001)
    cb = K.encrypt(b_bits, 1002)
    out = {"key_seed": np.int64(KEY_SEED), "a_bits": a_bits, "b_bits": b_bits, "ca": ca, "cb": cb}
    for op in ["NAND", "AND", "OR", "XOR", "XNOR", "NOR", "ANDNY", "ANDYN", "ORNY", "ORYN"]:
        res = np.stack([K.gate(op, ca[i], cb[i], mode=O.EXACT) for i in range(4)])
        out["out_" + op] = res
        print(op, K.decrypt(res))
    # MUX on all 8 triples
    s_bits = np.array([0, 0, 0, 0, 1, 1, 1, 1], dtype=np.int64)
    x_bits = np.array([0, 0, 1, 1, 0, 0, 1, 1], dtype=np.int64)
    y_bits = np.array([0, 1, 0, 1, 0, 1, 0, 1], dtype=np.int64)
    cs, cx, cy = K.encrypt(s_bits, 1003), K.encrypt(x_bits, 1004), K.encrypt(y_bits, 1005)
    res = np.stack([K.gate("MUX", cs[i], cx[i], cy[i], mode=O.EXACT) for i in range(8)])
    print("MUX", K.decrypt(res))
    out.update(s_bits=s_bits, x_bits=x_bits, y_bits=y_bits, cs=cs, cx=cx, cy=cy, out_MUX=res)
    # the pieces: bootstrap without keyswitch and keyswitch alone, on the NAND prologue of pair 3
    t = O.gate_prologue(p, "NAND", ca[3], cb[3])
    u = K.bootstrap_woks(t, mode=O.EXACT)
    out.update(pbs_in=t, pbs_out=u, ks_out=K.keyswitch(u))
    # checksums of the key material so a drifting keygen is detected
    out["bk_xor"] = np.bitwise_xor.reduce(K.bk.reshape(-1).view(np.uint32))
    out["ksk_xor"] = np.bitwise_xor.reduce(K.ksk.reshape(-1).view(np.uint32))
    out["bk_sum"] = np.uint32(K.bk.reshape(-1).view(np.uint32).sum(dtype=np.uint64) & 0xFFFFFFFF)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gates_seed42.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
