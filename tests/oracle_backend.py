"""Oracle-backed test backend — TEST INFRASTRUCTURE.  Runs the recorder's schedule gate by gate on the CPU oracle
(raw Torus32 words), so a whole circuit evaluated by the GPU engine can be compared word for word with a faithful
per-gate CPU evaluation of the same DAG."""
import ctypes as C
import os

import numpy as np

from peba1_b200 import tfhe_host as H
from peba1_b200.cuda import GATE_DTYPE

OPNAMES = ["NAND", "AND", "OR", "XOR", "XNOR", "NOR", "ANDNY", "ANDYN", "ORNY", "ORYN", "MUX"]


class OracleBackend:
    def __init__(self, oracle_keys):
        self.K = oracle_keys
        self.words = oracle_keys.words
        self.pool = np.zeros((0, self.words), dtype=np.int32)
        self.levels = 0
        self.struct = H.TestBackend(None, H.TestBackend.RESERVE(self._reserve), H.TestBackend.UPLOAD(self._upload),
                                    H.TestBackend.DOWNLOAD(self._download), H.TestBackend.GATES(self._gates))

    def install(self):
        H.load_library().peba1_host_set_test_backend(C.byref(self.struct))

    @staticmethod
    def uninstall():
        H.load_library().peba1_host_set_test_backend(None)

    def _reserve(self, user, slots):
        if slots > self.pool.shape[0]:
            fresh = np.zeros((slots - self.pool.shape[0], self.words), dtype=np.int32)
            fresh[:, -1] = -(1 << 29)
            self.pool = np.concatenate([self.pool, fresh])
        return 0

    def _upload(self, user, slots, count, words):
        arr = np.ctypeslib.as_array(words, shape=(count, self.words))
        for i in range(count):
            self.pool[slots[i]] = arr[i]
        return 0

    def _download(self, user, slot, words):
        np.ctypeslib.as_array(words, shape=(self.words,))[:] = self.pool[slot]
        return 0

    def _operand(self, g, q):
        if (g["flags"] >> (4 + q)) & 1:
            v = self.K.trivial(int(g["in"][q]))
        else:
            v = self.pool[g["in"][q]]
        return (-v).astype(np.int32) if (g["flags"] >> q) & 1 else v

    def _gates(self, user, gates, count):
        arr = np.frombuffer((C.c_char * (count * GATE_DTYPE.itemsize)).from_address(gates), dtype=GATE_DTYPE)
        threads = os.cpu_count() or 1
        outs = np.empty((count, self.words), dtype=np.int32)
        for op in np.unique(arr["op"]):
            idx = np.nonzero(arr["op"] == op)[0]
            a = np.stack([self._operand(arr[i], 0) for i in idx])
            b = np.stack([self._operand(arr[i], 1) for i in idx])
            c = np.stack([self._operand(arr[i], 2) for i in idx]) if op >= 10 else None
            if op in (11, 12):   # XOR3 = bootstrap(2(a+b+c) + 1/2), MAJ3 = bootstrap(a+b+c): prologue here, rest by the oracle
                s = 2 if op == 11 else 1
                t = (s * (a.astype(np.int64) + b + c)) & 0xFFFFFFFF
                if op == 11:
                    t[:, -1] = (t[:, -1] + (1 << 31)) & 0xFFFFFFFF
                t = t.astype(np.uint32).view(np.int32)
                outs[idx] = np.stack([self.K.keyswitch(self.K.bootstrap_woks(row)) for row in t])
                continue
            outs[idx] = self.K.gate_batch(OPNAMES[int(op)], a, b, c, threads=threads)
        for i in range(count):
            self.pool[arr[i]["out"]] = outs[i]
        self.levels += 1
        return 0
