"""Bit-level simulator behind peba1_host_set_test_backend — TEST INFRASTRUCTURE for the CPU suite.

It lets the recorder / level scheduler / Math.h circuit builder of libpeba1_tfhe.so run without a GPU: every pool slot
holds one plaintext bit; uploads are decrypted with the test's secret key, downloads return trivial samples
(a = 0, b = +-1/8), gates apply their truth table (honouring the NEG / CONST operand flags).  It also records the
schedule (gates per level) so tests can compare it with the reference's structure (SURVEY.md §8a [MODEL] numbers).
"""
import ctypes as C

import numpy as np

from peba1_b200 import tfhe_host as H
from peba1_b200.cuda import GATE_DTYPE

TRUTH = {0: lambda a, b: 1 - (a & b), 1: lambda a, b: a & b, 2: lambda a, b: a | b, 3: lambda a, b: a ^ b,
         4: lambda a, b: 1 - (a ^ b), 5: lambda a, b: 1 - (a | b), 6: lambda a, b: (1 - a) & b,
         7: lambda a, b: a & (1 - b), 8: lambda a, b: (1 - a) | b, 9: lambda a, b: a | (1 - b)}


class SimBackend:
    def __init__(self):
        self.keyset = None
        self.bits = np.zeros(0, dtype=np.int8)
        self.level_widths = []      # gates per peba1_cuda_gates call
        self.level_pbs = []
        self.op_counts = {}
        self.max_slot = -1
        self.struct = H.TestBackend(None, H.TestBackend.RESERVE(self._reserve), H.TestBackend.UPLOAD(self._upload),
                                    H.TestBackend.DOWNLOAD(self._download), H.TestBackend.GATES(self._gates))

    def install(self):
        H.load_library().peba1_host_set_test_backend(C.byref(self.struct))

    def reset_schedule(self):
        self.level_widths, self.level_pbs, self.op_counts = [], [], {}

    def _reserve(self, user, slots):
        if slots > self.bits.size:
            self.bits = np.concatenate([self.bits, np.zeros(slots - self.bits.size, dtype=np.int8)])
        return 0

    def _upload(self, user, slots, count, words):
        w = self.keyset.words
        arr = np.ctypeslib.as_array(words, shape=(count, w))
        dec = self.keyset.decrypt(arr)
        for i in range(count):
            self.bits[slots[i]] = dec[i]
        return 0

    def _download(self, user, slot, words):
        w = self.keyset.words
        arr = np.ctypeslib.as_array(words, shape=(w,))
        arr[:] = 0
        arr[w - 1] = (1 << 29) if self.bits[slot] else -(1 << 29)
        return 0

    def _operand(self, g, q):
        v = int(g["in"][q]) if (g["flags"] >> (4 + q)) & 1 else int(self.bits[g["in"][q]])
        return 1 - v if (g["flags"] >> q) & 1 else v

    def _gates(self, user, gates, count):
        arr = np.frombuffer((C.c_char * (count * GATE_DTYPE.itemsize)).from_address(gates), dtype=GATE_DTYPE)
        out = []
        pbs = 0
        for g in arr:
            op = int(g["op"])
            self.op_counts[op] = self.op_counts.get(op, 0) + 1
            if op == 10:
                s, t, f = (self._operand(g, q) for q in range(3))
                out.append(t if s else f)
                pbs += 2
            elif op in (11, 12):
                a, b, c = (self._operand(g, q) for q in range(3))
                out.append(a ^ b ^ c if op == 11 else int(a + b + c >= 2))
                pbs += 1
            else:
                out.append(TRUTH[op](self._operand(g, 0), self._operand(g, 1)))
                pbs += 1
        for g, v in zip(arr, out):   # outputs written after all inputs were read, like the two-kernel GPU path
            self.bits[g["out"]] = v
            self.max_slot = max(self.max_slot, int(g["out"]))
        self.level_widths.append(int(count))
        self.level_pbs.append(pbs)
        return 0
