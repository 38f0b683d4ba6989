"""Latency of narrow levels: python scripts/prof_small.py"""
import sys; sys.path.insert(0, '.')
import numpy as np
import peba1_b200 as P
from peba1_b200.cuda import make_gates
eng = P.Engine()
rng = np.random.default_rng(1)
p = eng.params
bk = rng.integers(-2**31, 2**31, size=p.n*6*2*1024, dtype=np.int64).astype(np.int32)
ksk = rng.integers(-2**31, 2**31, size=1024*8*4*631, dtype=np.int64).astype(np.int32)
eng.load_keys(bk, ksk)
eng.enable_timing(True)
for n in (1, 8, 81, 162, 324, 648, 1184):
    cts = rng.integers(-2**31, 2**31, size=(2*n, 631), dtype=np.int64).astype(np.int32)
    eng.reserve(3*n); eng.upload(0, cts)
    g = make_gates("NAND", np.arange(n), n+np.arange(n), 2*n+np.arange(n))
    res = []
    for teams in (8, 0):
        eng.set_option("tm_teams", teams)
        for rep in range(2):
            eng.gates(g)
        res.append(eng.stats().last_blind_rotate_ms)
    print(f"n {n:5d}: 8 teams/CTA {res[0]:.2f} ms   auto {res[1]:.2f} ms", flush=True)
