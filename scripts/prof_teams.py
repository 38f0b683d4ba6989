"""Throughput of the tensor-memory kernel vs teams per CTA (= warps per SM): python scripts/prof_teams.py"""
import sys; sys.path.insert(0, '.')
import numpy as np
import peba1_b200 as P
from peba1_b200.cuda import make_gates
eng = P.Engine()
rng = np.random.default_rng(1)
p = eng.params
eng.load_keys(rng.integers(-2**31, 2**31, size=p.n*6*2*1024, dtype=np.int64).astype(np.int32),
              rng.integers(-2**31, 2**31, size=1024*8*4*631, dtype=np.int64).astype(np.int32))
eng.enable_timing(True)
for teams in (1, 2, 4, 8):
    n = 148 * teams * 4
    eng.reserve(3*n); eng.upload(0, rng.integers(-2**31, 2**31, size=(2*n, 631), dtype=np.int64).astype(np.int32))
    g = make_gates("NAND", np.arange(n), n+np.arange(n), 2*n+np.arange(n))
    eng.set_option("tm_teams", teams)
    for rep in range(2):
        eng.gates(g)
    ms = eng.stats().last_blind_rotate_ms
    print(f"teams/CTA {teams} ({2*teams} warps/SM): {n} PBS in {ms:.2f} ms -> {n/ms*1e3:.0f} PBS/s, {ms/4:.2f} ms per wave", flush=True)
