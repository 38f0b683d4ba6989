"""One narrow launch through the low-latency kernel, for ncu: python scripts/prof_wide.py [count]"""
import sys; sys.path.insert(0, '.')
import numpy as np
import peba1_b200 as P
from peba1_b200.cuda import make_gates
n = int(sys.argv[1]) if len(sys.argv) > 1 else 81
eng = P.Engine()
rng = np.random.default_rng(1)
p = eng.params
bk = rng.integers(-2**31, 2**31, size=p.n*6*2*1024, dtype=np.int64).astype(np.int32)
ksk = rng.integers(-2**31, 2**31, size=1024*8*4*631, dtype=np.int64).astype(np.int32)
eng.load_keys(bk, ksk)
cts = rng.integers(-2**31, 2**31, size=(2*n, 631), dtype=np.int64).astype(np.int32)
eng.reserve(3*n); eng.upload(0, cts)
eng.enable_timing(True)
eng.set_option("wide_max_batch", 1 << 30)
g = make_gates("NAND", np.arange(n), n+np.arange(n), 2*n+np.arange(n))
for rep in range(3):
    eng.gates(g)
    print(f"n {n}: blind rotate {eng.stats().last_blind_rotate_ms:.2f} ms", flush=True)
