"""variant 4 with different start-up staggers: python scripts/prof_stagger.py [batch] [ns...]"""
import sys; sys.path.insert(0, '.')
import numpy as np
import peba1_b200 as P
from peba1_b200.cuda import make_gates
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
vals = [int(v) for v in sys.argv[2:]] or [0, 100, 200, 400, 800]
eng = P.Engine()
rng = np.random.default_rng(1)
p = eng.params
bk = rng.integers(-2**31, 2**31, size=p.n*6*2*1024, dtype=np.int64).astype(np.int32)
ksk = rng.integers(-2**31, 2**31, size=1024*8*4*631, dtype=np.int64).astype(np.int32)
eng.load_keys(bk, ksk)
cts = rng.integers(-2**31, 2**31, size=(2*n, 631), dtype=np.int64).astype(np.int32)
eng.reserve(3*n); eng.upload(0, cts)
eng.enable_timing(True)
g = make_gates("NAND", np.arange(n), n+np.arange(n), 2*n+np.arange(n))
eng.set_option("kernel_variant", 4)
for v in vals:
    eng.set_option("tmk_stagger_ns", v)
    for rep in range(2):
        eng.gates(g)
        s = eng.stats()
    print(f"stagger {v} ns/team: blind rotate {s.last_blind_rotate_ms:.2f} ms ({n/s.last_blind_rotate_ms*1e3:.0f} PBS/s)", flush=True)
