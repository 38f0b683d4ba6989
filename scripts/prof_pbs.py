"""Small launch of the blind-rotation kernel for ncu: python scripts/prof_pbs.py [teams] [batch]"""
import sys; sys.path.insert(0, '.')
import numpy as np
import peba1_b200 as P
from peba1_b200.cuda import make_gates
teams = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * teams * 2
eng = P.Engine()
rng = np.random.default_rng(1)
p = eng.params
bk = rng.integers(-2**31, 2**31, size=p.n*6*2*1024, dtype=np.int64).astype(np.int32)
ksk = rng.integers(-2**31, 2**31, size=1024*8*4*631, dtype=np.int64).astype(np.int32)
eng.load_keys(bk, ksk)
cts = rng.integers(-2**31, 2**31, size=(2*n, 631), dtype=np.int64).astype(np.int32)
eng.reserve(3*n); eng.upload(0, cts)
eng.enable_timing(True)
eng.set_option("teams_per_cta", teams if teams <= 6 else 6)
if teams >= 8:
    eng.set_option("kernel_variant", 1 if teams == 8 else 2)
g = make_gates("NAND", np.arange(n), n+np.arange(n), 2*n+np.arange(n))
for rep in range(3):
    eng.gates(g)
    s = eng.stats()
    print(f"teams {teams} n {n}: blind rotate {s.last_blind_rotate_ms:.2f} ms ({n/s.last_blind_rotate_ms*1e3:.0f} PBS/s), keyswitch {s.last_keyswitch_ms:.2f} ms", flush=True)
