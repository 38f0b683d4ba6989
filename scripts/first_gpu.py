import sys, time; sys.path.insert(0, '.')
import numpy as np
import peba1_b200 as P
eng = P.Engine()
print("fp64 peak TF:", eng.measure_fp64_peak())
rng = np.random.default_rng(1)
p = eng.params
bk = rng.integers(-2**31, 2**31, size=p.n*6*2*1024, dtype=np.int64).astype(np.int32)
ksk = rng.integers(-2**31, 2**31, size=1024*8*4*631, dtype=np.int64).astype(np.int32)
t=time.time(); eng.load_keys(bk, ksk); print("load keys s", time.time()-t)
n = 16384
cts = rng.integers(-2**31, 2**31, size=(2*n, 631), dtype=np.int64).astype(np.int32)
eng.reserve(3*n); eng.upload(0, cts)
eng.enable_timing(True)
from peba1_b200.cuda import make_gates
g = make_gates("NAND", np.arange(n), n+np.arange(n), 2*n+np.arange(n))
for teams in (4, 5, 6):
    eng.set_option("teams_per_cta", teams)
    for rep in range(2):
        eng.gates(g)
        s = eng.stats()
        print(f"teams {teams}: blind rotate {s.last_blind_rotate_ms:.2f} ms ({n/s.last_blind_rotate_ms*1e3:.0f} PBS/s), keyswitch {s.last_keyswitch_ms:.2f} ms", flush=True)
