"""Mid-size levels (between one wave of the low-latency kernel and a full persistent wave): which launch shape is
fastest?  python scripts/prof_mid.py [counts...]"""
import sys; sys.path.insert(0, '.')
import numpy as np
import peba1_b200 as P
from peba1_b200.cuda import make_gates
counts = [int(v) for v in sys.argv[1:]] or [296, 444, 592, 740, 888, 1036, 1183]
eng = P.Engine()
rng = np.random.default_rng(1)
p = eng.params
bk = rng.integers(-2**31, 2**31, size=p.n*6*2*1024, dtype=np.int64).astype(np.int32)
ksk = rng.integers(-2**31, 2**31, size=1024*8*4*631, dtype=np.int64).astype(np.int32)
eng.load_keys(bk, ksk)
nmax = max(counts)
cts = rng.integers(-2**31, 2**31, size=(2*nmax, 631), dtype=np.int64).astype(np.int32)
eng.reserve(3*nmax); eng.upload(0, cts)
eng.enable_timing(True)
def run(g):
    for rep in range(2):
        eng.gates(g)
    return eng.stats().last_blind_rotate_ms
for n in counts:
    g = make_gates("NAND", np.arange(n), nmax+np.arange(n), 2*nmax+np.arange(n))
    out = []
    eng.set_option("kernel_variant", 1); eng.set_option("wide_max_batch", 0)
    for t in (2, 4, 8):
        eng.set_option("tm_teams", t); out.append(f"T={t}: {run(g):6.2f}")
    eng.set_option("tm_teams", 0); out.append(f"auto: {run(g):6.2f}")
    eng.set_option("wide_max_batch", 1 << 30); out.append(f"wide: {run(g):6.2f}")
    eng.set_option("wide_max_batch", 0)
    eng.set_option("kernel_variant", 7); eng.set_option("pair_min_batch", 1); out.append(f"persistent8: {run(g):6.2f}")
    eng.set_option("kernel_variant", 6); out.append(f"persistent10: {run(g):6.2f}")
    eng.set_option("kernel_variant", 1); eng.set_option("pair_min_batch", -1)
    print(f"n {n:5d} ms: " + "  ".join(out), flush=True)
