"""One narrow circuit level end to end: blind rotation, keyswitch and wall time per peba1_cuda_gates call (+ sync)."""
import sys, time; sys.path.insert(0, '.')
import numpy as np
import peba1_b200 as P
from peba1_b200.cuda import make_gates
counts = [int(v) for v in sys.argv[1:]] or [1, 7, 81, 148, 296, 592]
eng = P.Engine()
rng = np.random.default_rng(1)
p = eng.params
bk = rng.integers(-2**31, 2**31, size=p.n*6*2*1024, dtype=np.int64).astype(np.int32)
ksk = rng.integers(-2**31, 2**31, size=1024*8*4*631, dtype=np.int64).astype(np.int32)
eng.load_keys(bk, ksk)
nmax = max(counts)
cts = rng.integers(-2**31, 2**31, size=(2*nmax, 631), dtype=np.int64).astype(np.int32)
eng.reserve(3*nmax); eng.upload(0, cts)
for n in counts:
    g = make_gates("NAND", np.arange(n), nmax+np.arange(n), 2*nmax+np.arange(n))
    eng.enable_timing(True)
    for rep in range(3):
        eng.gates(g)
    s = eng.stats()
    eng.enable_timing(False)
    eng.gates(g); eng.sync()
    t0 = time.perf_counter()
    for rep in range(10):
        eng.gates(g)
    eng.sync()
    wall = (time.perf_counter() - t0) / 10 * 1e3
    print(f"n {n:4d}: blind rotate {s.last_blind_rotate_ms:6.2f} ms, keyswitch {s.last_keyswitch_ms:6.3f} ms, wall per level {wall:6.2f} ms", flush=True)
