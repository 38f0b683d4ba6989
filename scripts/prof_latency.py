"""Narrow levels (wide_max_batch 0 = never the low-latency kernel, 2^30 = always): time per launch of the blind rotation for small batches, low-latency kernel (one CTA per bootstrap)
against the throughput kernel's narrow shapes; outputs must be identical.  python scripts/prof_latency.py [counts...]"""
import sys; sys.path.insert(0, '.')
import numpy as np
import peba1_b200 as P
from peba1_b200.cuda import make_gates
counts = [int(v) for v in sys.argv[1:]] or [1, 7, 37, 81, 148, 222, 296, 444, 592, 888, 1184]
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
for n in counts:
    g = make_gates("NAND", np.arange(n), nmax+np.arange(n), 2*nmax+np.arange(n))
    res = {}
    for name, wide in (("throughput", 0), ("wide", 1 << 30)):
        eng.set_option("wide_max_batch", wide)
        for rep in range(2):
            eng.gates(g)
            s = eng.stats()
        res[name] = (s.last_blind_rotate_ms, eng.download(2*nmax, n))
    same = bool(np.array_equal(res["throughput"][1], res["wide"][1]))
    print(f"n {n:5d}: throughput-kernel shapes {res['throughput'][0]:7.2f} ms, wide {res['wide'][0]:7.2f} ms, identical {same}", flush=True)
