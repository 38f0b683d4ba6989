"""Blind-rotation kernel variants side by side on one batch: python scripts/prof_variants.py [batch] [variants...]
Random keys (timing only); prints ms and PBS/s per variant, and checks that all variants give identical words."""
import sys; sys.path.insert(0, '.')
import numpy as np
import peba1_b200 as P
from peba1_b200.cuda import make_gates
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
variants = [int(v) for v in sys.argv[2:]] or [1, 3]
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
ref = None
for v in variants:
    eng.set_option("kernel_variant", v)
    for rep in range(3):
        eng.gates(g)
        s = eng.stats()
        print(f"variant {v} n {n}: blind rotate {s.last_blind_rotate_ms:.2f} ms ({n/s.last_blind_rotate_ms*1e3:.0f} PBS/s), keyswitch {s.last_keyswitch_ms:.2f} ms", flush=True)
    out = eng.download(2*n, n)
    if ref is None: ref = out
    else: print(f"variant {v} identical to variant {variants[0]}:", bool(np.array_equal(out, ref)), flush=True)
