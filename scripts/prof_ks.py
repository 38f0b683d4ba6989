"""Times the keyswitch kernel alone for each ks_group: python scripts/prof_ks.py [batch]"""
import sys; sys.path.insert(0, '.')
import numpy as np
import peba1_b200 as P
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
groups = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 2, 4, 6]
eng = P.Engine()
rng = np.random.default_rng(1)
p = eng.params
ksk = rng.integers(-2**31, 2**31, size=1024*8*4*631, dtype=np.int64).astype(np.int32)
bk = np.zeros(p.n*6*2*1024, dtype=np.int32)
eng.load_keys(bk, ksk)
u = rng.integers(-2**31, 2**31, size=(n, 1025), dtype=np.int64).astype(np.int32)
eng.enable_timing(True)
ref = None
for grp in groups:
    eng.set_option("ks_group", grp)
    for rep in range(2):
        out = eng.keyswitch_host(u)
    s = eng.stats()
    same = "" if ref is None else (" same-as-first" if np.array_equal(out, ref) else " MISMATCH")
    ref = out if ref is None else ref
    print(f"ks_group {grp} n {n}: keyswitch {s.last_keyswitch_ms:.3f} ms{same}", flush=True)
