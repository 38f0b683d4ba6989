"""Tiny end-to-end run for compute-sanitizer: every kernel variant, a MUX batch, linear ops, upload/download."""
import sys; sys.path.insert(0, '.')
import numpy as np
import peba1_b200 as P
eng = P.Engine()
rng = np.random.default_rng(1)
p = eng.params
bk = rng.integers(-2**31, 2**31, size=p.n*6*2*1024, dtype=np.int64).astype(np.int32)
ksk = rng.integers(-2**31, 2**31, size=1024*8*4*631, dtype=np.int64).astype(np.int32)
eng.load_keys(bk, ksk)
n = 11
ca, cb, cc = (rng.integers(-2**31, 2**31, size=(n, 631), dtype=np.int64).astype(np.int32) for _ in range(3))
ref = None
for variant in (1, 2, 0):
    eng.set_option("kernel_variant", variant)
    out = eng.gate_batch_host("MUX", ca, cb, cc)
    ref = out if ref is None else ref
    assert np.array_equal(out, ref)
eng.set_option("kernel_variant", 1)
eng.linear([0, -1], [-1, 0], [0, 1 << 29], [3, 4])
eng.download(0, 5)
print("sanitize_small ok")
