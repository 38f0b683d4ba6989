"""Every blind-rotation kernel variant on a SHORT key (n = 8 mask words: 8 CMux steps instead of 630) for
compute-sanitizer: a few CTAs per variant, persistent and non-persistent launches, odd counts (shadow jobs), the
keyswitch, device-side encryption.  All variants must agree word for word."""
import sys; sys.path.insert(0, '.')
import numpy as np
import peba1_b200 as P
from peba1_b200.cuda import Params
p = Params(8, 1024, 1, 3, 7, 8, 2)
eng = P.Engine(p)
rng = np.random.default_rng(1)
bk = rng.integers(-2**31, 2**31, size=p.n*6*2*1024, dtype=np.int64).astype(np.int32)
ksk = rng.integers(-2**31, 2**31, size=1024*8*4*(p.n+1), dtype=np.int64).astype(np.int32)
eng.load_keys(bk, ksk)
n = 37
ca, cb, cc = (rng.integers(-2**31, 2**31, size=(n, p.n + 1), dtype=np.int64).astype(np.int32) for _ in range(3))
eng.set_option("pair_min_batch", 1)
ref = None
for variant in (1, 0, 2, 3, 4, 5, 6, 7, 8):
    eng.set_option("kernel_variant", variant)
    out = eng.gate_batch_host("MUX", ca, cb, cc)
    ref = out if ref is None else ref
    assert np.array_equal(out, ref), variant
    print("variant", variant, "ok", flush=True)
eng.set_option("kernel_variant", 1)
eng.reserve(64)
eng.encrypt_bits(rng.integers(0, 2, size=p.n), rng.integers(0, 2, size=20), 3, seed=5)
eng.download(0, 30)
print("sanitize_tiny ok")
