"""Back-to-back asynchronous peba1_cuda_gates calls: every call must do its work (distinct output slots checked against
a synchronous run) and take its time (CUDA events between the calls)."""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
import peba1_b200 as P
from peba1_b200.cuda import make_gates
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
eng = P.Engine()
stream = torch.cuda.Stream()
eng.set_stream(stream.cuda_stream)
rng = np.random.default_rng(1)
p = eng.params
bk = rng.integers(-2**31, 2**31, size=p.n*6*2*1024, dtype=np.int64).astype(np.int32)
ksk = rng.integers(-2**31, 2**31, size=1024*8*4*631, dtype=np.int64).astype(np.int32)
eng.load_keys(bk, ksk)
K = 4
cts = rng.integers(-2**31, 2**31, size=((K + 1) * n, 631), dtype=np.int64).astype(np.int32)
eng.reserve((2 * K + 2) * n); eng.upload(0, cts)
gs = [make_gates("NAND", k * n + np.arange(n), (k + 1) * n + np.arange(n), (K + 1 + k) * n + np.arange(n)) for k in range(K)]
eng.gates(gs[0]); torch.cuda.synchronize()
ref = []
for k in range(K):                      # synchronous reference
    eng.gates(gs[k]); torch.cuda.synchronize()
    ref.append(eng.download((K + 1 + k) * n, n))
eng.upload((K + 1) * n, np.zeros((K * n, 631), dtype=np.int32))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
ev[0].record(stream)
for k in range(K):
    eng.gates(gs[k]); ev[k + 1].record(stream)
torch.cuda.synchronize()
for k in range(K):
    ok = np.array_equal(eng.download((K + 1 + k) * n, n), ref[k])
    print(f"call {k}: {ev[k].elapsed_time(ev[k + 1]):.2f} ms, output identical to the synchronous run: {ok}", flush=True)
