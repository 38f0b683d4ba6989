"""World-size-2 CPU test of the multi-GPU path's host logic (peba1_b200/gallery.py) on the gloo backend:
rank 0 encrypts, probe/bound are broadcast, templates scattered, each rank evaluates its shard of Function_f on the
bit-level simulator, match bits are all-gathered and rank 0 checks them against the plaintext distances."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["PEBA1_ROOT"]); sys.path.insert(0, os.path.join(os.environ["PEBA1_ROOT"], "tests"))
import numpy as np, torch, torch.distributed as dist
from peba1_b200 import tfhe_host as H, gallery as G
from sim_backend import SimBackend
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
sim = SimBackend(); sim.install()
ks = H.KeySet.generate(128, seed=42)          # the simulator decrypts uploads, so every rank holds the secret key
sim.keyset = ks
slots, bits, T, bound = 6, 8, 3, 40
probe, gal, expect = G.synthetic_gallery(world * T, slots, seed=5, planted=0.5, bound=bound)
client = None
if rank == 0:
    client = (ks.encrypt(G.bits_of(probe, bits), 1), ks.encrypt(G.bits_of([bound], 24), 2),
              ks.encrypt(G.bits_of(gal.reshape(-1), bits), 3))
pw, bw, mw = G.distribute(dist, rank, world, "cpu", torch, ks.words, slots, bits, T, client)
out = G.match_templates(ks, pw, bw, mw, slots, bits)
allout = G.gather_match_bits(dist, world, "cpu", torch, out)
assert allout.shape == (world * T, ks.words)
got = ks.decrypt(allout)
assert np.array_equal(got, expect), (rank, got, expect)
lo, hi = G.shard_bounds(world * T, world, rank)
assert hi - lo == T
if rank == 0:
    assert 0 < expect.sum() < world * T        # both outcomes present
    print("GLOO_OK", got.tolist())
dist.destroy_process_group()
'''


def test_shard_bounds():
    from peba1_b200 import gallery as G
    for total, world in [(10, 4), (8, 8), (3, 8), (65536, 8)]:
        cover = []
        for r in range(world):
            lo, hi = G.shard_bounds(total, world, r)
            cover += list(range(lo, hi))
        assert cover == list(range(total))


def test_world_size_2_gloo(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   PEBA1_ROOT=ROOT, OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "GLOO_OK" in outs[0]
