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
# BASELINE configs[4] plumbing (scripts/config5.py): no scatter at all — every rank derives ITS shard of the synthetic
# gallery from the seed, encrypts it where it is used (peba1_encrypt_array_list_device; on the simulator the words are
# made on the host), matches it in circuit mode 3, and only the encrypted match bits travel
total = 7
lo, hi = G.shard_bounds(total, world, rank)
probe5, mine5, expect5 = G.synthetic_gallery_shard(total, lo, hi, slots, seed=9, planted=0.5, bound=bound)
H.load_library().peba1_set_mode(3)
tv = G.encrypt_templates_on_device(ks, mine5, bits, seed=1000 + lo)
out5 = G.match_template_arrays(ks, ks.encrypt(G.bits_of(probe5, bits), 1), ks.encrypt(G.bits_of([bound], 24), 2), tv,
                               slots, bits)
H.load_library().peba1_set_mode(0)
sizes = [G.shard_bounds(total, world, r)[1] - G.shard_bounds(total, world, r)[0] for r in range(world)]
pad = np.zeros((max(sizes) - out5.shape[0], ks.words), dtype=np.int32)          # all_gather wants equal shapes
all5 = G.gather_match_bits(dist, world, "cpu", torch, np.concatenate([out5, pad]))
got5 = np.concatenate([ks.decrypt(all5[r * max(sizes): r * max(sizes) + sizes[r]]) for r in range(world)])
whole = G.synthetic_gallery_shard(total, 0, total, slots, seed=9, planted=0.5, bound=bound)[2]
assert np.array_equal(got5, whole), (rank, got5, whole)
if rank == 0:
    assert 0 < expect.sum() < world * T        # both outcomes present
    assert 0 < whole.sum() < total
    print("GLOO_OK", got.tolist(), got5.tolist())
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
