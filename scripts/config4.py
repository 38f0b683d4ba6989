"""BASELINE configs[3]: 1 probe vs a gallery of templates sharded over the GPUs of one box (NCCL scatter / all-gather).
   torchrun --nproc-per-node N scripts/config4.py --templates 1024 --mode 2
Prints one JSON line (rank 0).  Timed: recording + all circuit levels + gather of the encrypted match bits."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--templates", type=int, default=1024)
ap.add_argument("--mode", type=int, default=2)
ap.add_argument("--wave", type=int, default=64, help="templates recorded per batch on each GPU")
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import peba1_b200 as P
from peba1_b200 import tfhe_host as H, gallery as G
L = H.load_library()
L.peba1_host_set_device(local)
ks = H.KeySet.generate(128, seed=42) if rank == 0 else H.KeySet.placeholder(128)
eng = P.Engine.from_handle(ks.cuda_context(), P.default_params(128), device=local)
if dist is not None:
    (pa, na), (pb, nb) = eng.key_device_buffers()
    class _D:
        def __init__(s, p, n): s.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (p, False), "version": 2}
    for p_, n_ in ((pa, na), (pb, nb)):
        dist.broadcast(torch.as_tensor(_D(p_, n_), device=f"cuda:{local}"), 0)
    torch.cuda.synchronize()
    eng.keys_mark_loaded()
T = args.templates // world
slots, bits, bound, dev = 128, 8, 512, f"cuda:{local}"
probe, gal, expect = G.synthetic_gallery(world * T, slots, bound=bound)
client = None
t_enc = time.perf_counter()
if rank == 0:
    client = (ks.encrypt(G.bits_of(probe, bits), 501), ks.encrypt(G.bits_of([bound], 24), 502),
              ks.encrypt(G.bits_of(gal.reshape(-1), bits), 503))
t_enc = time.perf_counter() - t_enc
pw, bw, mw = G.distribute(dist, rank, world, dev, torch, ks.words, slots, bits, T, client)
L.peba1_set_mode(args.mode)
def barrier():
    if dist is not None: dist.barrier()
    torch.cuda.synchronize()
s0 = ks.stats()
barrier()
t0 = time.perf_counter()
outs = []
for w0 in range(0, T, args.wave):                      # waves bound the recorder's memory and the live ciphertexts
    outs.append(G.match_templates(ks, pw, bw, mw[w0:w0 + args.wave], slots, bits))
allout = G.gather_match_bits(dist, world, dev, torch, np.concatenate(outs))
barrier()
dt = time.perf_counter() - t0
tt = torch.tensor([dt], dtype=torch.float64, device=dev)
if dist is not None: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
s1 = ks.stats()
if rank == 0:
    got = ks.decrypt(allout)
    print(json.dumps({"workload": f"1 probe vs {world * T} templates (128-dim, 8-bit, 24-bit threshold), {world} GPU(s), {T} per GPU",
                      "circuit_mode": args.mode, "face_matches_per_s": world * T / float(tt.item()), "seconds": float(tt.item()),
                      "pbs_per_match": (s1.executed_pbs - s0.executed_pbs) / T, "pbs_per_s": (s1.executed_pbs - s0.executed_pbs) * world / float(tt.item()),
                      "levels": int(s1.levels - s0.levels), "verified": bool(np.array_equal(got, expect)),
                      "matches_found": int((got == 0).sum()), "client_encrypt_s": t_enc, "peak_live_ciphertexts_per_gpu": int(s1.peak_slots)}), flush=True)
if dist is not None: dist.destroy_process_group()
