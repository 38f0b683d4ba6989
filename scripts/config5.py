"""BASELINE configs[4]: 1:N identification over a synthetic gallery sharded across the GPUs of one box.
   torchrun --nproc-per-node N scripts/config5.py --gallery 65536 --subset 1024 --mode 2

The gallery is SYNTHETIC and generated where it is used: every rank derives its shard's plaintext templates from the
seed and encrypts them on its own GPU (peba1_encrypt_array_device) wave by wave, so no gallery
ciphertext (169 GB for 65,536 templates) crosses PCIe or NVLink.  Only rank 0's client material travels: the device-
format keys, the LWE key for the generator (benchmark only), the encrypted probe and bound.  The full gallery is
1.4e10 bootstraps (hours); `--subset` templates are timed (spread evenly over the ranks, each rank taking the first
ones of ITS shard of the full gallery) and the full-gallery time is extrapolated, as SURVEY.md §8d prescribes.
Timed: recording + all circuit levels of every wave + the NCCL all-gather of the encrypted match bits.  Not timed:
gallery generation.  Prints one JSON line (rank 0)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--gallery", type=int, default=65536)
ap.add_argument("--subset", type=int, default=256, help="templates actually matched (all ranks together)")
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
dev = f"cuda:{local}"
# every rank needs the LWE key for the on-device generator: all ranks generate the (deterministic) key set
ks = H.KeySet.generate(128, seed=42)
eng = P.Engine.from_handle(ks.cuda_context(), P.default_params(128), device=local)
slots, bits, bound = 128, 8, 512
lo, hi = G.shard_bounds(args.gallery, world, rank)           # this rank's shard of the FULL gallery
T = min(args.subset // world, hi - lo)                       # templates of it that are timed
probe, gal_local, expect_local = G.synthetic_gallery_shard(args.gallery, lo, lo + T, slots, bound=bound)
probe_w = ks.encrypt(G.bits_of(probe, bits), 501)            # the client's probe and bound: same seed on every rank
bound_w = ks.encrypt(G.bits_of([bound], 24), 502)
L.peba1_set_mode(args.mode)

def barrier():
    if dist is not None: dist.barrier()
    torch.cuda.synchronize()

s0 = ks.stats()
t_gen = 0.0
t_run = 0.0
outs = []
for w0 in range(0, T, args.wave):
    w1 = min(T, w0 + args.wave)
    tg = time.perf_counter()
    tmpl = G.encrypt_templates_on_device(ks, gal_local[w0:w1], bits, seed=7000 + lo + w0)
    torch.cuda.synchronize()
    t_gen += time.perf_counter() - tg
    barrier()
    tr = time.perf_counter()
    outs.append(G.match_template_arrays(ks, probe_w, bound_w, tmpl, slots, bits))
    torch.cuda.synchronize()
    t_run += time.perf_counter() - tr
barrier()
tr = time.perf_counter()
allout = G.gather_match_bits(dist, world, dev, torch, np.concatenate(outs))
barrier()
t_run += time.perf_counter() - tr
tt = torch.tensor([t_run, t_gen], dtype=torch.float64, device=dev)
if dist is not None: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
t_run, t_gen = (float(x) for x in tt.tolist())
s1 = ks.stats()
ex = torch.from_numpy(expect_local).to(dev)
if dist is not None:
    parts = [torch.empty_like(ex) for _ in range(world)]
    dist.all_gather(parts, ex)
    ex = torch.cat(parts)
if rank == 0:
    got = ks.decrypt(allout)
    rate = world * T / t_run
    print(json.dumps({
        "workload": f"1:N identification, synthetic gallery of {args.gallery} templates (128-dim, 8-bit, 24-bit threshold) "
                    f"sharded over {world} GPU(s); timed subset {world * T} templates ({T} per GPU, waves of {args.wave})",
        "circuit_mode": args.mode, "face_matches_per_s": rate, "seconds": t_run,
        "extrapolated_full_gallery_seconds": args.gallery / rate,
        "pbs_per_match": (s1.executed_pbs - s0.executed_pbs) / max(T, 1),
        "pbs_per_s": (s1.executed_pbs - s0.executed_pbs) * world / t_run,
        "gallery_generation": {"where": "on each GPU (peba1_cuda_encrypt_bits)", "seconds": t_gen,
                               "ciphertexts_per_gpu": int(s1.device_encryptions - s0.device_encryptions),
                               "ciphertexts_per_s_per_gpu": (s1.device_encryptions - s0.device_encryptions) / max(t_gen, 1e-9)},
        "gathered_bytes": int(allout.size * 4), "verified": bool(np.array_equal(got, ex.cpu().numpy())),
        "matches_found": int((got == 0).sum()), "peak_live_ciphertexts_per_gpu": int(s1.peak_slots)}), flush=True)
if dist is not None: dist.destroy_process_group()
