"""1:N encrypted identification over a sharded gallery (BASELINE configs[3] and [4]).

One process per GPU.  Rank 0 plays the client: it encrypts the probe, the 24-bit bound and the gallery.  The probe
and the bound are broadcast, template ciphertexts are scattered (each rank keeps `T` templates), every rank records
`Function_f` for all of its templates into ONE batch (so each of the 377 circuit levels is one launch pair), and the
encrypted match bits (one 2,524-byte ciphertext per template) are all-gathered.  No collective sits inside the hot
loop.  Works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests) or without one (world = 1).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import tfhe_host as H


def shard_bounds(total: int, world: int, rank: int):
    """Contiguous shard [lo, hi) of `total` templates for `rank` (sizes differ by at most one)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def bits_of(values, nbits: int) -> np.ndarray:
    v = np.asarray(values, dtype=np.int64).reshape(-1, 1)
    return ((v >> np.arange(nbits)) & 1).reshape(-1).astype(np.int32)


def synthetic_gallery(total: int, slots: int = 128, seed: int = 2024, planted: float = 0.25, bound: int = 512):
    """Probe, gallery and expected comparator bits.  The probe has no zero byte, so the reference's subtraction defect
    (wrong when the subtrahend is 0) stays silent; a fraction of the gallery are near-copies of the
    probe.  Polarity (Q3): the bit is 1 iff squared distance > bound."""
    rng = np.random.default_rng(seed)
    probe = rng.integers(1, 256, size=slots)
    gallery = rng.integers(0, 256, size=(total, slots))
    mask = (np.arange(total) % max(1, int(round(1 / planted)))) == 1 if planted > 0 else np.zeros(total, bool)
    gallery[mask] = np.clip(probe + rng.integers(-2, 3, size=(int(mask.sum()), slots)), 0, 255)
    dist2 = ((gallery.astype(np.int64) - probe) ** 2).sum(axis=1)
    return probe, gallery, (dist2 > bound).astype(np.int64)


def synthetic_gallery_shard(total: int, lo: int, hi: int, slots: int = 128, seed: int = 2024, planted: float = 0.25,
                            bound: int = 512):
    """Templates [lo, hi) of a `total`-template gallery whose template t depends on (seed, t) only, so each rank
    makes its own shard (BASELINE configs[4]: 65,536 templates) without anyone holding the whole gallery.  Same
    planting rule and polarity as synthetic_gallery.  Returns (probe, templates [hi-lo, slots], expected bits)."""
    probe = np.random.default_rng([seed, 0xFACE]).integers(1, 256, size=slots)
    period = max(1, int(round(1 / planted))) if planted > 0 else 0
    gallery = np.empty((hi - lo, slots), dtype=np.int64)
    for t in range(lo, hi):
        rng = np.random.default_rng([seed, t])
        if period and t % period == 1:
            gallery[t - lo] = np.clip(probe + rng.integers(-2, 3, size=slots), 0, 255)
        else:
            gallery[t - lo] = rng.integers(0, 256, size=slots)
    dist2 = ((gallery - probe) ** 2).sum(axis=1)
    return probe, gallery, (dist2 > bound).astype(np.int64)


def distribute(dist, rank: int, world: int, device, torch, words: int, slots: int, bits: int, per_rank: int,
               client=None):
    """client = (probe_words [slots*bits, words], bound_words [24, words], gallery_words [world*per_rank, slots*bits,
    words]) on rank 0, None elsewhere.  Returns this rank's (probe, bound, templates) as numpy arrays."""
    probe_t = torch.empty((slots * bits, words), dtype=torch.int32, device=device)
    bound_t = torch.empty((24, words), dtype=torch.int32, device=device)
    mine_t = torch.empty((per_rank, slots * bits, words), dtype=torch.int32, device=device)
    chunks = None
    if rank == 0:
        pw, bw, gw = client
        probe_t.copy_(torch.from_numpy(np.ascontiguousarray(pw)))
        bound_t.copy_(torch.from_numpy(np.ascontiguousarray(bw)))
        gw = np.ascontiguousarray(gw).reshape(world, per_rank, slots * bits, words)
        chunks = [torch.from_numpy(gw[r]).to(device) for r in range(world)]
    if dist is not None and world > 1:
        dist.broadcast(probe_t, 0)
        dist.broadcast(bound_t, 0)
        dist.scatter(mine_t, chunks, src=0)
    else:
        mine_t.copy_(chunks[0])
    return probe_t.cpu().numpy(), bound_t.cpu().numpy(), mine_t.cpu().numpy()


def distribute_device(dist, rank: int, world: int, device, torch, words: int, slots: int, bits: int, per_rank: int,
                      client=None):
    """As distribute, but what the collectives deliver STAYS on the device: returns torch int32 tensors
    (probe [slots*bits, words], bound [24, words], templates [per_rank, slots*bits, words]) for device_vectors()."""
    probe_t = torch.empty((slots * bits, words), dtype=torch.int32, device=device)
    bound_t = torch.empty((24, words), dtype=torch.int32, device=device)
    mine_t = torch.empty((per_rank, slots * bits, words), dtype=torch.int32, device=device)
    chunks = None
    if rank == 0:
        pw, bw, gw = client
        probe_t.copy_(torch.from_numpy(np.ascontiguousarray(pw)))
        bound_t.copy_(torch.from_numpy(np.ascontiguousarray(bw)))
        gw = np.ascontiguousarray(gw).reshape(world, per_rank, slots * bits, words)
        chunks = [torch.from_numpy(gw[r]).to(device) for r in range(world)]
    if dist is not None and world > 1:
        dist.broadcast(probe_t, 0)
        dist.broadcast(bound_t, 0)
        dist.scatter(mine_t, chunks, src=0)
    else:
        mine_t.copy_(chunks[0])
    torch.cuda.synchronize()
    return probe_t, bound_t, mine_t


def device_vectors(ks, tensor, count: int, slots: int, bits: int):
    """`count` vectors of `slots` ciphertext arrays of `bits` samples bound to the rows of a device tensor
    [count * slots * bits, words] by ONE device-side scatter into the ciphertext pool (peba1_sample_set_list_device)."""
    LSP = C.POINTER(H.LweSample)
    flat, out = [], []
    for _ in range(count):
        arrs = [ks.new_array(bits) for _ in range(slots)]
        flat.extend(arrs)
        out.append((arrs, (LSP * slots)(*arrs)))
    ks.set_list_device(flat, bits, tensor.data_ptr())
    return out


def _host_vector(ks, w, slots: int, bits: int):
    LSP = C.POINTER(H.LweSample)
    arrs = []
    for i in range(slots):
        a = ks.new_array(bits)
        ks.set_words(a, w[i * bits:(i + 1) * bits])
        arrs.append(a)
    return arrs, (LSP * slots)(*arrs)


def encrypt_templates_on_device(ks: "H.KeySet", templates: np.ndarray, bits: int = 8, seed: int = 1):
    """Plaintext templates [T, slots] -> per template a vector of `slots` ciphertext arrays whose encryptions are made
    on the GPU (peba1_encrypt_array_device): nothing but the plaintext bits crosses PCIe.  Benchmark-side stand-in for
    the clients who would upload their encrypted templates."""
    LSP = C.POINTER(H.LweSample)
    out = []
    T, slots = templates.shape
    flat = []
    for t in range(T):
        arrs = [ks.new_array(bits) for _ in range(slots)]
        flat.extend(arrs)
        out.append((arrs, (LSP * slots)(*arrs)))
    if flat:
        ks.encrypt_array_list_device(flat, bits, bits_of(templates.reshape(-1), bits), seed)  # one launch per wave
    return out


def match_template_arrays(ks: "H.KeySet", probe_w, bound_w, template_vecs, slots: int = 128, bits: int = 8) -> np.ndarray:
    """Function_f(probe, template_t, bound) for every template vector in `template_vecs` (as made by _host_vector or
    encrypt_templates_on_device), ONE batch: each of the 377 circuit levels is one launch pair.  Frees the template
    arrays; returns the comparator-bit ciphertexts [T, words]."""
    L = H.load_library()
    pa, pp = _host_vector(ks, probe_w, slots, bits)
    eb = ks.new_array(24)
    ks.set_words(eb, bound_w)
    T = len(template_vecs)
    rb = [ks.new_array(24) for _ in range(T)]
    L.peba1_batch_begin()
    for t in range(T):
        L.peba1_math_function_f(rb[t], pp, template_vecs[t][1], slots, eb, bits, ks.cloud)
    L.peba1_batch_end()
    out = np.stack([ks.get_words(rb[t], 1)[0] for t in range(T)]) if T else np.zeros((0, ks.words), np.int32)
    for arrs, _ in list(template_vecs) + [(pa, pp)]:
        for a in arrs:
            ks.free_array(bits, a)
    for a in rb + [eb]:
        ks.free_array(24, a)
    return out


def match_templates(ks: "H.KeySet", probe_w, bound_w, templates_w, slots: int = 128, bits: int = 8) -> np.ndarray:
    """As match_template_arrays, from host ciphertext words [T, slots*bits, words] (uploaded)."""
    return match_template_arrays(ks, probe_w, bound_w,
                                 [_host_vector(ks, templates_w[t], slots, bits) for t in range(templates_w.shape[0])],
                                 slots, bits)


def authenticate_batch(ks: "H.KeySet", probe_vecs, template_vecs, bound_w, r0_w, r1_w, slots: int = 128, bits: int = 8):
    """The PEBA1 protocol step of the server (src/main.cpp:534-570) for MANY clients at once:
    client k sent its encrypted probe and two encrypted random bytes r0_k, r1_k; the server holds the client's enrolled
    template.  For every k: b_k = Function_f(probe_k, template_k, bound) (1 iff squared distance > bound, quirk Q3), then
    y_k = Function_g(b_k, r0_k, r1_k) = r1_k if b_k else (-r0_k mod 256) (quirks Q1/Q4).  All K instances are recorded
    into one batch, so each circuit level of f (377) and g (116) is ONE launch pair whatever K is.
    probe_vecs / template_vecs: K vectors as made by _host_vector / encrypt_templates_on_device (freed here);
    r0_w, r1_w: [K, bits, words] ciphertext words.  Returns y as [K, bits, words]."""
    L = H.load_library()
    K = len(probe_vecs)
    assert len(template_vecs) == K and r0_w.shape[0] == K and r1_w.shape[0] == K
    eb = ks.new_array(24)
    ks.set_words(eb, bound_w)
    rb = [ks.new_array(24) for _ in range(K)]
    ys = [ks.new_array(bits + 1) for _ in range(K)]
    rs = []
    for k in range(K):
        e0, e1 = ks.new_array(bits), ks.new_array(bits)
        ks.set_words(e0, r0_w[k])
        ks.set_words(e1, r1_w[k])
        rs.append((e0, e1))
    L.peba1_batch_begin()
    for k in range(K):
        L.peba1_math_function_f(rb[k], probe_vecs[k][1], template_vecs[k][1], slots, eb, bits, ks.cloud)
    for k in range(K):
        L.peba1_math_function_g(ys[k], rb[k], rs[k][0], rs[k][1], bits, ks.cloud)
    L.peba1_batch_end()
    out = np.stack([ks.get_words(ys[k], bits) for k in range(K)]) if K else np.zeros((0, bits, ks.words), np.int32)
    for arrs, _ in list(probe_vecs) + list(template_vecs):
        for a in arrs:
            ks.free_array(bits, a)
    for e0, e1 in rs:
        ks.free_array(bits, e0)
        ks.free_array(bits, e1)
    for a in rb + [eb]:
        ks.free_array(24, a)
    for a in ys:
        ks.free_array(bits + 1, a)
    return out


def gather_match_bits(dist, world: int, device, torch, local_out: np.ndarray) -> np.ndarray:
    """All-gather of the encrypted match bits; returns [world*T, words] (rank-major) on every rank."""
    t = torch.from_numpy(np.ascontiguousarray(local_out)).to(device)
    if dist is None or world == 1:
        return local_out
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return torch.cat(parts).cpu().numpy()
