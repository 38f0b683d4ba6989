"""demo/gallery_nccl.cpp — the C++ multi-GPU host path (one process per GPU, NCCL broadcast of the device-format keys,
NCCL scatter of the gallery, Function_f with the reference's own signature, NCCL all-gather of the encrypted match bits).
On a one-GPU box the communicator has one rank; with two or more GPUs the same program shards the gallery."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "demo", "peba1_gallery_nccl")


def run(*args, timeout=600):
    out = subprocess.run([BIN, *map(str, args)], capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_driver_is_built_and_links_nccl():
    if not os.path.exists(BIN):
        pytest.skip("demo/peba1_gallery_nccl not built (make -C demo nccl)")
    libs = subprocess.run(["ldd", BIN], capture_output=True, text=True).stdout
    assert "libnccl" in libs and "libpeba1_tfhe" in libs and "libpeba1_cuda" in libs


@pytest.mark.gpu
def test_one_rank_faithful_and_folded():
    r = run("--gpus", 1, "--templates", 3, "--slots", 4, "--mode", 0)       # 4-slot vectors: 3 x 6,852 bootstraps (minimum: 24 XNOR + 48 MUX of 2)
    assert r["verified"] and r["templates"] == 3 and r["pbs_per_gpu"] == 3 * (4 * 1683 + 120)
    r = run("--gpus", 1, "--templates", 4, "--mode", 3)                     # full 128-slot vectors, folded circuit
    assert r["verified"] and r["matches_found"] == 1


@pytest.mark.gpu
def test_two_ranks_shard_the_gallery():
    import peba1_b200 as P
    if P.load_library().peba1_cuda_device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    r = run("--gpus", 2, "--templates", 8, "--mode", 3)
    assert r["verified"] and r["n_gpus"] == 2 and r["templates_per_gpu"] == 4 and r["matches_found"] == 2
    assert r["key_broadcast_bytes"] > 120e6
