"""Full-size parity through a checksum (BASELINE configs[1]: 16,384 gates in one batch).

scripts/ubench/kbench.c drives the CUDA path through the plain C ABI on pseudo-random keys and inputs (xorshift, fixed
seed) and prints an FNV-1a checksum of all batch x (n+1) output words; scripts/ubench/kbench_oracle.c runs the CPU oracle
on exactly the same stream and prints the same checksum.  tests/golden/kbench_checksums.json holds what each side printed
(GPU: on B200, profiles/r1zzz_kbench_two_buffers.log; oracle: in the build container).  Equal checksums = every output
word identical, at a size where a word-by-word comparison in the CPU suite would take minutes of oracle time."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "kbench_checksums.json")))
UB = os.path.join(ROOT, "scripts", "ubench")


def _checksum(cmd):
    out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    return json.loads(out.strip().splitlines()[-1])["checksum"]


def test_recorded_gpu_and_oracle_checksums_agree():
    for case, v in GOLD.items():
        if case.startswith("_") or v["oracle"] is None:
            continue
        assert v["gpu"] == v["oracle"], case
        if v.get("oracle_exact") is not None:   # the exact-integer form of the oracle on the same workload
            assert v["gpu"] == v["oracle_exact"], case


def test_oracle_reproduces_the_gpu_checksum_of_a_narrow_level():
    """81 gates (the median level width of a face match): about a second of oracle time."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "kbench"])
    got = _checksum([os.path.join(UB, "kbench_oracle_bin"), "batch=81", f"threads={os.cpu_count() or 1}"])
    assert got == GOLD["batch=81 op=0"]["gpu"]


def test_exact_integer_oracle_rederives_a_recorded_checksum():
    """The `oracle_exact` entries are not just recorded: here the oracle's EXACT-INTEGER form (O(N^2) negacyclic products, no
    floating point) recomputes the 200-gate batch — two waves of the low-latency kernel on the GPU — and must print the
    checksum the GPU printed (about 30 s on 8 cores); the FP64-FFT form does the mid-size shapes in seconds."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "kbench"])
    o = os.path.join(UB, "kbench_oracle_bin")
    threads = f"threads={os.cpu_count() or 1}"
    assert _checksum([o, "batch=200", threads, "mode=0"]) == GOLD["batch=200 op=0"]["gpu"]
    assert _checksum([o, "batch=648", threads]) == GOLD["batch=648 op=0"]["gpu"]
    assert _checksum([o, "batch=1036", threads]) == GOLD["batch=1036 op=0"]["gpu"]


@pytest.mark.skipif(os.environ.get("PEBA1_SLOW") != "1", reason="two minutes on 8 cores; PEBA1_SLOW=1 runs it")
def test_oracle_reproduces_the_gpu_checksum_at_full_size():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "kbench"])
    got = _checksum([os.path.join(UB, "kbench_oracle_bin"), "batch=16384", f"threads={os.cpu_count() or 1}"])
    assert got == GOLD["batch=16384 op=0"]["gpu"]


def _kbench_bin():
    """the prebuilt harness travels with the snapshot; compile it against the prebuilt library if it is missing"""
    path = os.path.join(UB, "kbench_bin")
    if not os.path.exists(path):
        csrc = os.path.join(ROOT, "peba1_b200", "csrc")
        subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), "-o", path,
                               os.path.join(UB, "kbench.c"), "-L" + csrc, "-lpeba1_cuda",
                               "-Wl,-rpath,$ORIGIN/../../peba1_b200/csrc"])
    return path


@pytest.mark.gpu
def test_gpu_checksums_equal_the_oracle_at_full_size():
    """16,384 bootsNAND through the default path (persistent wide-batch kernel + keyswitch) and through the one-buffer
    kernel it replaced, 81 through the low-latency kernel and the split keyswitch, 16,384 bootsMUX (two bootstraps and one
    keyswitch each): the oracle's checksums."""
    k = _kbench_bin()
    want = GOLD["batch=16384 op=0"]["oracle"]
    assert _checksum([k, "batch=16384", "steps=1", "warmup=0"]) == want
    assert _checksum([k, "batch=16384", "steps=1", "warmup=0", "kernel_variant=8"]) == want
    assert _checksum([k, "batch=81", "steps=1", "warmup=0"]) == GOLD["batch=81 op=0"]["oracle"]
    for n in (200, 648, 1036):   # two waves of the low-latency kernel; narrow shapes of 5 and 7 teams per CTA
        assert _checksum([k, f"batch={n}", "steps=1", "warmup=0"]) == GOLD[f"batch={n} op=0"]["oracle"], n
    assert _checksum([k, "batch=16384", "op=10", "steps=1", "warmup=0"]) == GOLD["batch=16384 op=10"]["oracle"]  # bootsMUX
