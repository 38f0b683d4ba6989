"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/peba1_cuda.h declares,
and fails loudly (no fallback) without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import peba1_b200 as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "peba1_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(peba1_cuda_[a-z0-9_]+)\s*\(", src)))


def test_header_symbol_list_matches_binding():
    assert _header_symbols() == sorted(P.DECLARED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = P.load_library()
    for s in _header_symbols():
        assert hasattr(L, s), f"libpeba1_cuda.so does not export {s}"


def test_struct_layouts():
    assert C.sizeof(P.Gate) == 40
    assert C.sizeof(P.Params) == 28
    from peba1_b200.cuda import GATE_DTYPE
    assert GATE_DTYPE.itemsize == 40 and GATE_DTYPE.fields["out"][1] == 32


def test_no_device_fails_loudly():
    L = P.load_library()
    if L.peba1_cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(P.PEBA1Error) as ei:
        P.Engine()
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_bad_params_rejected():
    L = P.load_library()
    h = C.c_void_p()
    bad = P.Params(630, 2048, 1, 3, 7, 8, 2)
    assert L.peba1_cuda_create(0, C.byref(bad), C.byref(h)) == -1
    assert b"unsupported parameter set" in L.peba1_cuda_last_error()
    assert L.peba1_cuda_create(0, None, C.byref(h)) == -1


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (test infrastructure only)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "peba1_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "tfhe_oracle" not in txt and "pyoracle" not in txt, os.path.join(dirpath, f)


def test_host_library_exports_every_declared_function():
    """libpeba1_tfhe.so defines every function the TFHE-compatible headers, Math.h mirror and host extensions declare
    (C symbols by name; the C++ entry points of peba1_math.h — std::vector arguments like the reference's — mangled)."""
    import subprocess
    from peba1_b200 import tfhe_host as H
    H.load_library()
    nm = subprocess.run(["nm", "-D", "--defined-only", H.lib_path()], capture_output=True, text=True, check=True).stdout
    exported = [ln.split()[-1] for ln in nm.splitlines() if ln.strip()]
    inc = os.path.join(ROOT, "include")
    headers = [os.path.join(inc, "peba1_host.h"), os.path.join(inc, "peba1_math.h")] + \
              [os.path.join(inc, "tfhe", f) for f in sorted(os.listdir(os.path.join(inc, "tfhe")))]
    pat = re.compile(r"\b((?:boots|peba1_|new_|delete_|export_|import_|tfhe_random_|HE_|Function_|compare_bit|minimum)"
                     r"[A-Za-z0-9_]*)\s*\(")
    declared = set()
    for h in headers:
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        declared |= set(pat.findall(src))
    declared = {d for d in declared if not d.startswith("peba1_cuda_")}   # libpeba1_cuda's, checked above
    assert len(declared) > 60
    missing = [d for d in sorted(declared) if d not in exported and not any(d in e for e in exported if e.startswith("_Z"))]
    assert not missing, missing


def test_plain_c_caller_fails_loudly_without_a_device():
    """A C program that binds nothing but include/peba1_cuda.h (scripts/ubench/kbench.c, the Python-free bench harness)
    links against the library and, without a device, stops at peba1_cuda_create with the library's own message."""
    import subprocess
    L = P.load_library()
    if L.peba1_cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "peba1_b200", "csrc"),
                           "../../scripts/ubench/kbench_bin"])
    r = subprocess.run([os.path.join(ROOT, "scripts", "ubench", "kbench_bin"), "batch=8"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr and r.stdout == ""


def test_every_engine_option_is_documented_in_the_header():
    """peba1_cuda_set_option's names (csrc/peba1_cuda.cu) all appear in include/peba1_cuda.h"""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "peba1_b200", "csrc", "peba1_cuda.cu")).read()
    hdr = open(os.path.join(root, "include", "peba1_cuda.h")).read()
    names = re.findall(r'strcmp\(name, "([a-z_0-9]+)"\)', src[src.index("int peba1_cuda_set_option("):])
    assert len(names) >= 10
    assert [n for n in names if n not in hdr] == []
