#!/bin/bash
# ASan + UBSan run of the host side (recorder, level scheduler, circuit builder, key generation, tfhe_io) — SURVEY.md §5.
# Builds libpeba1_tfhe from the same sources with -fsanitize=address,undefined into a scratch directory and runs the CPU
# test files that drive it (simulator / oracle backends; no GPU needed).  Output: profiles/r2_host_asan_ubsan.log
set -e
cd "$(dirname "$0")/.."
OUT=${TMPDIR:-/tmp}/peba1_asan
mkdir -p "$OUT"
CXX=/usr/bin/g++   # the distribution compiler ships the sanitizer runtimes (/opt/gcc does not)
$CXX -O1 -g -std=c++17 -fPIC -fno-omit-frame-pointer -fsanitize=address,undefined -fno-sanitize-recover=undefined \
    -Iinclude -pthread -shared -o "$OUT/libpeba1_tfhe.so" \
    peba1_b200/host/tfhe_front.cpp peba1_b200/host/math_circuits.cpp peba1_b200/host/tfhe_io.cpp \
    -Lpeba1_b200/csrc -lpeba1_cuda -Wl,-rpath,"$PWD/peba1_b200/csrc" -pthread
ASAN=$($CXX -print-file-name=libasan.so)
UBSAN=$($CXX -print-file-name=libubsan.so)
# python itself is not instrumented: leak checking would report the interpreter's own allocations
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=1:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
LD_PRELOAD="$ASAN:$UBSAN" PEBA1_TFHE_LIB="$OUT/libpeba1_tfhe.so" python -m pytest -x -q \
    tests/test_circuits.py tests/test_tfhe_io.py tests/test_host_rng.py tests/test_oracle_backend.py \
    "tests/test_refmath_parity.py::test_product_circuits_on_the_simulator_match_the_reference_counts_and_bits" 2>&1
