#!/bin/bash
# Builds what-if variants of libpeba1_cuda (WRONG results by construction) to see which resource bounds the blind
# rotation: scripts/ablate.sh <n>  ->  scripts/ubench/libablate_<n>.so ; run with PEBA1_CUDA_LIB=... scripts/prof_pbs.py
set -e
cd "$(dirname "$0")/../peba1_b200/csrc"
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -diag-suppress 20011,20014 \
     -DPEBA1_ABLATE=$1 -shared -o ../../scripts/ubench/libablate_$1.so peba1_cuda.cu -lcudart
