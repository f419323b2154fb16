#!/bin/bash
# Build one named experiment variant of the library (headline size only: -DSA_ONLY_4096, ~20 s).
# usage: profiles/microbench/build_var.sh NAME [-Dflags ...]   -> profiles/microbench/variants/libsa_NAME.so
set -e
NAME=$1; shift
OUT=profiles/microbench/variants
mkdir -p $OUT
env -u CC -u CXX nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off \
  -diag-suppress 177 -shared -DSA_ONLY_4096 "$@" -Xptxas -v -o $OUT/libsa_$NAME.so spectrum_analyzer_b200/csrc/sa_lib.cu 2>&1 \
  | grep -A2 "spectrum_kernel_v2.*Li0ELb1EEE" | grep -E "spill|Used" | sed "s/^/$NAME: /"
