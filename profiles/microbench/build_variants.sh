#!/bin/bash
# Build timing-experiment variants of the library: one .so per SA_EXP value (see sa_spectrum_v2.cuh).
# usage: profiles/microbench/build_variants.sh 0 1 2 4 ...   (run from the repo root; builds in parallel)
set -e
OUT=profiles/microbench/variants
mkdir -p $OUT
for v in "$@"; do
  env -u CC -u CXX nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off \
    -diag-suppress 177 -shared -DSA_EXP=$v -o $OUT/libsa_exp$v.so spectrum_analyzer_b200/csrc/sa_lib.cu &
done
wait
ls -la $OUT
