#!/bin/bash
# Time the headline workload with each experiment build.  usage (GPU box): run_variants.sh STATS v1 v2 ...
STATS=$1; shift
for v in "$@"; do
  [ -f $PWD/profiles/microbench/variants/libsa_exp$v.so ] || { echo "variant $v not built"; continue; }
  printf "SA_EXP=%s stats=%s: " $v $STATS
  SA_LIB_PATH=$PWD/profiles/microbench/variants/libsa_exp$v.so timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --stats $STATS \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f M spectra/s' % (d['value']/1e6))"
done
