# usage: run_vars.sh OUTNAME v1 v2 ...  -> gpurun_out/OUTNAME.log (timing-experiment builds in profiles/microbench/variants)
# a variant may carry bench arguments after a colon: v3h:--stats=3
OUT=$1; shift
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e"
pr() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%s %.2f M spectra/s frac %.3f' % (sys.argv[1], d['value']/1e6, d['roofline']['frac']))" "$1"; }
( for v in "$@"; do lib=${v%%:*}; extra=""; [ "$lib" != "$v" ] && extra=${v#*:}
  if [ "$lib" = "main" ]; then timeout 300 $B $extra | pr $v; else SA_LIB_PATH=$PWD/profiles/microbench/variants/libsa_$lib.so timeout 300 $B $extra | pr $v; fi; done ) > gpurun_out/$OUT.log 2>&1
