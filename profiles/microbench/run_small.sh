for v in main small3 small4; do
  if [ "$v" != "main" ]; then export SA_LIB_PATH=$PWD/profiles/microbench/variants/libsa_$v.so; fi
  for c in "cfg5 256" "64 hann" "512 hann"; do python profiles/bench_configs.py --only "$c" 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$v', d['config'][:40], '%.3e'%d['frames_per_s'], '%.3f'%d['hbm_frac_of_measured'])"; done
done > gpurun_out/small_vars.log 2>&1
