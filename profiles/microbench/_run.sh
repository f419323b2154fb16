timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/e1_tests.log 2>&1
python profiles/microbench/time_vars.py main a4 > gpurun_out/e1.log 2>&1
exp() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null; ncu -i gpurun_out/$1.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null | gzip > gpurun_out/$1_src.csv.gz; rm -f gpurun_out/$1.ncu-rep; }
NCU="ncu --set full --clock-control none --import-source on -s 3 -c 1"
python bench.py --steps 2 --warmup 3 --frames 262144 --no-cpu --no-e2e > gpurun_out/e1_plain_cfg2.log 2>&1 && $NCU -k regex:spectrum_kernel_v2 -o gpurun_out/e1_cfg2 python bench.py --steps 2 --warmup 3 --frames 262144 --no-cpu --no-e2e > gpurun_out/e1_ncu_cfg2.log 2>&1
exp e1_cfg2
