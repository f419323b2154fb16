# round-2 secondary measurements on one B200: cfg3 as specified at N=1, every config with clocks, ncu captures of the
# kernels behind cfg2 / cfg3 / cfg4 / cfg5 (each run first without ncu)
python profiles/run_cfg3_sharded.py --hours 10 > gpurun_out/multi_cfg3_n1.log 2>&1
python profiles/bench_configs.py > gpurun_out/r02_configs.jsonl 2> gpurun_out/r02_configs.err
# the .ncu-rep files are too big to travel back together: export the two CSV pages on the box, keep only those
exp() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null; ncu -i gpurun_out/$1.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null | gzip > gpurun_out/$1_src.csv.gz; rm -f gpurun_out/$1.ncu-rep; }
NCU="ncu --set full --clock-control none --import-source on -s 3 -c 1"
python bench.py --steps 2 --warmup 3 --frames 262144 --no-cpu --no-e2e > gpurun_out/r02_plain_cfg2.log 2>&1 && $NCU -k regex:spectrum_kernel_v2 -o gpurun_out/r02_cfg2 python bench.py --steps 2 --warmup 3 --frames 262144 --no-cpu --no-e2e > gpurun_out/r02_ncu_cfg2.log 2>&1
exp r02_cfg2
python profiles/run_one_config.py 2048 512 400000 0 2 7 > gpurun_out/r02_plain_cfg3.log 2>&1 && $NCU -k regex:spectrum_kernel_v2 -o gpurun_out/r02_cfg3 python profiles/run_one_config.py 2048 512 400000 0 2 7 > gpurun_out/r02_ncu_cfg3.log 2>&1
exp r02_cfg3
python profiles/run_one_config.py 16384 16384 32768 3 3 3 3 50 15000 > gpurun_out/r02_plain_cfg4.log 2>&1 && $NCU -k regex:spectrum_kernel_v2 -o gpurun_out/r02_cfg4 python profiles/run_one_config.py 16384 16384 32768 3 3 3 3 50 15000 > gpurun_out/r02_ncu_cfg4.log 2>&1
exp r02_cfg4
python profiles/run_one_config.py 256 256 2097152 4 0 7 > gpurun_out/r02_plain_cfg5.log 2>&1 && $NCU -k regex:spectrum_kernel_small -o gpurun_out/r02_cfg5 python profiles/run_one_config.py 256 256 2097152 4 0 7 > gpurun_out/r02_ncu_cfg5.log 2>&1
exp r02_cfg5
python profiles/run_one_config.py 32768 32768 16384 0 1 7 > gpurun_out/r02_plain_32k.log 2>&1 && $NCU -k regex:spectrum_kernel -o gpurun_out/r02_32k python profiles/run_one_config.py 32768 32768 16384 0 1 7 > gpurun_out/r02_ncu_32k.log 2>&1
exp r02_32k
