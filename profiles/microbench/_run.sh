timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/g1_tests.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/g1_bench.json 2> gpurun_out/g1_bench.err
sleep 20
timeout 900 python profiles/bench_configs.py > gpurun_out/g1_configs.jsonl 2> gpurun_out/g1_configs.err
exp() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null; ncu -i gpurun_out/$1.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null | gzip > gpurun_out/$1_src.csv.gz; rm -f gpurun_out/$1.ncu-rep; }
NCU="ncu --set full --clock-control none --import-source on -s 3 -c 1"
python bench.py --steps 2 --warmup 3 --frames 262144 --no-cpu --no-e2e > gpurun_out/g1_plain_cfg2.log 2>&1 && $NCU -k regex:spectrum_kernel_v2 -o gpurun_out/g1_cfg2 python bench.py --steps 2 --warmup 3 --frames 262144 --no-cpu --no-e2e > gpurun_out/g1_ncu_cfg2.log 2>&1
exp g1_cfg2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/g1_launches.csv python bench.py --steps 2 --warmup 3 --frames 262144 --no-cpu --no-e2e > gpurun_out/g1_launch_run.log 2>&1
python profiles/run_one_config.py 1024 1024 1000000 1 1 7 > gpurun_out/g1_plain_1024.log 2>&1 && $NCU -k regex:spectrum_kernel_v2 -o gpurun_out/g1_1024 python profiles/run_one_config.py 1024 1024 1000000 1 1 7 > gpurun_out/g1_ncu_1024.log 2>&1
exp g1_1024
