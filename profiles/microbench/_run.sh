timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/h4_tests.log 2>&1
timeout 900 python profiles/bench_configs.py > gpurun_out/h4_configs.jsonl 2> gpurun_out/h4_configs.err
exp() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null; ncu -i gpurun_out/$1.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null | gzip > gpurun_out/$1_src.csv.gz; rm -f gpurun_out/$1.ncu-rep; }
NCU="ncu --set full --clock-control none --import-source on -s 3 -c 1"
python profiles/run_one_config.py 256 256 2097152 4 0 7 > gpurun_out/h4_plain_cfg5.log 2>&1 && $NCU -k regex:spectrum_kernel_small -o gpurun_out/h4_cfg5 python profiles/run_one_config.py 256 256 2097152 4 0 7 > gpurun_out/h4_ncu_cfg5.log 2>&1
exp h4_cfg5
