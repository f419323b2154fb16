timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r1_tests.log 2>&1
timeout 600 python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
sleep 15
timeout 900 python profiles/bench_configs.py > gpurun_out/r1_configs.jsonl 2> gpurun_out/r1_configs.err
exp() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null; ncu -i gpurun_out/$1.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null | gzip > gpurun_out/$1_src.csv.gz; rm -f gpurun_out/$1.ncu-rep; }
NCU="ncu --set full --clock-control none --import-source on -s 3 -c 1"
python profiles/run_one_config.py 16384 16384 32768 3 3 3 3 50 15000 > gpurun_out/r1_plain_cfg4.log 2>&1 && $NCU -k regex:spectrum_kernel_v2 -o gpurun_out/r1_cfg4 python profiles/run_one_config.py 16384 16384 32768 3 3 3 3 50 15000 > gpurun_out/r1_ncu_cfg4.log 2>&1
exp r1_cfg4
