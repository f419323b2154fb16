exp() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null; ncu -i gpurun_out/$1.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null | gzip > gpurun_out/$1_src.csv.gz; rm -f gpurun_out/$1.ncu-rep; }
NCU="ncu --set full --clock-control none --import-source on -s 3 -c 1"
python profiles/run_one_config.py 2048 512 400000 0 2 7 > gpurun_out/o1_plain_cfg3.log 2>&1 && $NCU -k regex:spectrum_kernel_v2 -o gpurun_out/o1_cfg3 python profiles/run_one_config.py 2048 512 400000 0 2 7 > gpurun_out/o1_ncu_cfg3.log 2>&1
exp o1_cfg3
python profiles/run_one_config.py 16384 16384 32768 3 3 3 3 50 15000 > gpurun_out/o1_plain_cfg4.log 2>&1 && $NCU -k regex:spectrum_kernel_v2 -o gpurun_out/o1_cfg4 python profiles/run_one_config.py 16384 16384 32768 3 3 3 3 50 15000 > gpurun_out/o1_ncu_cfg4.log 2>&1
exp o1_cfg4
