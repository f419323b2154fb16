timeout 300 python profiles/microbench/time_vars.py main s5 s7 s8 g1 s7g1 m16 m18 h8 main > gpurun_out/l1.log 2>&1
