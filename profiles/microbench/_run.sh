timeout 300 python profiles/microbench/time_vars.py base2 reord base2 reord > gpurun_out/j1.log 2>&1
SA_LIB_PATH=$PWD/profiles/microbench/variants/libsa_reord.so timeout 300 python -m pytest tests/test_median_stress_gpu.py -x -q -m gpu -k "4096" > gpurun_out/j1_tests.log 2>&1
