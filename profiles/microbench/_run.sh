for v in p18; do SA_LIB_PATH=$PWD/profiles/microbench/variants/libsa_$v.so timeout 300 python -m pytest tests/test_median_stress_gpu.py tests/test_parity_gpu.py -x -q -m gpu -k "4096" > gpurun_out/h1_tests_$v.log 2>&1; done
timeout 300 python profiles/microbench/time_vars.py p17nosoa p18nosoa p18 p17nosoa p18nosoa p18 > gpurun_out/h1.log 2>&1
