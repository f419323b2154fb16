timeout 900 python -m pytest tests -q -m gpu > gpurun_out/q3_tests.log 2>&1
timeout 300 python profiles/microbench/time_vars.py --limit 3 50 12000 prev main > gpurun_out/q3.log 2>&1
