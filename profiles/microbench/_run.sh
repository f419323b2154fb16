timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/i1_smoke.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/i1_tests.log 2>&1
timeout 600 python bench.py > gpurun_out/i1_bench.json 2> gpurun_out/i1_bench.err
sleep 15
timeout 900 python profiles/bench_configs.py > gpurun_out/i1_configs.jsonl 2> gpurun_out/i1_configs.err
