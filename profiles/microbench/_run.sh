T="timeout 300 python profiles/microbench/time_vars.py"
{ $T --n 16384 --frames 65536 --scaling 3 --window 3 --stats 3 --limit 3 50 15000 prev pad main prev pad main
$T --n 16384 --frames 65536 --scaling 1 prev pad main
$T --n 8192 --frames 196608 --scaling 1 prev pad main
$T --n 8192 --frames 196608 --scaling 3 --stats 3 prev pad main; } > gpurun_out/p3.log 2>&1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/p3_tests.log 2>&1
