T="timeout 300 python profiles/microbench/time_vars.py"
{ $T --n 16384 --frames 65536 --scaling 3 --window 3 --stats 3 --limit 3 50 15000 prev main prev main
$T --n 16384 --frames 65536 --scaling 3 --window 3 --stats 3 prev main
$T --n 4096 --scaling 3 --stats 3 prev main
$T --n 4096 --scaling 3 --stats 7 prev main
$T --n 2048 --frames 786432 --scaling 3 --stats 1 prev main
$T --n 8192 --frames 196608 --scaling 3 --stats 3 prev main; } > gpurun_out/n1.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/n1_tests.log 2>&1
