T="timeout 300 python profiles/microbench/time_vars.py"
{ $T --n 16384 --frames 65536 --scaling 3 --window 3 --stats 3 --limit 3 50 15000 oldfull main full_p4 oldfull full_p4
$T --n 8192 --frames 196608 --scaling 1 oldfull main full_p4
$T --n 16384 --frames 65536 --scaling 1 oldfull main full_p4; } > gpurun_out/f4.log 2>&1
