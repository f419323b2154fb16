T="timeout 300 python profiles/microbench/time_vars.py"
{ $T --n 256 --frames 4194304 --scaling 2 --limit 3 50 12000 prev inregall main prev inregall main
$T --n 512 --frames 2097152 --scaling 2 --limit 2 0 8000 prev inregall main
$T --n 512 --frames 2097152 --scaling 0 prev inregall main
$T --n 256 --frames 4194304 --scaling 4 --window 0 prev main; } > gpurun_out/h6.log 2>&1
