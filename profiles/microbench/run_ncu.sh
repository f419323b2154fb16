# usage: run_ncu.sh OUT variant kernel-regex   (variant = lib name in profiles/microbench/variants or "main")
OUT=$1; V=$2; K=$3
if [ "$V" != "main" ]; then export SA_LIB_PATH=$PWD/profiles/microbench/variants/libsa_$V.so; fi
python bench.py --steps 2 --warmup 3 --frames 262144 --no-cpu --no-e2e > gpurun_out/${OUT}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/${OUT}_prof python bench.py --steps 2 --warmup 3 --frames 262144 --no-cpu --no-e2e > gpurun_out/${OUT}_ncu.log 2>&1
