# multi-GPU measurements of round 2 (run on an N-GPU box): usage run_multi.sh NGPUS
G=$1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4 8; do
  [ $n -le $G ] || continue
  timeout 600 $TR --nproc-per-node $n --master-port $((29500+n)) profiles/run_cfg3_sharded.py --hours 10 > gpurun_out/multi_cfg3_n$n.log 2>&1
done
timeout 600 python profiles/threads_one_process.py --gpus $G > gpurun_out/multi_threads_n$G.log 2>&1
timeout 900 $TR --nproc-per-node $G --master-port 29599 bench.py --gpus $G --steps 10 --warmup 3 > gpurun_out/multi_bench_n$G.json 2> gpurun_out/multi_bench_n$G.err
nvidia-smi topo -m > gpurun_out/multi_topo.txt 2>&1
