#!/bin/bash
# usage: tools/gpu_multi.sh N
N=$1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_$N.txt
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/bench_n$N.log 2>&1
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 1 --impl reference ) > gpurun_out/bench_ref_n$N.log 2>&1
grep -o '"value": [0-9.]*' gpurun_out/bench_n$N.log | head -2; grep -o '"index": {[^}]*}' gpurun_out/bench_n$N.log; grep -o '"n_gpus": [0-9]*' gpurun_out/bench_n$N.log | head -1
tail -4 gpurun_out/bench_n$N.log | cut -c1-400; tail -3 gpurun_out/bench_ref_n$N.log | cut -c1-300
