#!/bin/bash
# first GPU visit: smoke, parity tests, instruction-rate microbenchmark, short bench runs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
( time python __graft_entry__.py --smoke ) > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
./tools/ubench_dpx > gpurun_out/ubench_dpx.log 2>&1
( time timeout 600 python bench.py --genome-mbp 100 --steps 5 --warmup 3 ) > gpurun_out/bench_100mbp.log 2>&1
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_3gbp.log 2>&1
tail -5 gpurun_out/smoke.log; tail -15 gpurun_out/pytest_gpu.log; cat gpurun_out/ubench_dpx.log; tail -3 gpurun_out/bench_100mbp.log; tail -3 gpurun_out/bench_3gbp.log
