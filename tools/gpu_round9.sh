#!/bin/bash
mkdir -p gpurun_out
./tools/ubench_gather > gpurun_out/ubench_gather3.log 2>&1
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( python bench.py --steps 10 --warmup 3 --no-other-configs ) > gpurun_out/bench_k14.log 2>&1
( python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs --ktab-k 15 ) > gpurun_out/bench_k15.log 2>&1
( timeout 600 python tools/compare_ref_cuda.py ) > gpurun_out/compare_ref_cuda.log 2>&1
grep sweep gpurun_out/ubench_gather3.log; tail -4 gpurun_out/pytest_gpu.log | head -2
for f in bench_k14 bench_k15; do echo "== $f"; grep -o '"value": [0-9.]*' gpurun_out/$f.log | head -2; grep -o '"stage_ms": {[^}]*}' gpurun_out/$f.log; grep -o '"index": {[^}]*}' gpurun_out/$f.log; done
grep -v "^$" gpurun_out/compare_ref_cuda.log | head -3 | cut -c1-400
