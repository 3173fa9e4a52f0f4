#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_default.log 2>&1
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs --pairs 0"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:pipe_seed_match -s 3 -c 1 -f -o gpurun_out/prof_seed_match_r01g $B > gpurun_out/ncu_seed.log 2>&1
tail -6 gpurun_out/pytest_gpu.log | head -3
grep -o '"value": [0-9.]*' gpurun_out/bench_default.log | head -2; grep -o '"stage_ms": {[^}]*}' gpurun_out/bench_default.log
grep -o '"paired_end": {.*' gpurun_out/bench_default.log | cut -c1-1500
tail -3 gpurun_out/bench_default.log | cut -c1-300
tail -2 gpurun_out/ncu_seed.log
