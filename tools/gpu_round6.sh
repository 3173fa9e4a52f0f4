#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/bench_default.log 2>&1
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:pipe_seed_match -s 3 -c 1 -f -o gpurun_out/prof_seed_match_r01c $B > gpurun_out/ncu_seed.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:gotoh_pair -s 3 -c 1 -f -o gpurun_out/prof_gotoh_pair_r01c $B > gpurun_out/ncu_gotoh.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'nvb::pipe|nvb::gotoh|cub::DeviceScan' -s 60 -c 48 --csv --log-file gpurun_out/launches_r01c.csv $B > gpurun_out/ncu_launches.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
grep -o '"value": [0-9.]*' gpurun_out/bench_default.log | head -2; grep -o '"stage_ms": {[^}]*}' gpurun_out/bench_default.log; grep -o '"cpu_baseline": {[^}]*}' gpurun_out/bench_default.log
