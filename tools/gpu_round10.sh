#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( timeout 900 python tools/compare_ref_cuda.py ) > gpurun_out/compare_ref_cuda.log 2>&1
( python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_default.log 2>&1
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'nvb::|cub::' -c 400 --csv --log-file gpurun_out/launches_r01e.csv $B > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:pipe_seed_match -s 3 -c 1 -f -o gpurun_out/prof_seed_match_r01e $B > gpurun_out/ncu_seed.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:gotoh_pair -s 3 -c 1 -f -o gpurun_out/prof_gotoh_pair_r01e $B > gpurun_out/ncu_gotoh.log 2>&1
tail -4 gpurun_out/pytest_gpu.log | head -2
grep -v "^$" gpurun_out/compare_ref_cuda.log | cut -c1-500
grep -o '"value": [0-9.]*' gpurun_out/bench_default.log | head -2; grep -o '"stage_ms": {[^}]*}' gpurun_out/bench_default.log
tail -2 gpurun_out/ncu_seed.log gpurun_out/ncu_gotoh.log
