#!/bin/bash
# One GPU session: the -m gpu suite, the default bench line, the reference-CUDA comparison and the ncu captures that
# profiles/ is built from.  Run as: gpurun --timeout 1800 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-rXX}
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_default.log 2>&1
( timeout 900 python tools/compare_ref_cuda.py ) > gpurun_out/compare_ref_cuda.log 2>&1
( timeout 900 python tools/bench_full.py ) > gpurun_out/bench_full.log 2>&1
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs --pairs 0"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'nvb::|cub::' -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:pipe_seed_match -s 2 -c 2 -f -o gpurun_out/prof_seed_match_$TAG $B > gpurun_out/ncu_seed.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:gotoh_pair_kernel -s 3 -c 1 -f -o gpurun_out/prof_gotoh_pair_$TAG $B > gpurun_out/ncu_gotoh.log 2>&1
tail -4 gpurun_out/pytest_gpu.log | head -2
grep -o '"value": [0-9.]*' gpurun_out/bench_default.log | head -2; grep -o '"stage_ms": {[^}]*}' gpurun_out/bench_default.log
grep -v "^$" gpurun_out/compare_ref_cuda.log | cut -c1-400
