#!/bin/bash
# ncu evidence for round 1: launch list of one bench run + full captures of the two dominant kernels
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'nvb::|cub::' -c 400 --csv --log-file gpurun_out/launches_r01.csv $B > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:pipe_seed_match -s 3 -c 1 -f -o gpurun_out/prof_seed_match_r01 $B > gpurun_out/ncu_seed.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:gotoh_pair -s 3 -c 1 -f -o gpurun_out/prof_gotoh_pair_r01 $B > gpurun_out/ncu_gotoh.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:pipe_expand_hits -s 3 -c 1 -f -o gpurun_out/prof_locate_r01 $B > gpurun_out/ncu_locate.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q ) > gpurun_out/pytest_pipeline.log 2>&1
tail -3 gpurun_out/ncu_launches.log gpurun_out/ncu_seed.log gpurun_out/ncu_gotoh.log gpurun_out/ncu_locate.log gpurun_out/pytest_pipeline.log; ls -la gpurun_out
