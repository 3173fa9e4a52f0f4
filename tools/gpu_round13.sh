#!/bin/bash
mkdir -p gpurun_out
( python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs --ktab-k 16 ) > gpurun_out/bench_k16.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:gotoh_full_pair -s 1 -c 1 -f -o gpurun_out/prof_full_pair_local python tools/bench_full.py --one > gpurun_out/ncu_full.log 2>&1
grep -o '"value": [0-9.]*' gpurun_out/bench_k16.log | head -2; grep -o '"stage_ms": {[^}]*}' gpurun_out/bench_k16.log; grep -o '"index": {[^}]*}' gpurun_out/bench_k16.log
tail -3 gpurun_out/bench_k16.log | cut -c1-300
tail -2 gpurun_out/ncu_full.log
