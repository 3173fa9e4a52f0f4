#!/bin/bash
mkdir -p gpurun_out
./tools/ubench_gather > gpurun_out/ubench_gather.log 2>&1
ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:gather -c 12 --csv --log-file gpurun_out/ubench_gather_ncu.csv ./tools/ubench_gather > /dev/null 2>&1
( time timeout 900 python -m pytest tests/test_gpu_gotoh.py tests/test_gpu_pipeline.py tests/test_cpp_mirror.py -m gpu -x -q ) > gpurun_out/pytest_gpu_b.log 2>&1
( time timeout 900 python tools/compare_ref_cuda.py ) > gpurun_out/compare_ref_cuda.log 2>&1
( time timeout 900 python bench.py --steps 10 --warmup 3 --ktab-k 13 ) > gpurun_out/bench_k13.log 2>&1
( time timeout 900 python bench.py --steps 10 --warmup 3 --ktab-k 14 --no-cpu-baseline ) > gpurun_out/bench_k14.log 2>&1
( time timeout 900 python bench.py --steps 10 --warmup 3 --ktab-k 13 --no-dedup --no-cpu-baseline ) > gpurun_out/bench_k13_nodedup.log 2>&1
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ktab-k 13"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:pipe_seed_match -s 3 -c 1 -f -o gpurun_out/prof_seed_match_r01b $B > gpurun_out/ncu_seed.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:gotoh_pair -s 3 -c 1 -f -o gpurun_out/prof_gotoh_pair_r01b $B --no-dedup > gpurun_out/ncu_gotoh.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'nvb::pipe|nvb::gotoh|cub::DeviceScan' -s 40 -c 60 --csv --log-file gpurun_out/launches_r01b.csv $B > gpurun_out/ncu_launches.log 2>&1
cat gpurun_out/ubench_gather.log; tail -3 gpurun_out/pytest_gpu_b.log; grep -v "^$" gpurun_out/compare_ref_cuda.log | tail -4
for f in bench_k13 bench_k14 bench_k13_nodedup; do echo "== $f"; grep -o '"value": [0-9.]*' gpurun_out/$f.log | head -2; grep -o '"stage_ms": {[^}]*}' gpurun_out/$f.log; grep -o '"GCUPS": [0-9.]*' gpurun_out/$f.log; grep -o '"clocks": {[^}]*}' gpurun_out/$f.log; done
