#!/bin/bash
# L2 fetch-granularity variants of the FM block gathers inside the real kernel + final ncu traffic capture
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-other-configs"
( $B --steps 10 --warmup 3 ) > gpurun_out/bench_ld64.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:pipe_seed_match -s 3 -c 1 -f -o gpurun_out/prof_seed_match_r01d $B --steps 2 --warmup 3 > gpurun_out/ncu_seed.log 2>&1
NVB_NVCC_EXTRA='-DNVB_FM_LD_QUAL=""' python -m nvbio_b200.build --force > gpurun_out/build_plain.log 2>&1
( $B --steps 10 --warmup 3 ) > gpurun_out/bench_ldplain.log 2>&1
NVB_NVCC_EXTRA='-DNVB_FM_LD_QUAL=".L2::128B"' python -m nvbio_b200.build --force > gpurun_out/build_128.log 2>&1
( $B --steps 10 --warmup 3 ) > gpurun_out/bench_ld128.log 2>&1
for f in bench_ld64 bench_ldplain bench_ld128; do echo "== $f"; grep -o '"value": [0-9.]*' gpurun_out/$f.log | head -1; grep -o '"seed_match": [0-9.]*' gpurun_out/$f.log; grep -o '"locate_windows": [0-9.]*' gpurun_out/$f.log; done
