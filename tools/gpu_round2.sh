#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_3gbp_r2.log 2>&1
( time timeout 900 python bench.py --steps 5 --warmup 3 --no-dedup --no-cpu-baseline ) > gpurun_out/bench_3gbp_nodedup.log 2>&1
( time timeout 900 python bench.py --steps 5 --warmup 3 --ktab-k 0 --sa-interval 16 --no-dedup --no-cpu-baseline ) > gpurun_out/bench_3gbp_plain.log 2>&1
( time timeout 900 python bench.py --steps 5 --warmup 3 --ktab-k 13 --no-cpu-baseline ) > gpurun_out/bench_3gbp_k13.log 2>&1
( time timeout 900 python bench.py --steps 5 --warmup 3 --ktab-k 11 --no-cpu-baseline ) > gpurun_out/bench_3gbp_k11.log 2>&1
( time timeout 900 python tools/compare_ref_cuda.py ) > gpurun_out/compare_ref_cuda.log 2>&1
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_reference.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
for f in bench_3gbp_r2 bench_3gbp_nodedup bench_3gbp_plain bench_3gbp_k13 bench_3gbp_k11; do echo "== $f"; grep -o '"value": [0-9.]*' gpurun_out/$f.log | head -1; grep -o '"stage_ms": {[^}]*}' gpurun_out/$f.log; grep -o '"GCUPS": [0-9.]*' gpurun_out/$f.log; done
cat gpurun_out/compare_ref_cuda.log | grep -v "^$" | tail -6
tail -3 gpurun_out/bench_reference.log
