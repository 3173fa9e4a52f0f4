#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/bench_default.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
grep -o '"value": [0-9.]*' gpurun_out/bench_default.log | head -2; grep -o '"stage_ms": {[^}]*}' gpurun_out/bench_default.log; grep -o '"cpu_baseline": {[^}]*}' gpurun_out/bench_default.log
