#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_gotoh.py -m gpu -x -q -k full ) > gpurun_out/pytest_full.log 2>&1
( timeout 900 python tools/bench_full.py ) > gpurun_out/bench_full.log 2>&1
tail -5 gpurun_out/pytest_full.log
cat gpurun_out/bench_full.log | cut -c1-300
