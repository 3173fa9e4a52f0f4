#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python tools/compare_ref_cuda.py ) > gpurun_out/compare_ref_cuda.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k c1 ) > gpurun_out/pytest_c1.log 2>&1
./tools/ubench_gather > gpurun_out/ubench_gather4.log 2>&1
grep -v "^$" gpurun_out/compare_ref_cuda.log | cut -c1-500
tail -3 gpurun_out/pytest_c1.log
grep ordered gpurun_out/ubench_gather4.log
