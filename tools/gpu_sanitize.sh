#!/bin/bash
# compute-sanitizer memcheck over a small-size slice of the GPU test-suite
mkdir -p gpurun_out
T="tests/test_gpu_fm.py::test_golden_fixtures tests/test_gpu_fm.py::test_device_suffix_sort_repetitive tests/test_gpu_fm.py::test_ktab_gives_identical_ranges tests/test_gpu_fm.py::test_filter_rank_locate tests/test_gpu_fm.py::test_sampled_sa_interval tests/test_gpu_gotoh.py::test_reference_asserted_problems tests/test_gpu_gotoh.py::test_golden_random tests/test_gpu_gotoh.py::test_mixed_batch_fallback_list tests/test_gpu_gotoh.py::test_quality_table_scheme tests/test_gpu_gotoh.py::test_full_matrix_reference_strings tests/test_gpu_gotoh.py::test_full_matrix_packed_path tests/test_gpu_gotoh.py::test_full_matrix_golden_gpu tests/test_gpu_gotoh.py::test_full_matrix_traceback_vs_oracle tests/test_gpu_pipeline.py::test_seed_extend_4bit_reads_with_N tests/test_gpu_pipeline.py::test_unaligned_2bit_reads_and_streaming_api tests/test_gpu_pipeline.py::test_paired_end_vs_oracle tests/test_gpu_traceback.py"
( time timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest $T -x -q ) > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck.log
grep -E "ERROR SUMMARY|Invalid|passed|failed|exit" gpurun_out/sanitizer_memcheck.log | tail -12
