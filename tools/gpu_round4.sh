#!/bin/bash
mkdir -p gpurun_out
./tools/ubench_gather > gpurun_out/ubench_gather2.log 2>&1
ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:gather --csv --log-file gpurun_out/ubench_gather2_ncu.csv ./tools/ubench_gather > /dev/null 2>&1
cat gpurun_out/ubench_gather2.log
