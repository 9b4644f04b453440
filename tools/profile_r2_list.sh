#!/bin/sh
# round 2: every launch of two keyword batches (single lane) with duration, DRAM traffic, grid and dynamic shared memory
set -x
mkdir -p gpurun_out
BATCHES=2 timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__shared_mem_per_block_dynamic,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r2_list_${TAG:-a}.csv python tools/prof_keyword.py > gpurun_out/r2_list_${TAG:-a}.log 2>&1
tail -3 gpurun_out/r2_list_${TAG:-a}.log
