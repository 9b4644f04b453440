#!/bin/sh
# end-of-round captures with the final code: launch list of the bench command, DRAM traffic per kernel, full capture of eval_dp
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-vector > gpurun_out/launches_bench.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/traffic.csv python tools/prof_keyword.py > gpurun_out/traffic.log 2>&1
