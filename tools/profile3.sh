#!/bin/sh
# DRAM traffic of every launch of one batch, per kernel (for roofline.traffic), plus a full capture of the tcgen05 vector kernel.
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/traffic.csv python tools/prof_keyword.py > gpurun_out/traffic.log 2>&1
VEC_GEMM_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:vec_gemm_topk -s 2 -c 1 -f -o gpurun_out/prof_vec_gemm python tools/vec_bench.py > gpurun_out/prof_vec_gemm.log 2>&1
