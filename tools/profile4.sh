#!/bin/sh
# round-1 final captures: the heaviest eval_dp launch (source-level) and the tcgen05 vector kernel
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_dp_kernel -s 61 -c 2 -f -o gpurun_out/prof4_eval_dp python tools/prof_keyword.py > gpurun_out/prof4_eval.log 2>&1
VEC_GEMM_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:vec_gemm_topk -s 2 -c 1 -f -o gpurun_out/prof4_vec_gemm python tools/vec_bench.py > gpurun_out/prof4_vec_gemm.log 2>&1
