#!/bin/sh
# Round-2 evidence, one gpurun call (10 M documents; ~7 GPU-minutes).  gpurun brings back at most 64 MiB: the keyword capture is
# exported to CSV on the box and its .ncu-rep only kept when small enough.
#  1. launch list of two keyword batches (Detailed = the keyword side of the hybrid headline) with time + DRAM bytes per launch
#  2. the tcgen05 vector kernel at 1024 x 10^6 x 768 with the tensor-pipe counters, and the GEMV kernel
#  3. one --set full capture of the first launches of the heaviest keyword kernels (resolve + Words levels)
#  4. the cfg 4 matrix
set -x
mkdir -p gpurun_out
export DOCS=10000000 VOCAB=1500000
BATCHES=2 SCORING=detailed timeout 170 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__shared_mem_per_block_dynamic,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r02_list_10m.csv python tools/prof_keyword.py > gpurun_out/r02_list_10m.log 2>&1
VEC_GEMM_ONLY=1 timeout 70 ncu --set full --metrics sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --import-source on -k regex:vec_gemm_topk -s 2 -c 1 -f -o gpurun_out/r02_vec_gemm python tools/vec_bench.py > gpurun_out/r02_vec_gemm.log 2>&1
VEC_GEMV_ONLY=1 timeout 60 ncu --set full --clock-control none -k regex:vec_dist -s 2 -c 1 -f -o gpurun_out/r02_vec_dist python tools/vec_bench.py > gpurun_out/r02_vec_dist.log 2>&1
BATCHES=1 SCORING=detailed timeout 200 ncu --set full --clock-control none -k regex:"eval_dp_kernel|walk_kernel|scatter_kernel|scatter_big_kernel|act_compact_kernel" -c 14 -f -o gpurun_out/r02_keyword_full python tools/prof_keyword.py > gpurun_out/r02_keyword_full.log 2>&1
ncu -i gpurun_out/r02_keyword_full.ncu-rep --page raw --csv > gpurun_out/r02_keyword_full_raw.csv 2>/dev/null
[ "$(stat -c %s gpurun_out/r02_keyword_full.ncu-rep 2>/dev/null || echo 0)" -gt 30000000 ] && rm -f gpurun_out/r02_keyword_full.ncu-rep
timeout 150 python tools/cfg4_matrix.py > gpurun_out/r02_cfg4_matrix.json 2> gpurun_out/r02_cfg4_matrix.err
ls -la gpurun_out/; du -sh gpurun_out
