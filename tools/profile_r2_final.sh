#!/bin/sh
# Round-2 evidence, one gpurun call (10 M documents; needs ~6 GPU-minutes):
#  1. launch list of two keyword batches (Detailed = the keyword side of the hybrid headline) with time + DRAM bytes per launch
#  2. --set full captures of the heaviest kernels of the first steps (resolve + Words levels)
#  3. the tcgen05 vector kernel at 1024 x 10^6 x 768 with the tensor-pipe counters, and the GEMV kernel
set -x
mkdir -p gpurun_out
export DOCS=10000000 VOCAB=1500000
BATCHES=2 SCORING=detailed timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__shared_mem_per_block_dynamic,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r02_list_10m.csv python tools/prof_keyword.py > gpurun_out/r02_list_10m.log 2>&1
for k in eval_dp_kernel walk_kernel scatter_kernel scatter_big_kernel act_compact_kernel; do
  BATCHES=1 SCORING=detailed timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^(void )?$k" -c 3 -f -o gpurun_out/r02_$k python tools/prof_keyword.py > gpurun_out/r02_$k.log 2>&1
done
VEC_GEMM_ONLY=1 timeout 600 ncu --set full --metrics sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --import-source on -k regex:vec_gemm_topk -s 2 -c 1 -f -o gpurun_out/r02_vec_gemm python tools/vec_bench.py > gpurun_out/r02_vec_gemm.log 2>&1
VEC_GEMV_ONLY=1 timeout 600 ncu --set full --clock-control none -k regex:vec_dist -s 2 -c 1 -f -o gpurun_out/r02_vec_dist python tools/vec_bench.py > gpurun_out/r02_vec_dist.log 2>&1
ls -la gpurun_out/r02_*
