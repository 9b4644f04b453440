#!/bin/sh
# Run on the GPU box (gpurun): ncu launch list of the bench command + one --set full capture of the top kernels.
# Outputs under gpurun_out/ ; summaries are copied into profiles/ by tools/summarise_profiles.py (run in the build container).
set -x
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-vector --docs ${DOCS:-1000000} --vocab ${VOCAB:-400000}"
# 1. every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
# 2. the heavy kernels, --set full, 3 launches each, skipping the first (small) launches
for K in eval_dp_kernel scatter_kernel lev_match_kernel act_compact_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 40 -c 3 -f -o gpurun_out/prof_$K $BENCH > gpurun_out/prof_$K.log 2>&1
done
# 3. the vector GEMV
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vec_dist_kernel -s 3 -c 2 -f -o gpurun_out/prof_vec_dist python tools/vec_bench.py > gpurun_out/prof_vec.log 2>&1
ls -la gpurun_out
