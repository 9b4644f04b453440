#!/bin/sh
# ncu --set full of the heavy eval_dp / scatter launches (launches 1..8 of the third identical batch; 30 steps per batch)
set -x
mkdir -p gpurun_out
S=${SKIP:-61}
for K in eval_dp_kernel scatter_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 7 -f -o gpurun_out/prof2_$K python tools/prof_keyword.py > gpurun_out/prof2_$K.log 2>&1
done
