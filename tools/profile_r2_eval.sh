#!/bin/sh
# round 2: full ncu capture of the first eval_dp / scatter launches of a 10 M-document keyword batch (resolve + Words levels)
set -x
mkdir -p gpurun_out
export DOCS=${DOCS:-10000000} VOCAB=${VOCAB:-1500000} BATCHES=1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"${KERNEL:-eval_dp}" -c ${COUNT:-4} -f -o gpurun_out/r2_${TAG:-eval10m} python tools/prof_keyword.py > gpurun_out/r2_${TAG:-eval10m}.log 2>&1
tail -2 gpurun_out/r2_${TAG:-eval10m}.log
