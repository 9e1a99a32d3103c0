#!/bin/bash
# ncu full capture of the pair-mode K3 kernel (batch 1024) on a 12.5M-row shard.
set -x
mkdir -p gpurun_out
MMA="python bench.py --rows ${ROWS:-12500000} --batch ${BATCH:-1024} --steps 3 --warmup 3 --no-extras"
$MMA > gpurun_out/plain_mma_pair.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_mma_topk -s 2 -c 1 -f -o gpurun_out/prof_mma_pair $MMA > gpurun_out/ncu_mma_pair.log 2>&1
tail -n 1 gpurun_out/plain_mma_pair.log | cut -c1-400
