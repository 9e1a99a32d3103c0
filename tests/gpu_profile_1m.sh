#!/bin/bash
# ncu full capture of the batch-1 scan at BASELINE configs[1] size (1M x 384)
set -x
mkdir -p gpurun_out
CMD="python bench.py --rows 1000000 --steps 5 --warmup 3 --no-extras"
$CMD > gpurun_out/plain_1m.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_scan_topk -s 4 -c 1 -f -o gpurun_out/prof_1m $CMD > gpurun_out/ncu_1m.log 2>&1
tail -n 1 gpurun_out/plain_1m.log | cut -c1-200
