#!/bin/bash
# ncu full capture of the predicate scan (cfg3: 10M x 768, 10 % namespace selectivity, top-50)
set -x
mkdir -p gpurun_out
CMD="python tests/bench_configs.py cfg3 --rows3 ${ROWS3:-4000000}"
$CMD > gpurun_out/plain_cfg3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_scan_topk -s 20 -c 1 -f -o gpurun_out/prof_cfg3 $CMD > gpurun_out/ncu_cfg3.log 2>&1
tail -n 1 gpurun_out/plain_cfg3.log | cut -c1-600
