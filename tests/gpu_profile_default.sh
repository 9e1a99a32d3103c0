#!/bin/bash
# ncu evidence for the DEFAULT bench command (100M x 384, batch 1, top-10, 1 GPU): per-launch list of the
# query kernels and one --set full capture of the dominant kernel.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-extras"
$CMD > gpurun_out/plain_default.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_scan_topk|k_mma_topk|k_merge_keys|k_prep_queries|k_exchange_merge" -c 50 --csv \
    --log-file gpurun_out/launches_default.csv $CMD > gpurun_out/ncu_default_list.log 2>&1
$CMD > gpurun_out/plain_default2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_scan_topk -s 4 -c 1 -f -o gpurun_out/prof_default $CMD > gpurun_out/ncu_default_full.log 2>&1
tail -n 1 gpurun_out/plain_default.log | cut -c1-300
