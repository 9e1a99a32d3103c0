#!/bin/bash
# ncu evidence for the two hot kernels (one gpurun call; every ncu run is preceded by the same
# command exiting 0 without ncu).  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
ROWS=${ROWS:-12500000}
SCAN="python bench.py --rows $ROWS --steps 3 --warmup 3 --no-extras"
MMA="python bench.py --rows $ROWS --batch 1024 --steps 3 --warmup 3 --no-extras"
KERNELS='regex:k_scan_topk|k_mma_topk|k_merge_keys|k_prep_queries|k_normalise_append|k_count_nonfinite'
$SCAN > gpurun_out/plain_scan.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -c 400 --csv \
    --log-file gpurun_out/launches_scan.csv $SCAN > gpurun_out/ncu_scan_list.log 2>&1
$MMA > gpurun_out/plain_mma.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -c 400 --csv \
    --log-file gpurun_out/launches_mma.csv $MMA > gpurun_out/ncu_mma_list.log 2>&1
$SCAN > gpurun_out/plain_scan2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_scan_topk -s 2 -c 2 -f -o gpurun_out/prof_scan $SCAN > gpurun_out/ncu_scan_full.log 2>&1
$MMA > gpurun_out/plain_mma2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_mma_topk -s 2 -c 2 -f -o gpurun_out/prof_mma $MMA > gpurun_out/ncu_mma_full.log 2>&1
tail -2 gpurun_out/plain_scan.log gpurun_out/plain_mma.log
ls -la gpurun_out/
