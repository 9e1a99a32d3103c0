#!/bin/bash
set -x
mkdir -p gpurun_out
CMD="env PROBE_BATCHES=64 PROBE_PATHS=mma python tests/quick_probe.py"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mma_topk -s 6 -c 1 -f -o gpurun_out/r02i_prof_k3_b64 $CMD > gpurun_out/r02i_ncu.log 2>&1
