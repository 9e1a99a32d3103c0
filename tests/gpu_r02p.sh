#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
CODDVEC_K3_TIMELINE=1 PROBE_BATCHES=4,16,64,128,256,1024 PROBE_PATHS=mma python tests/quick_probe.py > $O/r02p_timeline_raw.log 2>&1
for b in 4 16 64 128 256; do grep -A13 "timeline\] B=$b " $O/r02p_timeline_raw.log | tail -14; done | tee $O/r02p_timeline.log
CODDVEC_K3_TIMELINE=1 PROBE_ROWS=12500000 PROBE_BATCHES=64,256,1024 PROBE_PATHS=mma python tests/quick_probe.py > $O/r02p_timeline12_raw.log 2>&1
for b in 64 256 1024; do grep -A13 "timeline\] B=$b " $O/r02p_timeline12_raw.log | tail -14; done | tee -a $O/r02p_timeline.log
