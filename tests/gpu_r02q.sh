#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
CODDVEC_K3_TRACE=1 CODDVEC_K3_TIMELINE=1 PROBE_BATCHES=4,64,256 PROBE_PATHS=mma python tests/quick_probe.py > $O/r02q_trace_raw.log 2>&1
for b in 4 64 256; do grep -A78 "timeline\] B=$b " $O/r02q_trace_raw.log | tail -79; done > $O/r02q_trace.log
