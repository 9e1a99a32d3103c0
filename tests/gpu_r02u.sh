#!/bin/bash
# rank-based list insertion + bound warp + PDL'd query prep for every batch, against the previous commit
mkdir -p gpurun_out
O=gpurun_out
L=$O/r02u_ab.log
: > $L
run() { local label=$1; shift; env "$@" PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/$label /" | tee -a $L; }
BS=4,16,64,128,256,512,1024
OLD="CODDVEC_LIB=$PWD/ab_libs/d43732d.so CODDVEC_LIB_ALLOW_MISSING=1"
for rep in 1 2; do
  run old_1m PROBE_BATCHES=$BS $OLD
  run new_1m PROBE_BATCHES=$BS
  run old_100k PROBE_ROWS=100000 PROBE_BATCHES=$BS $OLD
  run new_100k PROBE_ROWS=100000 PROBE_BATCHES=$BS
done
for rep in 1 2 3; do
  run old_12m PROBE_ROWS=12500000 PROBE_BATCHES=16,64,256,1024 $OLD
  run new_12m PROBE_ROWS=12500000 PROBE_BATCHES=16,64,256,1024
done
run old_k50 PROBE_K=50 PROBE_BATCHES=16,64,256 $OLD
run new_k50 PROBE_K=50 PROBE_BATCHES=16,64,256
run old_d768 PROBE_DIM=768 PROBE_BATCHES=16,64,256 $OLD
run new_d768 PROBE_DIM=768 PROBE_BATCHES=16,64,256
run old_d64 PROBE_DIM=64 PROBE_BATCHES=16,64,256 $OLD
run new_d64 PROBE_DIM=64 PROBE_BATCHES=16,64,256
CODDVEC_K3_TRACE=1 CODDVEC_K3_TIMELINE=1 PROBE_BATCHES=256 PROBE_PATHS=mma python tests/quick_probe.py > $O/r02u_trace_raw.log 2>&1
grep -A78 "timeline\] B=256 " $O/r02u_trace_raw.log | tail -79 > $O/r02u_trace.log
timeout 1200 python -m pytest tests/test_mma_gpu.py tests/test_cfg2_gpu.py tests/test_properties_gpu.py tests/test_striped_gpu.py tests/test_recall_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee $O/r02u_pytest.log
