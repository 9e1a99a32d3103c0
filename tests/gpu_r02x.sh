#!/bin/bash
# pilot (two-pass first tile) on / off, against the previous commit's library; parity suites
mkdir -p gpurun_out
O=gpurun_out
L=$O/r02x_ab.log
: > $L
run() { local label=$1; shift; env "$@" PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/$label /" | tee -a $L; }
PREV="CODDVEC_LIB=$PWD/ab_libs/49b0bb3.so CODDVEC_LIB_ALLOW_MISSING=1"
BS=4,16,64,128,256,512,1024
for rep in 1 2; do
  run prev_1m PROBE_BATCHES=$BS $PREV
  run pilot_1m PROBE_BATCHES=$BS
  run nopilot_1m PROBE_BATCHES=$BS CODDVEC_K3_PILOT=0
  run prev_100k PROBE_ROWS=100000 PROBE_BATCHES=$BS $PREV
  run pilot_100k PROBE_ROWS=100000 PROBE_BATCHES=$BS
  run prev_12m PROBE_ROWS=12500000 PROBE_BATCHES=16,64,256,1024 $PREV
  run pilot_12m PROBE_ROWS=12500000 PROBE_BATCHES=16,64,256,1024
done
run prev_k50 PROBE_K=50 PROBE_BATCHES=16,64,256 $PREV
run pilot_k50 PROBE_K=50 PROBE_BATCHES=16,64,256
run prev_d768 PROBE_DIM=768 PROBE_BATCHES=16,64,256 $PREV
run pilot_d768 PROBE_DIM=768 PROBE_BATCHES=16,64,256
run prev_d64 PROBE_DIM=64 PROBE_BATCHES=16,64,256 $PREV
run pilot_d64 PROBE_DIM=64 PROBE_BATCHES=16,64,256
timeout 1500 python -m pytest tests/test_mma_gpu.py tests/test_cfg2_gpu.py tests/test_properties_gpu.py tests/test_striped_gpu.py tests/test_full_size_gpu.py tests/test_recall_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee $O/r02x_pytest.log
