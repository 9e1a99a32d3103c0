#!/bin/bash
# PILOT as separate instantiations (chosen for single CTAs on short scans): A/B against the previous commit; parity; host probe
mkdir -p gpurun_out
O=gpurun_out
L=$O/r02y_ab.log
: > $L
run() { local label=$1; shift; env "$@" PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/$label /" | tee -a $L; }
PREV="CODDVEC_LIB=$PWD/ab_libs/49b0bb3.so CODDVEC_LIB_ALLOW_MISSING=1"
BS=4,16,64,128,256,1024
for rep in 1 2; do
  run prev_1m PROBE_BATCHES=$BS $PREV
  run new_1m PROBE_BATCHES=$BS
  run prev_100k PROBE_ROWS=100000 PROBE_BATCHES=$BS $PREV
  run new_100k PROBE_ROWS=100000 PROBE_BATCHES=$BS
  run prev_12m PROBE_ROWS=12500000 PROBE_BATCHES=16,64,256 $PREV
  run new_12m PROBE_ROWS=12500000 PROBE_BATCHES=16,64,256
done
run prev_3m PROBE_ROWS=3000000 PROBE_BATCHES=16,64,128 $PREV
run new_3m PROBE_ROWS=3000000 PROBE_BATCHES=16,64,128
run new_3m_nopilot PROBE_ROWS=3000000 PROBE_BATCHES=16,64,128 CODDVEC_K3_PILOT=0
timeout 1500 python -m pytest tests/test_mma_gpu.py tests/test_cfg2_gpu.py tests/test_properties_gpu.py tests/test_striped_gpu.py tests/test_recall_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee $O/r02y_pytest.log
python tests/host_scaling_probe.py 2>&1 | tail -4 | tee $O/r02y_host_scaling.jsonl
