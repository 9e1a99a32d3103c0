#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
: > $O/r02m_exact.log
for thr in 1 100000 1 100000; do
  CODDVEC_K3_EXACT_HEADS_MIN_B=$thr PROBE_BATCHES=16,32,64,128,256,512,1024 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/exact_min_b=$thr /" | tee -a $O/r02m_exact.log
  CODDVEC_K3_EXACT_HEADS_MIN_B=$thr PROBE_ROWS=12500000 PROBE_BATCHES=64,256,1024 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/exact_min_b=$thr /" | tee -a $O/r02m_exact.log
done
for ew in 4 8; do
  CODDVEC_K3_EW=$ew PROBE_BATCHES=256,512,1024 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/ew=$ew /" | tee -a $O/r02m_exact.log
  CODDVEC_K3_EW=$ew PROBE_ROWS=12500000 PROBE_BATCHES=256,1024 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/ew=$ew /" | tee -a $O/r02m_exact.log
done
timeout 600 python -m pytest tests/test_mma_gpu.py tests/test_cfg2_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee $O/r02m_pytest.log
