#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_mma_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee $O/r02k_pytest_mma.log
: > $O/r02k_ab.log
for which in new old new old; do
  case $which in
    new) unset CODDVEC_LIB CODDVEC_LIB_ALLOW_MISSING;;
    old) export CODDVEC_LIB=$PWD/ab_libs/2988303.so CODDVEC_LIB_ALLOW_MISSING=1;;
  esac
  PROBE_BATCHES=1,4,16,64,128,256,1024 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/$which /" | tee -a $O/r02k_ab.log
  PROBE_ROWS=100000 PROBE_BATCHES=16,256 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/$which /" | tee -a $O/r02k_ab.log
done
unset CODDVEC_LIB CODDVEC_LIB_ALLOW_MISSING
for ew in 4 8; do
  CODDVEC_K3_EW=$ew PROBE_BATCHES=16,64,256,1024 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/ew=$ew /" | tee -a $O/r02k_ew.log
done
CODDVEC_K3_DEBUG=1 PROBE_BATCHES=16,64,256 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | grep "k3 debug" | sort | uniq -c | sort -rn | head -6 | tee $O/r02k_debug.log
