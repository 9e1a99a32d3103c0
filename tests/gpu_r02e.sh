#!/bin/bash
set -x
mkdir -p gpurun_out
T=r02e
O=gpurun_out
timeout 600 python -m pytest tests/test_mma_gpu.py tests/test_integration_doc_gpu.py -m gpu -x -q 2>&1 | tail -8 | tee $O/${T}_pytest_mma.log
CODDVEC_K3_DEBUG=1 PROBE_BATCHES=1,16,256,1024 PROBE_PATHS=mma python tests/quick_probe.py > $O/${T}_debug_1m.log 2>&1
grep -E "k3 debug" $O/${T}_debug_1m.log | sort | uniq -c | sort -rn | head -12
for mode in 0 1 3; do
  CODDVEC_K3_SPLIT_LAUNCHES=$mode PROBE_BATCHES=1,4,16,64,128,256,512,1024 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/split_launches=$mode /" | tee -a $O/${T}_probe_modes.log
done
PROBE_K=16 PROBE_BATCHES=16,256,1024 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/ew=8 /" | tee -a $O/${T}_probe_ew.log
CODDVEC_K3_EW=4 PROBE_K=16 PROBE_BATCHES=16,256,1024 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/ew=4 /" | tee -a $O/${T}_probe_ew.log
PROBE_ROWS=100000 PROBE_BATCHES=1,16,256 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | tee $O/${T}_probe_100k.log
PROBE_ROWS=12500000 PROBE_BATCHES=16,256,1024 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | tee $O/${T}_probe_12m.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/${T}_pytest_gpu.log
