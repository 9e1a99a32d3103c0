#!/bin/bash
set -x
mkdir -p gpurun_out
T=r02d
O=gpurun_out
CODDVEC_K3_DEBUG=1 PROBE_BATCHES=1,16,256,1024 PROBE_PATHS=mma python tests/quick_probe.py > $O/${T}_debug_1m.log 2>&1
grep -E "k3 debug" $O/${T}_debug_1m.log | sort | uniq -c | sort -rn | head -30
for mode in 0 1 3; do
  CODDVEC_K3_SPLIT_LAUNCHES=$mode PROBE_BATCHES=1,4,16,64,256,1024 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/split_launches=$mode /" | tee -a $O/${T}_probe_modes.log
  CODDVEC_K3_SPLIT_LAUNCHES=$mode PROBE_ROWS=100000 PROBE_BATCHES=1,16 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/split_launches=$mode /" | tee -a $O/${T}_probe_modes.log
done
timeout 900 python -m pytest tests/test_mma_gpu.py tests/test_integration_doc_gpu.py tests/test_native_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee $O/${T}_pytest.log
