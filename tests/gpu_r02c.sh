#!/bin/bash
# attribution of the K3 fixed cost: fused prep / fused tail on and off in one binary, then ncu of the fused kernel
set -x
mkdir -p gpurun_out
T=r02c
O=gpurun_out
for mode in 0 1 2 3; do
  CODDVEC_K3_SPLIT_LAUNCHES=$mode PROBE_BATCHES=1,16,256 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/split_launches=$mode /" | tee -a $O/${T}_probe_modes.log
  CODDVEC_K3_SPLIT_LAUNCHES=$mode PROBE_ROWS=100000 PROBE_BATCHES=1,16 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/split_launches=$mode /" | tee -a $O/${T}_probe_modes.log
done
CMD="env PROBE_BATCHES=16 PROBE_PATHS=mma python tests/quick_probe.py"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mma_topk -s 6 -c 1 -f -o $O/${T}_prof_k3_b16 $CMD > $O/${T}_ncu_k3_b16.log 2>&1
timeout 900 python -m pytest tests/test_native_gpu.py tests/test_properties_gpu.py tests/test_recall_gpu.py tests/test_store_scenarios.py tests/test_integration_doc_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee $O/${T}_pytest_rest.log
