#!/bin/bash
# Last sources of the round: the K3 capture behind the traffic stamp + the GPU suite.
mkdir -p gpurun_out
O=gpurun_out
T=r02j
P="env PROBE_ROWS=12500000 PROBE_BATCHES=1024 PROBE_PATHS=mma python tests/quick_probe.py"
ncu --set full --clock-control none --import-source on -k regex:k_mma_topk -s 3 -c 1 -f -o $O/${T}_prof_k3_12m_b1024 $P > $O/${T}_ncu_k3_12m_b1024.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/${T}_pytest_gpu.log
