#!/bin/bash
# Round-2 second GPU call: the whole GPU suite on the restructured K3 (query tile built in-kernel, fused tail,
# hi/lo split, self-resetting state) + merged K2 branches, then the cfg2 batch sweep.
set -x
mkdir -p gpurun_out
T=r02b
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/${T}_pytest_gpu.log
PROBE_BATCHES=1,2,4,8,16,64,128,256,512,1024 PROBE_PATHS=scan,mma python tests/quick_probe.py 2>&1 | tee $O/${T}_probe_1m.log
PROBE_ROWS=100000 PROBE_BATCHES=1,4,16,64 PROBE_PATHS=scan,mma python tests/quick_probe.py 2>&1 | tee $O/${T}_probe_100k.log
PROBE_ROWS=12500000 PROBE_BATCHES=16,256,1024 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | tee $O/${T}_probe_12m.log
PROBE_DIM=768 PROBE_BATCHES=1,2,16,256 PROBE_PATHS=scan,mma python tests/quick_probe.py 2>&1 | tee $O/${T}_probe_768.log
