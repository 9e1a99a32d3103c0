#!/bin/bash
# predicate-scan / candidate-buffer iteration: K2 parity tests, then the selectivity probe
set -x
mkdir -p gpurun_out
python -m pytest tests/test_native_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_pred.log
CODDVEC_HYPOTHESIS_EXAMPLES=40 python -m pytest tests/test_properties_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee -a gpurun_out/pytest_pred.log
python tests/pred_probe.py 2>&1 | tee gpurun_out/pred_probe_${TAG:-new}.log
