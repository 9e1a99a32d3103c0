#!/bin/bash
# first GPU contact: native parity tests + a quick timing probe
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests/test_native_gpu.py -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_native.log
python tests/quick_probe.py 2>&1 | tee gpurun_out/probe.log
