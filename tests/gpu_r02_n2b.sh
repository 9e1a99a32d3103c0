#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_striped_gpu.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r02_n2b_striped.log
tail -5 gpurun_out/r02_n2b_striped.log
CODDVEC_DEVICES=0,1 python -m pytest tests/test_store_scenarios.py tests/test_client.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02_n2b_store_on_two_gpus.log
