#!/bin/bash
mkdir -p gpurun_out
python tests/cross_process_check.py 2>&1 | tail -5 | tee gpurun_out/r02_cross_process_check.log
python -m pytest tests/test_client.py tests/test_batching.py tests/test_store_scenarios.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_crossproc_pytest.log
