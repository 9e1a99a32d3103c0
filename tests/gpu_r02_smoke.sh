#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_last_smoke.log
python -m pytest tests/test_cfg2_gpu.py tests/test_abi.py tests/test_integration_doc_gpu.py -m gpu -x -q 2>&1 | tail -2 | tee -a gpurun_out/r02_last_smoke.log
