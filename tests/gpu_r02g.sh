#!/bin/bash
set -x
mkdir -p gpurun_out
T=r02g
O=gpurun_out
OLD=$PWD/ab_libs/2988303.so
timeout 600 python -m pytest tests/test_striped_gpu.py -m gpu -x -q 2>&1 | tail -12 | tee $O/${T}_pytest_striped.log
for i in 1 2; do
  for which in new old; do
    if [ $which = old ]; then export CODDVEC_LIB=$OLD CODDVEC_LIB_ALLOW_MISSING=1; else unset CODDVEC_LIB CODDVEC_LIB_ALLOW_MISSING; fi
    PROBE_BATCHES=1,2,4,16,64,128,256,1024 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/$which /" | tee -a $O/${T}_ab_1m.log
    PROBE_ROWS=100000 PROBE_BATCHES=4,16,256 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/$which /" | tee -a $O/${T}_ab_100k.log
    PROBE_ROWS=12500000 PROBE_BATCHES=16,256,1024 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/$which /" | tee -a $O/${T}_ab_12m.log
  done
done
unset CODDVEC_LIB CODDVEC_LIB_ALLOW_MISSING
# assertion build (device-side bounds checks that trap) over the reduced subset + the K3 / K2 / striped suites
CODDVEC_LIB=$PWD/codd_query_engine_b200/libcoddvec_checked.so python tests/sanitizer_subset.py 2>&1 | tail -8 | tee $O/${T}_checked_subset.log
CODDVEC_LIB=$PWD/codd_query_engine_b200/libcoddvec_checked.so timeout 1200 python -m pytest tests/test_mma_gpu.py tests/test_native_gpu.py tests/test_properties_gpu.py tests/test_striped_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee $O/${T}_checked_pytest.log
timeout 900 python -m pytest tests/test_mma_gpu.py tests/test_client.py tests/test_store_scenarios.py tests/test_batching.py -m gpu -x -q 2>&1 | tail -5 | tee $O/${T}_pytest.log
