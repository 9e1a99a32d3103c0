#!/bin/bash
# cleaned-up K3 (bound warp only, rank insertion, timeline stamps compiled out): whole GPU suite + A/B + assertion build
mkdir -p gpurun_out
O=gpurun_out
L=$O/r02w_ab.log
: > $L
run() { local label=$1; shift; env "$@" python tests/quick_probe.py 2>&1 | sed "s/^/$label /" | tee -a $L; }
OLD="CODDVEC_LIB=$PWD/ab_libs/d43732d.so CODDVEC_LIB_ALLOW_MISSING=1"
for rep in 1 2; do
  run old_1m PROBE_PATHS=mma PROBE_BATCHES=4,16,64,128,256,512,1024 $OLD
  run new_1m PROBE_PATHS=mma PROBE_BATCHES=4,16,64,128,256,512,1024
  run old_12m PROBE_PATHS=mma PROBE_ROWS=12500000 PROBE_BATCHES=16,64,256,1024 $OLD
  run new_12m PROBE_PATHS=mma PROBE_ROWS=12500000 PROBE_BATCHES=16,64,256,1024
done
run scan_1m PROBE_PATHS=scan PROBE_BATCHES=1,2,4
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/r02w_pytest_gpu.log
CODDVEC_LIB=$PWD/codd_query_engine_b200/libcoddvec_checked.so python tests/sanitizer_subset.py 2>&1 | tail -3 | tee $O/r02w_checked_subset.log
CODDVEC_LIB=$PWD/codd_query_engine_b200/libcoddvec_checked.so timeout 1200 python -m pytest tests/test_mma_gpu.py tests/test_native_gpu.py tests/test_properties_gpu.py tests/test_striped_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee $O/r02w_checked_pytest.log
