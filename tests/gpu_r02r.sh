#!/bin/bash
# once-per-tile cross-CTA publications: A/B against the previous commit, per-tile trace, parity tests
mkdir -p gpurun_out
O=gpurun_out
L=$O/r02r_ab.log
: > $L
run() { local label=$1; shift; env "$@" PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/$label /" | tee -a $L; }
for rep in 1 2; do
  run new PROBE_BATCHES=4,16,32,64,128,256,512,1024
  run old PROBE_BATCHES=4,16,32,64,128,256,512,1024 CODDVEC_LIB=$PWD/ab_libs/d43732d.so CODDVEC_LIB_ALLOW_MISSING=1
done
for rep in 1 2; do
  run new12m PROBE_ROWS=12500000 PROBE_BATCHES=16,64,256,1024
  run old12m PROBE_ROWS=12500000 PROBE_BATCHES=16,64,256,1024 CODDVEC_LIB=$PWD/ab_libs/d43732d.so CODDVEC_LIB_ALLOW_MISSING=1
done
run new100k PROBE_ROWS=100000 PROBE_BATCHES=16,64,256
run old100k PROBE_ROWS=100000 PROBE_BATCHES=16,64,256 CODDVEC_LIB=$PWD/ab_libs/d43732d.so CODDVEC_LIB_ALLOW_MISSING=1
run newk50 PROBE_K=50 PROBE_BATCHES=16,64,256
run oldk50 PROBE_K=50 PROBE_BATCHES=16,64,256 CODDVEC_LIB=$PWD/ab_libs/d43732d.so CODDVEC_LIB_ALLOW_MISSING=1
run newd768 PROBE_DIM=768 PROBE_BATCHES=16,64,256
run oldd768 PROBE_DIM=768 PROBE_BATCHES=16,64,256 CODDVEC_LIB=$PWD/ab_libs/d43732d.so CODDVEC_LIB_ALLOW_MISSING=1
CODDVEC_K3_TRACE=1 CODDVEC_K3_TIMELINE=1 PROBE_BATCHES=4,64,256 PROBE_PATHS=mma python tests/quick_probe.py > $O/r02r_trace_raw.log 2>&1
for b in 4 64 256; do grep -A78 "timeline\] B=$b " $O/r02r_trace_raw.log | tail -79; done > $O/r02r_trace.log
timeout 900 python -m pytest tests/test_mma_gpu.py tests/test_cfg2_gpu.py tests/test_properties_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee $O/r02r_pytest.log
