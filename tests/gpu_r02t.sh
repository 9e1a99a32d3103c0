#!/bin/bash
# bound warp (refresh mode 3) against the per-thread modes and the previous commit
mkdir -p gpurun_out
O=gpurun_out
L=$O/r02t_ab.log
: > $L
run() { local label=$1; shift; env "$@" PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/$label /" | tee -a $L; }
BS=4,16,64,128,256,1024
for rows in 1000000 100000; do
for rep in 1 2; do
  run old_$rows PROBE_ROWS=$rows PROBE_BATCHES=$BS CODDVEC_LIB=$PWD/ab_libs/d43732d.so CODDVEC_LIB_ALLOW_MISSING=1
  run r3_$rows PROBE_ROWS=$rows PROBE_BATCHES=$BS
  run r3prep0_$rows PROBE_ROWS=$rows PROBE_BATCHES=$BS CODDVEC_K3_FUSED_PREP_MAX_B=0
  run r1_$rows PROBE_ROWS=$rows PROBE_BATCHES=$BS CODDVEC_K3_REFRESH=1
done
done
run old_12m PROBE_ROWS=12500000 PROBE_BATCHES=16,64,256,1024 CODDVEC_LIB=$PWD/ab_libs/d43732d.so CODDVEC_LIB_ALLOW_MISSING=1
run r3_12m PROBE_ROWS=12500000 PROBE_BATCHES=16,64,256,1024
run old_k50 PROBE_K=50 PROBE_BATCHES=16,64,256 CODDVEC_LIB=$PWD/ab_libs/d43732d.so CODDVEC_LIB_ALLOW_MISSING=1
run r3_k50 PROBE_K=50 PROBE_BATCHES=16,64,256
CODDVEC_K3_TRACE=1 CODDVEC_K3_TIMELINE=1 PROBE_BATCHES=4,64,256 PROBE_PATHS=mma python tests/quick_probe.py > $O/r02t_trace_raw.log 2>&1
for b in 4 64 256; do grep -A78 "timeline\] B=$b " $O/r02t_trace_raw.log | tail -79; done > $O/r02t_trace.log
timeout 900 python -m pytest tests/test_mma_gpu.py tests/test_cfg2_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee $O/r02t_pytest.log
