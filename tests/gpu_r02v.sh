#!/bin/bash
# where does the 12.5M-row steady state lose 5 %?  bound warp on / off, PDL on / off, against the previous commit
mkdir -p gpurun_out
O=gpurun_out
L=$O/r02v_ab.log
: > $L
run() { local label=$1; shift; env "$@" PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/$label /" | tee -a $L; }
OLD="CODDVEC_LIB=$PWD/ab_libs/d43732d.so CODDVEC_LIB_ALLOW_MISSING=1"
for rep in 1 2; do
  run old_12m PROBE_ROWS=12500000 PROBE_BATCHES=16,64,256,1024 $OLD
  run new_12m PROBE_ROWS=12500000 PROBE_BATCHES=16,64,256,1024
  run r2_12m PROBE_ROWS=12500000 PROBE_BATCHES=16,64,256,1024 CODDVEC_K3_REFRESH=2
  run nopdl_12m PROBE_ROWS=12500000 PROBE_BATCHES=16,64,256,1024 CODDVEC_K3_PDL=0
  run prep16_12m PROBE_ROWS=12500000 PROBE_BATCHES=16,64 CODDVEC_K3_FUSED_PREP_MAX_B=16
done
run old_1m PROBE_BATCHES=4,16,64,128,256 $OLD
run new_1m PROBE_BATCHES=4,16,64,128,256
