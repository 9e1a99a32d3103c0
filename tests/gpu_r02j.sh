#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
: > $O/r02j_variants.log
for which in new vA vB vC old new vC old; do
  case $which in
    new) unset CODDVEC_LIB CODDVEC_LIB_ALLOW_MISSING;;
    old) export CODDVEC_LIB=$PWD/ab_libs/2988303.so CODDVEC_LIB_ALLOW_MISSING=1;;
    *) export CODDVEC_LIB=$PWD/ab_libs/$which.so CODDVEC_LIB_ALLOW_MISSING=1;;
  esac
  PROBE_BATCHES=16,64,256 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/$which /" | tee -a $O/r02j_variants.log
done
